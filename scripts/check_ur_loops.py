#!/usr/bin/env python
"""Static check of a kernel's SASS for the hazard documented in kicp_device.cuh: a uniform register (URx, shared by
the 32 lanes of a warp) that is READ inside a loop body, not written inside it, and WRITTEN AGAIN LATER in the kernel
(at an address after the loop).  Lanes that leave a per-lane loop early can reach that later write while their siblings
are still looping and still reading the old value.  A uniform register with no write after the loop (kernel
parameters loaded once, e.g. the state pointer) cannot be clobbered and is not reported.  Memory descriptors
(desc[URx]) are ignored."""
import re, subprocess, sys
lib, fun = sys.argv[1], sys.argv[2]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], capture_output=True, text=True).stdout
ins = []
for l in sass.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for a, t in ins:
    m = re.search(r"BRA\s+(?:P\d+,\s*)?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) <= a: loops.append((int(m.group(1), 16), a))
def writes(t):
    m = re.match(r"(?:@!?U?P\d+\s+)?\S+\s+(UR\d+)", t)
    w = set()
    if m:
        w.add(m.group(1))
        if ".64" in t.split()[0] or "LDCU.64" in t: w.add("UR%d" % (int(m.group(1)[2:]) + 1))
    return w
bad = 0
for lo, hi in sorted(set(loops)):
    body = [(a, t) for a, t in ins if lo <= a <= hi]
    # a loop that contains a full-warp or CTA barrier is warp-synchronous by construction: no lane can leave it early
    # (__syncwarp() compiles to WARPSYNC.ALL, or to BRA.DIV -> a warpsync stub when the compiler expects convergence)
    if any(("WARPSYNC.ALL" in t) or ("BRA.DIV" in t) or t.startswith("BAR.SYNC") or (" BAR.SYNC" in t) for a, t in body):
        continue
    # the same holds for a loop whose body executes a full-warp collective unconditionally (SHFL / VOTE / REDUX / MATCH with
    # all 32 lanes named: the hardware holds every lane at the instruction until all have arrived), which is how the pooled
    # registration kernel is written — ptxas proves those loops convergent and drops the explicit WARPSYNC
    def collective(t):
        if t.startswith("@"):
            return False  # predicated: not executed by every lane
        op = t.split()[0]
        return op.startswith(("SHFL.", "VOTE.", "VOTEU.", "REDUX", "CREDUX", "MATCH."))
    if any(collective(t) for a, t in body):
        continue
    # a counted loop whose trip count is a uniform register (back-edge predicate = ISETP counter, URx): every lane runs the
    # same number of iterations, nobody leaves early
    back = ins[addr[hi]][1]
    mb = re.match(r"@(!?)(P\d+)\s+BRA", back)
    if mb:
        setter = None
        for a, t in body:
            ms = re.match(r"(?:@!?U?P\d+\s+)?ISETP\S*\s+(P\d+),", t)
            if ms and ms.group(1) == mb.group(2): setter = t
        if setter is not None and re.search(r",\s*UR\d+,", setter):
            continue
    defs = set()
    for a, t in body: defs |= writes(t)
    later = set()
    for a, t in ins:
        if a > hi: later |= writes(t)
    reads = {}
    for a, t in body:
        t2 = re.sub(r"^(?:@!?U?P\d+\s+)?\S+\s+UR\d+", "", t) if writes(t) else t
        for r in re.findall(r"(?<!desc\[)UR\d+", t2):
            if r not in defs and r in later: reads.setdefault(r, []).append((a, t))
    if reads:
        bad += 1
        print("loop 0x%x..0x%x (%d instr) reads uniform regs that are rewritten after the loop:" % (lo, hi, len(body)))
        for r, uses in reads.items():
            print("   %s: %s" % (r, "; ".join("0x%x %s" % (a, t[:50]) for a, t in uses[:2])))
print("loops: %d, hazardous: %d" % (len(set(loops)), bad))

#!/usr/bin/env python
"""Static check of a kernel's SASS: list uniform-register (URx) READS inside loop bodies (a backward branch to an
earlier address) whose defining write is outside the loop.  Such reads are only safe if the loop is warp-uniform; for
per-lane loops they are the hazard documented in kicp_device.cuh.  UR used as memory descriptors (desc[URx]) are
reported separately (they hold a constant descriptor for the whole kernel)."""
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
    defs = set()
    for a, t in body: defs |= writes(t)
    reads = {}
    for a, t in body:
        t2 = re.sub(r"^(?:@!?U?P\d+\s+)?\S+\s+UR\d+", "", t) if writes(t) else t
        for r in re.findall(r"(?<!desc\[)UR\d+", t2):
            if r not in defs: reads.setdefault(r, []).append((a, t))
    if reads:
        bad += 1
        print("loop 0x%x..0x%x (%d instr) reads loop-invariant uniform regs:" % (lo, hi, len(body)))
        for r, uses in reads.items():
            print("   %s: %s" % (r, "; ".join("0x%x %s" % (a, t[:50]) for a, t in uses[:2])))
print("loops: %d, with loop-invariant UR reads: %d" % (len(set(loops)), bad))

#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` capture of the registration kernel (read here, no GPU needed): DRAM bytes
per launch, stamped with the SHA-256 of the kernel source it was captured from — bench.py reports `traffic` only when that hash
matches the kicp_register.cu it is running.   usage: python scripts/make_ncu_traffic.py gpurun_out/reg_r02_final.ncu-rep"""
import csv, hashlib, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U, V = rows[0], rows[1], rows[2]
def get(name):
    i = H.index(name)
    v, u = float(V[i]), U[i]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
src = os.path.join(ROOT, "kinematic-icp_b200", "csrc", "kicp_register.cu")
out = {"kernel": "k_register<true>", "workload": "cfg4, one launch = one registration (4 passes), right after the bench's L2 flush",
       "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
       "dram_bytes_per_launch": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
       "lts_sectors": get("lts__t_sectors.sum"), "lts_bytes_per_launch": get("lts__t_sectors.sum") * 32,
       "duration_us_under_ncu": get("gpu__time_duration.sum") / 1e3 if U[H.index("gpu__time_duration.sum")] == "ns" else get("gpu__time_duration.sum"),
       "source": os.path.basename(rep) + " (ncu --set full --clock-control none, bench.py --steps 2 --warmup 1)",
       "kernel_source_sha256": hashlib.sha256(open(src, "rb").read()).hexdigest()}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

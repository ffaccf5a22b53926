"""Per-kernel totals and shares from an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv
--log-file LIST.csv python bench.py ...`).  Per-launch times under ncu are cold-cache and serialised: the SHARES are what
is compared with bench.py's CUDA-event timing, not the absolutes.

    python profiles/launch_list_summary.py LIST.csv "COMMAND THAT WAS PROFILED" > profiles/rN_launches_<workload>.txt
"""
import csv, re, sys
from collections import defaultdict

path, cmd = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(path, errors="ignore") if l.startswith('"'))]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void ", "", r[ik])
    name = re.sub(r"\(.*$", "", name).replace("dmi::", "")
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {cmd}")
print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes.")
print("# launches  total_ms  share  kernel")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{cnt[k]:6d} {tot[k]:10.3f} {100 * tot[k] / total:6.2f}%  {k}")
print(f"# total {total:.3f} ms")

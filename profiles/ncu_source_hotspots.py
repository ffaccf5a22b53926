"""Per-instruction view of an `ncu --set full --import-source on` report: groups the SASS instructions of the first
kernel by their executed count (a loop body shows up as one group), prints each group's share of the executed
instructions and of the warp-stall samples, the instruction mix of the biggest group and the most-stalled
instructions outside it.   usage: python profiles/ncu_source_hotspots.py gpurun_out/<name>.ncu-rep"""
import csv
import re
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, isrc = hdr.index("Address"), hdr.index("Source")
ist, iex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
print("#", rows[0][1][:120])
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) > iex and r[ia].startswith("0x"):
        data.append((r[isrc].strip(), int(r[ist] or 0), int(r[iex] or 0)))
tot_ex, tot_st = sum(d[2] for d in data), sum(d[1] for d in data)
print(f"# {len(data)} SASS instructions, {tot_ex:.3e} executed (warp level), {tot_st} stall samples")
groups = {}
for s, smp, ex in data:
    g = groups.setdefault(ex, [0, 0])
    g[0] += 1
    g[1] += smp
print("# executed-count  #instr  share-of-executed  share-of-stall-samples")
for ex, (n, smp) in sorted(groups.items(), key=lambda kv: -kv[1][1])[:10]:
    print(f"{ex:14d} {n:7d} {100.0 * ex * n / tot_ex:17.1f}% {100.0 * smp / tot_st:22.1f}%")
hot = max(groups, key=lambda ex: ex * groups[ex][0])
mix, st = Counter(), Counter()
for s, smp, ex in data:
    if ex == hot:
        op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0]
        mix[op] += 1
        st[op] += smp
print(f"# instruction mix of the hottest group ({groups[hot][0]} instructions x {hot} executions):")
for op, n in mix.most_common():
    print(f"{op:8s} {n:4d}   {100.0 * st[op] / max(groups[hot][1], 1):5.1f}% of the group's stall samples")
print("# most-stalled instructions outside that group:")
for s, smp, ex in sorted((d for d in data if d[2] != hot), key=lambda d: -d[1])[:12]:
    print(f"{100.0 * smp / tot_st:5.1f}%  executed {ex:10d}  {s[:100]}")

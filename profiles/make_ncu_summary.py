"""Turns an `ncu --set full` capture (.ncu-rep) into an entry of profiles/ncu_summary.json, the file bench.py reads its
`roofline.traffic` / `roofline.ncu` from (nothing measured is pasted into bench.py).

    python profiles/make_ncu_summary.py KEY REPORT.ncu-rep "COMMAND THAT WAS PROFILED" [KERNEL_MS_PER_STEP_OF_THE_SAME_BUILD]

KEY = "<workload>/<kernel>/n<gpus>" for the integration kernel (e.g. config5/auto/n1) or "coloration/colorize/n1".
Values are means over the captured launches."""
import csv, json, os, subprocess, sys

key, rep, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
kernel_ms = float(sys.argv[4]) if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(name):
    i = hdr.index(name)
    vals = []
    for r in data:
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        if u.startswith("gbyte"): v *= 1e9
        elif u.startswith("mbyte"): v *= 1e6
        elif u.startswith("kbyte"): v *= 1e3
        elif u in ("ms", "msecond"): v *= 1e-3
        elif u in ("us", "usecond"): v *= 1e-6
        elif u in ("ns", "nsecond"): v *= 1e-9
        vals.append(v)
    return sum(vals) / len(vals)


pct = lambda n: col(n) / 100.0
stall = lambda n: col(f"smsp__average_warps_issue_stalled_{n}_per_issue_active.ratio")
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
entry = {
    "kernel": data[0][hdr.index("Kernel Name")].split("(")[0],
    "launches_captured": len(data),
    "dram_bytes_per_launch": col("dram__bytes_read.sum") + col("dram__bytes_write.sum"),
    "gpu_time_s_per_launch_under_ncu": col("gpu__time_duration.sum"),
    "registers_per_thread": col("launch__registers_per_thread"),
    "warps_active_frac": pct("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "issue_slots_busy": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_eligible_per_cycle": col("smsp__warps_eligible.avg.per_cycle_active"),
    "pipe_fp64": pct("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "pipe_fma": pct("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "pipe_alu": pct("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "pipe_xu": pct("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    "pipe_lsu": pct("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    "l1_hit": pct("l1tex__t_sector_hit_rate.pct"), "l2_hit": pct("lts__t_sector_hit_rate.pct"),
    "warp_instructions_per_launch": col("smsp__inst_executed.sum"),
    "stalls_per_issue": {n: stall(n) for n in ("long_scoreboard", "wait", "barrier", "not_selected", "short_scoreboard",
                                               "math_pipe_throttle", "branch_resolving", "no_instruction")},
    "command": cmd, "report": os.path.basename(rep), "summary_written_at_commit": commit,
    "kernel_ms_per_step_of_this_build": kernel_ms,
    "note": "ncu replays each launch with cold caches and serialised: durations are not bench values",
}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_summary.json")
allv = json.load(open(path)) if os.path.exists(path) else {}
allv[key] = entry
json.dump(allv, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(entry, indent=1))

"""Turns the ncu exports in gpurun_out/ into the tracked summaries under profiles/ (run in the dev container).

    python scripts/summarize_profiles.py r01
"""
import csv, os, re, shutil, sys
from collections import defaultdict

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)

def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")
    return name

lines = []
# ---- launch list of the bench command
src = os.path.join(G, f"{R}_launches_bench.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(OUT, f"{R}_launches_bench.csv"))
    rows = [l for l in open(src) if not l.startswith("==")]
    agg = defaultdict(list)
    for row in csv.DictReader(rows):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
            v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
            agg[short(row["Kernel Name"])].append(v)
    tot = sum(sum(v) for v in agg.values())
    lines += [f"## Launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (ncu, cold cache, serialised)", "",
              f"command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv` — raw file `{R}_launches_bench.csv`; "
              "per-launch times are serialised and cold-cache: compare SHARES.", "",
              "| kernel | launches | total µs | share | median µs |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| `{k}` | {len(v)} | {sum(v):.1f} | {100*sum(v)/tot:.1f} % | {sorted(v)[len(v)//2]:.1f} |")
    lines.append("")

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
]

def table(raw, title, note):
    if not os.path.exists(raw):
        return
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    # one column per distinct kernel (first occurrence after warm-up = last occurrence)
    pick = {}
    for r in data:
        pick[short(r[ix["Kernel Name"]])] = r
    names = list(pick)
    lines.extend([f"## {title}", "", note, "", "| metric | " + " | ".join(f"`{n}`" for n in names) + " |", "|---|" + "---:|" * len(names)])
    for key, label in WANT:
        if key not in ix:
            continue
        vals = []
        for n in names:
            v = pick[n][ix[key]]
            try:
                vals.append(f"{float(v.replace(',', '')):.4g} {units[ix[key]]}".strip())
            except ValueError:
                vals.append(v)
        lines.append(f"| {label} | " + " | ".join(vals) + " |")
    lines.append("")

table(os.path.join(G, f"{R}_query_point.raw.csv"), "`k_query_point` at bench size (16 777 216 queries, C2 mesh, Morton-ordered batch)",
      "command: `ncu --set full --clock-control none --import-source on -k regex:'k_query_point|k_unpack_results' -s 2 -c 2 python scripts/prof_driver.py` "
      "(PROF_NQ=2^24).  `DRAM read + write` of the traversal launch plus its unpack pass is the `roofline.traffic` figure bench.py reports.")
table(os.path.join(G, f"{R}_build_refit.raw.csv"), "Build / refit / ray kernels at C2 size (1 310 720 triangles; rays: 1 M random rays)",
      "command: `ncu --set full --clock-control none --import-source on -k regex:'k_scene|k_morton|k_onesweep|k_leaves|k_merge|k_refit|k_deep|k_query_ray' "
      "-s 12 -c 16 python scripts/prof_driver.py`.")

table(os.path.join(G, f"{R}_build10m.raw.csv"), "Build kernels at C3 size (9 999 392-triangle heightfield, 30-bit keys)",
      "command: `ncu --set full --clock-control none --import-source on -k regex:'k_onesweep|k_merge|k_leaves|k_scene|k_morton|k_deep' -s 11 -c 11 "
      "python scripts/prof_build_big.py` (taken before the depth pass was reworked: `k_deep_fix` reads packed range lengths since).")
table(os.path.join(G, f"{R}_refit_wave.raw.csv"), "Wavefront refit + its one-time plan at C4 size (3 998 792-triangle cloth)",
      "command: `ncu --set full --clock-control none --import-source on -k regex:'k_refit|k_plan' -s 0 -c 12 python scripts/prof_refit_driver.py`. "
      "`k_plan_*` (+ 4 sort passes, not captured here) run once per build; a refit is `k_refit_leaves` + `k_refit_levels` + `k_refit_climb`.")

table(os.path.join(G, f"{R}_query_ray.raw.csv"), "`k_query_ray` on C3 (9 999 392-triangle heightfield, 4096 x 4096 primary rays in row-major order)",
      "command: `ncu --set full --clock-control none --import-source on -k regex:k_query_ray -s 2 -c 1 python scripts/prof_ray_driver.py`.")

open(os.path.join(OUT, f"{R}_summary.md"), "w").write(f"# ncu summaries, round {R}\n\n" + "\n".join(lines) + "\n")
print("wrote", os.path.join(OUT, f"{R}_summary.md"))

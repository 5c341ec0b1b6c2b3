"""Turns the round-2 measurement artefacts in gpurun_out/ into the tracked summaries under profiles/ (needs no GPU).
usage: python scripts/summarize_r02.py

  (produced by `gpurun -- bash scripts/profile_pass.sh r2f`)
  gpurun_out/r2f_bench.json, r2f_bench_ref.json        -> profiles/r02_bench_n1.json, r02_bench_reference_arm.json
  gpurun_out/r2f_launches_c3.csv                        -> profiles/r02_launches_c3.csv (+ _summary.csv)
  gpurun_out/r2f_prof_{c3,c3_primary,c4,c5}.ncu-rep
                                                        -> profiles/r02_ncu_full_{c3,c4,c5}.csv, the *_traffic_*.json files bench.py reads,
                                                           profiles/r02_budget_{bounce_small,primary,trace8,volume_paths}.md (phase budgets)
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"
KEEP = ["ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def last_json(path):
    return [l for l in Path(path).read_text().splitlines() if l.startswith("{")][-1]


def budget(rep, regex, index, symbol, phases, units=None):
    cmd = [sys.executable, str(ROOT / "scripts" / "ncu_phase_budget.py"), str(rep), regex, str(index), symbol, str(P / "phases" / phases)]
    if units:
        cmd.append(str(units))
    return subprocess.run(cmd, capture_output=True, text=True).stdout


for src, dst in (("r2f_bench.json", "r02_bench_n1.json"), ("r2f_bench_ref.json", "r02_bench_reference_arm.json"), ("r2f_bench_short.json", "r02_bench_short_for_launch_list.json")):
    if (G / src).exists():
        (P / dst).write_text(last_json(G / src) + "\n")

f = G / "r2f_launches_c3.csv"
if f.exists():
    rows = list(csv.reader(open(f)))
    start = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["# ncu launch list summary — bench.py --steps 2 --warmup 3 --spp 64 --side= --no-cpu-baseline --no-e2e (workload c3, two waves of 32 spp per step), round 2 final build",
           "# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised: compare SHARES",
           f"# launches captured: {sum(v[0] for v in agg.values())}, total {tot / 1e3:.2f} ms", "kernel,launches,total_us,share"]
    out += [f"{k},{v[0]},{v[1]:.1f},{v[1] / tot:.4f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    (P / "r02_launches_c3_summary.csv").write_text("\n".join(out) + "\n")
    (P / "r02_launches_c3.csv").write_bytes(f.read_bytes())
    print("\n".join(out[2:9]))

WL = {"c3": ("r2f_prof_c3.ncu-rep", ("k_bounce_small",), "bounce_traffic", "scripts/profile_step.py c3 8 (one wave of 8 spp = 16.6 M paths): k_bounce_small at bounce 0, 1, 2"),
      "c4": ("r2f_prof_c4.ncu-rep", ("k_trace8<(bool)0", "k_trace8<0"), "extend_traffic", "scripts/profile_step.py c4 8 (one wave of 8 spp): k_raygen, then k_trace8 closest (<0,..>) / k_shade_surface / k_trace8 any-hit (<1,..>) at bounce 0, 1, 2"),
      "c5": ("r2f_prof_c5.ncu-rep", ("k_volume_paths",), "volume_traffic", "scripts/profile_step.py c5 8 (one wave of 8 spp): k_volume_paths (every volume path to completion)")}
for wl, (repname, pats, stem, title) in WL.items():
    rep = G / repname
    if not rep.exists():
        continue
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    ii = [idx[k] for k in KEEP if k in idx]
    with open(P / f"r02_ncu_full_{wl}.csv", "w") as fo:
        fo.write(f"# ncu --set full --clock-control none --import-source on, {title}\n")
        w = csv.writer(fo)
        w.writerow([hdr[i] for i in ii])
        w.writerow([units[i] for i in ii])
        for r in data:
            w.writerow([r[i] for i in ii])
    tot = []
    for r in data:
        n = r[idx["Kernel Name"]]
        b = sum(float(r[idx[m]].replace(",", "")) * UNIT[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        print(wl, n[:34].ljust(34), r[idx["gpu__time_duration.sum"]][:7], units[idx["gpu__time_duration.sum"]], "dram %4.0f MB" % (b / 1e6), "issue",
              r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]][:4], "lanes", r[idx["smsp__thread_inst_executed_per_inst_executed.ratio"]][:5])
        if any(p in n for p in pats):
            tot.append(b)
    if tot:
        json.dump({"workload": wl, "kernel": pats[0], "dram_bytes_per_launch": sum(tot) / len(tot), "launches_captured": len(tot), "per_launch_bytes": tot,
                   "wave_paths": 8 * 1920 * 1080,   # the capture renders one wave of 8 spp at 1080p; bench.py scales to its own wave size
                   "source": f"ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, {title}, round 2"},
                  open(P / f"{stem}_{wl}.json", "w"), indent=1)

# ---- phase budgets ----
rep = G / "r2f_prof_c3.ncu-rep"
if rep.exists():
    md = ["# k_bounce_small — per-phase instruction budget (round 2, final build)\n",
          "Source: `ncu --set full --clock-control none --import-source on` on `scripts/profile_step.py c3 8` (Cornell GI depth 3, 1080p, one wave of 8 spp = 16.6 M paths),",
          "joined per SASS instruction with `nvdisasm -gi` line info by `scripts/ncu_phase_budget.py` (phase = source line ranges in `profiles/phases/k_bounce_small.json`).",
          "`warp instr` = `Instructions Executed` summed over the phase's SASS instructions; `per unit` = per warp-tile of 32 queue entries.\n"]
    for i, (b, entries) in enumerate(((0, None), (1, None), (2, None)), start=1):
        md.append(f"## bounce {b}\n")
        md.append(budget(rep, "k_bounce_small", i, "4fast14k_bounce_smallILb1", "k_bounce_small.json"))
    (P / "r02_budget_bounce_small.md").write_text("\n".join(md))
rep = G / "r2f_prof_c3_primary.ncu-rep"
if rep.exists():
    md = ["# k_primary — per-phase instruction budget (round 2, final build: screen-space candidate masks)\n",
          "Source: `ncu --set full` on `scripts/profile_step.py c3 8`, one wave of 8 spp = 16.6 M paths (36 % of them outside the scissor).\n",
          budget(rep, "k_primary", 1, "4fast9k_primaryILb0ELb0", "k_primary.json")]
    (P / "r02_budget_primary.md").write_text("\n".join(md))
rep = G / "r2f_prof_c4.ncu-rep"
if rep.exists():
    md = ["# k_trace8 — per-phase instruction budget on the 999,698-triangle scene (round 2, final build)\n",
          "Source: `ncu --set full` on `scripts/profile_step.py c4 8` (one wave of 8 spp = 16.6 M paths, device-built PLOC tree), bounce 0 / 1 closest hit and the bounce-1 any-hit launch.\n"]
    for title, rx, i, sym in (("closest hit, bounce 0 (coherent primary rays)", r"k_trace8<\(bool\)0", 1, "k_trace8ILb0ELb0ELi8"),
                              ("closest hit, bounce 1 (incoherent)", r"k_trace8<\(bool\)0", 2, "k_trace8ILb0ELb0ELi8"),
                              ("any hit, bounce 1 (shadow rays)", r"k_trace8<\(bool\)1", 2, "k_trace8ILb1ELb0ELi8")):
        md.append(f"## {title}\n")
        md.append(budget(rep, rx, i, sym, "k_trace8.json"))
    (P / "r02_budget_trace8.md").write_text("\n".join(md))
rep = G / "r2f_prof_c5.ncu-rep"
if rep.exists():
    md = ["# k_volume_paths — per-phase instruction budget on workload c5 (round 2)\n",
          "Source: `ncu --set full` on `scripts/profile_step.py c5 8` (one wave of 8 spp, 3-D texture lookups).\n", budget(rep, "k_volume_paths", 1, "4fast14k_volume_pathsILb0ELi4", "k_volume_paths.json")]
    (P / "r02_budget_volume_paths.md").write_text("\n".join(md))
print("written:", sorted(p.name for p in P.glob("r02_*")))

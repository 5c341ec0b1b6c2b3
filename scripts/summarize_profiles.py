"""Turns the artefacts a gpurun measurement pass left in gpurun_out/ into the tracked summaries under profiles/.
usage: python scripts/summarize_profiles.py <suffix in gpurun_out, e.g. v4> [round tag, default r01]

Expected inputs (all optional):
  gpurun_out/bench_<wl>_<sfx>.json         bench.py line of workload c3 / c4 / c5
  gpurun_out/bench_c3_ref_<sfx>.json       bench.py --impl reference line
  gpurun_out/launches_c3_<sfx>.csv         ncu --metrics gpu__time_duration.sum launch list of a short bench.py run
  gpurun_out/prof_<wl>_<sfx>.ncu-rep       ncu --set full capture of scripts/profile_step.py <wl> 8
"""
import csv, json, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"
sfx = sys.argv[1]
tag = sys.argv[2] if len(sys.argv) > 2 else "r01"


def last_json(path):
    return [l for l in Path(path).read_text().splitlines() if l.startswith("{")][-1]


for wl in ("c3", "c4", "c5"):
    f = G / f"bench_{wl}_{sfx}.json"
    if f.exists():
        line = last_json(f); (P / f"{tag}_bench_{wl}.json").write_text(line + "\n")
        d = json.loads(line); r = d["roofline"]
        print(wl, "value %.0f Ms/s" % d["value"], "mrays %.0f" % d["mrays_per_s"], "e2e %.0f" % (d["e2e"]["value"] if d.get("e2e") else -1),
              "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None, "| %s: %.0f GB/s frac %.3f traffic %s alg %.0f MB" %
              (r["kernel"][:24], r["achieved"], r["frac"], r["traffic"], r["algorithmic_bytes_per_launch"] / 1e6), {k: round(v, 3) for k, v in r["share_of_step"].items()})
f = G / f"bench_c3_ref_{sfx}.json"
if f.exists():
    (P / f"{tag}_bench_c3_reference_arm.json").write_text(last_json(f) + "\n"); print("reference arm", json.loads(last_json(f))["value"], "Msamples/s")

f = G / f"launches_c3_{sfx}.csv"
if f.exists():
    rows = list(csv.reader(open(f))); start = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[start]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[start + 1:]:
        if len(r) <= vi: continue
        v = float(r[vi].replace(",", "")); v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["# ncu launch list summary — bench.py --steps 2 --warmup 3 --spp 16 --no-cpu-baseline --no-e2e (workload c3), build " + sfx,
           "# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised: compare SHARES",
           f"# launches captured: {sum(v[0] for v in agg.values())}, total {tot / 1e3:.2f} ms", "kernel,launches,total_us,share"]
    out += [f"{k},{v[0]},{v[1]:.1f},{v[1] / tot:.4f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    (P / f"{tag}_launches_c3_final_summary.csv").write_text("\n".join(out) + "\n"); (P / f"{tag}_launches_c3_final.csv").write_bytes(f.read_bytes())
    print("\n".join(out[3:10]))

KEEP = ["ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
# workload -> (dominant kernel patterns, traffic file stem, title)
WL = {"c3": (("k_bounce_small",), "bounce_traffic", "k_primary, then k_bounce_small at bounce 0, 1, 2 (shade + shadow rays + next closest hit fused)"),
      "c4": (("k_trace<0", "k_trace<(bool)0"), "extend_traffic", "k_trace closest (<0,0>) / any-hit (<1,0>) and k_shade_surface at bounce 0, 1, 2"),
      "c5": (("k_volume_paths",), "volume_traffic", "k_primary, then k_volume_paths (every volume path to completion)")}
for wl, (pats, stem, title) in WL.items():
    rep = G / f"prof_{wl}_{sfx}.ncu-rep"
    if not rep.exists(): continue
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}; ii = [idx[k] for k in KEEP if k in idx]
    with open(P / f"{tag}_ncu_full_{wl}_final.csv", "w") as fo:
        fo.write(f"# ncu --set full --clock-control none, scripts/profile_step.py {wl} 8, build {sfx}: {title}\n")
        w = csv.writer(fo); w.writerow([hdr[i] for i in ii]); w.writerow([units[i] for i in ii]); [w.writerow([r[i] for i in ii]) for r in data]
    tot = []
    for r in data:
        n = r[idx["Kernel Name"]]
        b = sum(float(r[idx[m]].replace(",", "")) * UNIT[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        print(wl, n[:30].ljust(30), r[idx["gpu__time_duration.sum"]][:7], units[idx["gpu__time_duration.sum"]], "dram %4.0f MB" % (b / 1e6), "issue",
              r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]][:4], "thr/inst", r[idx["smsp__thread_inst_executed_per_inst_executed.ratio"]][:5])
        if any(p in n for p in pats): tot.append(b)
    if tot:
        json.dump({"workload": wl, "kernel": pats[0], "dram_bytes_per_launch": sum(tot) / len(tot), "launches_captured": len(tot), "per_launch_bytes": tot,
                   "source": f"ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, scripts/profile_step.py {wl} 8, build {sfx}"},
                  open(P / f"{stem}_{wl}.json", "w"), indent=1)

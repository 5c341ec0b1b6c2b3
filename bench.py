#!/usr/bin/env python
"""bench.py — headline benchmark of the GPU path tracer (BASELINE.json: Msamples/s & Mrays/s, GIIntegrator 1080p).

    python bench.py --gpus N --steps K --warmup W [--workload c3|c4|c5] [--impl ours|reference]

One "step" = one full render of the workload (every pixel x every sample) through the wavefront kernels.

  workload c3 (default): Cornell box, GIIntegrator(maxDepth 3), 1920x1080, 1024 spp (BASELINE configs[2]); the spp are
           SPLIT across the N ranks (total work fixed -> "scaling": "strong"), scene replicated, per-rank SUM buffers
           reduced to rank 0 with one NCCL reduce, then divided by spp.
  workload c4: Cornell walls + ~1M-triangle displaced sphere, GI depth 3, 1080p, 64 spp (BVH-traversal-bound).
  workload c5: VolumePathTracing through a procedural 256^3 density grid, 1080p, 64 spp.

JSON line keys follow the driver contract; `value` = whole-job Msamples/s with everything resident in HBM (CUDA events,
barrier + synchronize on both sides, max over ranks); `e2e` = the same through the host-buffer C ABI call (scene
re-upload H2D + render + NCCL reduce + image D2H inside the timed region, host wall clock, max over ranks).
`--impl reference` times the reference's own CPU implementation (oracle/_ref compiled from /root/reference when
present, else the oracle port) on all host threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (scene builder key, integrator, max_depth, width, height, spp, cpu sample spp, cpu pixel stride)
    "c3": dict(scene="cornell", integrator="gi", max_depth=3, width=1920, height=1080, spp=1024, cpu_spp=4, cpu_stride=1,
               desc="Cornell box (34 tris + quad light), GIIntegrator depth 3, 1920x1080, 1024 spp"),
    "c4": dict(scene="mesh1m", integrator="gi", max_depth=3, width=1920, height=1080, spp=64, cpu_spp=1, cpu_stride=30,
               desc="Cornell walls + 999698-triangle displaced sphere, GIIntegrator depth 3, 1920x1080, 64 spp"),
    "c5": dict(scene="volume", integrator="volume", max_depth=16, width=1920, height=1080, spp=64, cpu_spp=1, cpu_stride=2,
               desc="VolumePathTracing depth 16 through a procedural 256^3 density grid + quad light, 1920x1080, 64 spp"),
}


def build_scene(kind):
    from xraytracer_b200 import scenes
    if kind == "cornell":
        return scenes.cornell_box("quad")
    if kind == "mesh1m":
        return scenes.cornell_mesh_scene(707, 707)
    if kind == "volume":
        return scenes.volume_scene(n=256, abs_color=(0.01, 0.01, 0.01), scat_color=(0.05, 0.05, 0.05))
    raise ValueError(kind)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [(ts, line) for ts, line in self.rows if t0 <= ts <= t1 + 0.2]
        if not inside and self.rows:   # timed region shorter than the sampling period: the sample closest to it
            inside = [min(self.rows, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))]
        for ts, line in inside:
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except Exception:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_baseline(desc, cam, wl, integ_id):
    """The reference's CPU renderer (compiled reference if available, else the oracle port) on ALL host threads over a
    bounded sample of the workload. Returns (Msamples/s, dict)."""
    from xraytracer_b200 import api, capi
    kind = "reference" if capi.have_reference() else "port"
    cpu = api.ReferenceScene(desc) if kind == "reference" else api.OracleScene(desc)
    # all host cores this process may use — NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    W, H = wl["width"], wl["height"]
    stride = wl["cpu_stride"]
    spp = wl["cpu_spp"]
    nx, ny = (W + stride - 1) // stride, (H + stride - 1) // stride
    _, sec, _ = cpu.render(cam, W, H, spp, integ_id, wl["max_depth"], nthreads=cores, pixel_stride=stride)
    samples = nx * ny * spp
    sample = f"{W}x{H} frame, every {stride}th pixel in x and y ({nx}x{ny} pixels), {spp} spp of {wl['spp']}" if stride > 1 else \
        f"{W}x{H}, {spp} spp of {wl['spp']}"
    return samples / sec / 1e6, {"value": samples / sec / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample,
                                 "seconds": sec}


def run_reference(args, wl, rank, world):
    """--impl reference: rank 0 alone times the CPU implementation; other ranks exit 0."""
    if rank != 0:
        return
    from xraytracer_b200 import capi, scenes
    host = build_scene(wl["scene"])
    desc = host.flatten()
    cam = scenes.make_camera(wl["width"], wl["height"])
    integ_id = capi.INTEGRATOR_NAMES.index(wl["integrator"])
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        v, info = cpu_baseline(desc, cam, wl, integ_id)
        if i >= args.warmup:
            vals.append(info["seconds"])
        last = info
    W, H = wl["width"], wl["height"]
    stride = wl["cpu_stride"]
    samples = ((W + stride - 1) // stride) * ((H + stride - 1) // stride) * wl["cpu_spp"]
    total_s = sum(vals)
    value = samples * len(vals) / total_s / 1e6
    last = dict(last)
    last["value"] = value
    last.pop("seconds", None)
    line = {"impl": "reference", "metric": "Msamples/s, " + wl["desc"], "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / len(vals), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "step": "bounded CPU sample: " + last["sample"]},
            "cpu_baseline": last, "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """Exactly ONE JSON line may reach stdout: libraries (NCCL prints its version banner there) are diverted to stderr by
    pointing fd 1 at fd 2 for the whole run; emit() writes the result line to the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override the workload's spp")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.spp:
        wl["spp"] = args.spp
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from xraytracer_b200 import api, capi, scenes
    from xraytracer_b200 import dist as xdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the render path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)

    W, H, spp_total = wl["width"], wl["height"], wl["spp"]
    integ_id = capi.INTEGRATOR_NAMES.index(wl["integrator"])
    # spp split: rank r renders sample indices [lo, hi)
    lo, hi = xdist.sample_range(spp_total, rank, world)
    my_spp = hi - lo

    host = build_scene(wl["scene"])
    desc = host.flatten()
    cam = scenes.make_camera(W, H)
    scene = api.GpuScene(desc, local_rank)
    info = scene.info()
    out = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    pinned = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()

    def step(flags, want_stats):
        """One render of this rank's spp share into `out` (per-pixel SUM), then the NCCL reduce and the 1/spp scale."""
        st = None
        if my_spp > 0:
            st = scene.render_device(cam, W, H, my_spp, integ_id, wl["max_depth"], out.data_ptr(), stream.cuda_stream,
                                     flags=flags | capi.FLAG_SUM_ONLY, seed=1234, sample_offset=lo, spp_total=spp_total,
                                     want_stats=want_stats)
        else:
            out.zero_()
        xdist.reduce_image(out, spp_total, dst=0)  # one NCCL reduce(sum), then image /= n_samples on rank 0
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- instrumented pass (untimed): node / triangle counters for the algorithmic-bytes figure ----
    cst = scene.render_device(cam, W, H, max(1, min(my_spp, 4)), integ_id, wl["max_depth"], out.data_ptr(), stream.cuda_stream,
                              flags=capi.FLAG_COUNTERS | capi.FLAG_SUM_ONLY, seed=1234, sample_offset=lo, spp_total=spp_total)
    torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step(0, False)
    barrier()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    stats = []
    for _ in range(args.steps):
        stats.append(step(capi.FLAG_STAGE_TIMES, True))
    e1.record(stream)
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clk = clocks.stop(t0, t1) if rank == 0 else None

    # aggregate ray counts over ranks
    agg = np.zeros(12, dtype=np.float64)
    for st in stats:
        if st:
            agg += np.array([st["closest_rays"], st["shadow_rays"], st["kernel_launches"], st["extend_ms"], st["shade_ms"],
                             st["connect_ms"], st["extend_launches"], st["tracking_steps"], st["primary_hits"],
                             st["bounce_entries"], st["bounce_launches"], st["shade_launches"]], dtype=np.float64)
    if world > 1:
        t = torch.tensor(agg, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        agg_all = t.cpu().numpy()
    else:
        agg_all = agg

    # ---- e2e: the host-facing call — scene H2D + render + reduce + image D2H, host wall clock ----
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            scene.upload()  # flattened scene + BVH from pinned host memory -> HBM
            step(0, False)
            if rank == 0:
                pinned.copy_(out, non_blocking=True)
            torch.cuda.synchronize(dev)
        e2e_step()
        barrier()
        tw0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        tw = time.perf_counter() - tw0
        if world > 1:
            t = torch.tensor([tw], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tw = float(t.item())
        cam_bytes = 18 * 4 + 10 * 4
        e2e = {"value": W * H * spp_total * args.steps / tw / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(info["upload_bytes"]) + cam_bytes, "d2h_bytes_per_step": W * H * 3 * 4,
               "ms_per_step": 1e3 * tw / args.steps,
               "call": "xrtg_scene_upload + xrtg_render_device + NCCL reduce + D2H to pinned host"}

    if rank == 0:
        samples = W * H * spp_total * args.steps
        value = samples / (ms * 1e-3) / 1e6
        rays = agg_all[0] + agg_all[1]
        peak, peak_src = measured_peak_hbm()
        n_cl = max(cst["closest_rays"], 1)
        nodes_per_ray = cst["nodes_visited"] / n_cl
        tris_per_ray = cst["tris_tested"] / n_cl
        ext_ms, ext_launches, closest_r0 = agg[3], max(agg[6], 1), agg[0]
        bvh_bytes = 64.0 * info["n_bvh_nodes"] + 48.0 * info["n_triangles"]
        paths_r0 = float(W) * H * my_spp * args.steps
        traffic = None
        # measured DRAM bytes per launch of the dominant kernel (one ncu --set full capture, summarised by scripts/summarize_profiles.py)
        is_volume = wl["integrator"].startswith("volume")
        tp = ROOT / "profiles" / (f"bounce_traffic_{args.workload}.json" if agg[9] > 0 else
                                  (f"volume_traffic_{args.workload}.json" if is_volume else f"extend_traffic_{args.workload}.json"))
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        share = {"extend": agg[3] / ms, "shade": agg[4] / ms, "connect": agg[5] / ms}
        if agg[9] > 0:
            # ---- small scene: the dominant kernel is k_bounce_small (shade + shadow rays + next closest hit + next RR fused) ----
            # achieved = COMPULSORY HBM bytes of its launches / their CUDA-event time, rank 0:
            #   per queue entry  : 64 B in (48 B ray/throughput/path word + 16 B hit record) + 16 B radiance read + 16 B radiance write
            #   per survivor     : 64 B out (ray + hit record of the next bounce); survivors = entries of bounce >= 1 = entries - primary hits
            # The triangles are read from shared memory (plane-paired block staged once per CTA), reported as `fetch`.
            entries, launches_b, b_ms = agg[9], max(agg[10], 1), agg[4]
            hbm_bytes = entries * 96.0 + (entries - agg[8]) * 64.0
            achieved = hbm_bytes / (b_ms * 1e-3) / 1e9 if b_ms > 0 else 0.0
            traced = (closest_r0 - paths_r0)      # closest-hit rays traced inside the bounce kernels
            smem_bytes = 80.0 * (traced * info["small_records_all"] + agg[1] * info["small_records_occ"]) if info["small_records_all"] else \
                64.0 * info["n_triangles"] * (traced + agg[1])
            roofline = {"bound": "hbm", "kernel": "k_bounce_small (per bounce: shade + NEE shadow rays + next closest hit + next Russian roulette, fused)",
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": hbm_bytes / launches_b,
                        "bytes_per_entry_hbm": hbm_bytes / max(entries, 1.0), "entries_per_launch": entries / launches_b,
                        "fetch": {"achieved": smem_bytes / (b_ms * 1e-3) / 1e9 if b_ms > 0 else 0.0, "unit": "GB/s",
                                  "level": "shared memory (every ray reads the whole plane-paired triangle block, 80 B per record; upper bound: "
                                           "shadow rays skip planes no lane of the warp can reach)",
                                  "records_closest": info["small_records_all"], "records_occluders": info["small_records_occ"]},
                        "launches": int(launches_b), "avg_launch_ms": b_ms / launches_b, "share_of_step": share,
                        "note": "latency / dependency bound (ncu: ~57 % issue-active at 31 of 32 lanes, 4.6 warps per scheduler), not memory bound: a "
                                "36-triangle scene cannot saturate HBM; the fraction says how far the queue traffic is from the HBM roof"}
        elif wl["integrator"].startswith("volume") and agg[7] > 0 and agg[4] > agg[3]:
            # ---- volume workloads: the dominant kernel is the path kernel k_volume_paths (delta tracking, run to completion) ----
            # achieved = algorithmic bytes / CUDA-event time of its launches, rank 0: SURVEY §8(d)'s 8 voxels x 4 B = 32 B per tracking
            # step + per path that enters it 64 B in (ray + hit record) and a 16 B radiance read-modify-write. The voxel gathers of
            # a 64 MiB grid are served by L2 (ncu: ~80 % L2 hit rate), so DRAM traffic is far below this figure.
            v_ms, v_launches = agg[4], max(agg[11], 1)
            hbm_bytes = agg[7] * 32.0 + agg[8] * 96.0
            achieved = hbm_bytes / (v_ms * 1e-3) / 1e9 if v_ms > 0 else 0.0
            roofline = {"bound": "hbm", "kernel": "k_volume_paths (volume paths run to completion: lockstep delta-tracking walk + inline closest hits)",
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": hbm_bytes / v_launches, "tracking_steps_per_launch": agg[7] / v_launches,
                        "tracking_steps_per_s": agg[7] / (v_ms * 1e-3) if v_ms > 0 else 0.0, "paths_per_launch": agg[8] / v_launches,
                        "launches": int(v_launches), "avg_launch_ms": v_ms / v_launches, "share_of_step": share,
                        "note": "instruction-issue bound on the tracking step (3 draws, log, 3 exp, 8 voxel gathers + trilinear weights per "
                                "step); the gathers hit L2, so the HBM fraction only says how far the walk is from the memory roof"}
        else:
            # ---- deep BVH: roofline of the dominant stage, closest-hit traversal, rank 0's launches ----
            # (k_primary = ray generation fused with the bounce-0 hit on shallow BVHs, k_extend_simple / k_trace<closest> after)
            # achieved  = COMPULSORY HBM bytes of those launches / their CUDA-event time:
            #               fused primary launch : 16 B radiance init per path + 64 B (48 B ray + 16 B hit) per primary HIT
            #               every other launch   : 32 B ray read + 16 B hit write per ray
            #               + the BVH and triangle arrays once per launch.
            #             BVH nodes / triangles re-fetched per ray are served by L1/L2, not HBM; they are reported separately
            #             as `fetch` = rays x (64 B x nodes visited + 48 B x triangles tested) / time  (SURVEY §8(d)'s B_ray).
            if agg[8] > 0:   # fused primary kernel in use
                hbm_bytes = paths_r0 * 16.0 + agg[8] * 64.0 + (closest_r0 - paths_r0) * 48.0 + ext_launches * bvh_bytes
            else:
                hbm_bytes = closest_r0 * 48.0 + ext_launches * bvh_bytes
            fetch_bytes = closest_r0 * (64.0 * nodes_per_ray + 48.0 * tris_per_ray)
            achieved = hbm_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
            roofline = {"bound": "hbm", "kernel": "closest-hit stage (k_primary fused raygen+bounce 0, k_extend_simple / k_trace after)", "achieved": achieved, "peak": peak,
                        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": hbm_bytes / ext_launches,
                        "bytes_per_ray_hbm": hbm_bytes / max(closest_r0, 1.0), "bvh_bytes": bvh_bytes,
                        "fetch": {"achieved": fetch_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0, "unit": "GB/s",
                                  "level": "L1/L2 (BVH node + triangle fetches, 64 B and 48 B records)",
                                  "bytes_per_ray": 64.0 * nodes_per_ray + 48.0 * tris_per_ray,
                                  "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray},
                        "launches": int(ext_launches), "avg_launch_ms": ext_ms / ext_launches, "share_of_step": share,
                        "note": "node/triangle counters from an instrumented run (XRTG_FLAG_COUNTERS) of the same kernels on the same scene"}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            _, cpu = cpu_baseline(desc, cam, wl, integ_id)
            cpu.pop("seconds", None)
        line = {
            "metric": "Msamples/s, " + wl["desc"], "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "spp_per_gpu": my_spp, "integrator": wl["integrator"],
                       "max_depth": wl["max_depth"], "rng": "Philox4x32-7 counter RNG keyed (seed,pixel)/(sample,block)",
                       "parallelism": f"spp split over {world} GPU(s), scene replicated, one NCCL reduce(sum) of the {W}x{H}x3 fp32 buffer",
                       "l2_policy": "per-wave working set (ray + hit queues and the per-path radiance of 8.3M paths, 0.5-1.5 GB) exceeds the 126 MB L2; no flush needed",
                       "triangles": info["n_triangles"], "bvh_nodes": info["n_bvh_nodes"]},
            "mrays_per_s": rays / (ms * 1e-3) / 1e6, "rays_per_sample": rays / samples,
            "gpu_launches": int(agg_all[2]),
            "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "scene_build_ms": info["build_ms"], "scene_upload_ms": info["upload_ms"],
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

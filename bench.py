#!/usr/bin/env python
"""bench.py — headline benchmark of the GPU path tracer (BASELINE.json: Msamples/s & Mrays/s, GIIntegrator 1080p).

    python bench.py --gpus N --steps K --warmup W [--workload c3|c4|c5] [--side c4,c5,c1,c2,exact] [--impl ours|reference]

One "step" = one full render of the workload (every pixel x every sample) through the wavefront kernels.

  workload c3 (headline): Cornell box, GIIntegrator(maxDepth 3), 1920x1080, 1024 spp (BASELINE configs[2]); the spp are SPLIT
           across the N ranks (total work fixed -> "scaling": "strong"), scene replicated, the per-rank SUM buffers meet in
           one fused peer-memory reduce + `image /= spp` kernel per rank (CUDA IPC over NVLink; NCCL reduce if IPC is refused).
  workload c4: Cornell box + 999,698-triangle displaced sphere, GI depth 3, 1080p, 64 spp (BVH-traversal-bound; configs[3]).
  workload c5: VolumePathTracing through a procedural 256^3 density grid, 1080p, 256 spp (configs[4]).

The JSON line carries the headline workload at top level (driver contract) and, under "workloads", the same measurements for
the side workloads c4 and c5 taken in the same run on the same ranks, plus "exact_mode_c3": the throughput of the exact
instantiation (per-pixel mt19937 stream, no FMA) that the drop-in NormalRenderer / ParallelRenderer names map to.

  value     whole-job Msamples/s with everything resident in HBM: flags = 0, no statistics, no host synchronisation inside the
            timed region; CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
  e2e       the same metric through the plug-in call a C++ user makes: xrtg_scene_upload + xrtg_render(host buffer) — scene
            H2D from pinned memory, render, `image /= spp`, image D2H, host wall clock. At N > 1 rank 0 drives ALL N GPUs
            through one xrtg_scene_create_multi handle (single process, no Python / NCCL on the path); the other ranks wait.
  roofline  dominant kernel of each workload, stage times from a separate UNTIMED instrumented pass (XRTG_FLAG_STAGE_TIMES).
  mrays_per_s counts reference-equivalent rays (Scene::intersect + Scene::occluded calls the reference would make);
  mrays_traced_per_s counts the rays the kernels really traced (the primary scissor resolves some samples without a ray).

`--impl reference` times the reference's own CPU implementation (oracle/_ref compiled from /root/reference when present, else
the oracle port) on all host threads, on a bounded sample of the same workload; its process loads oracle/ libraries only.
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
    "c1": dict(scene="cornell", integrator="normal", max_depth=1, width=512, height=512, spp=16, cpu_spp=16, cpu_stride=1,
               desc="Cornell box, NormalIntegrator visibility, 512x512, 16 spp"),
    "c2": dict(scene="cornell_tri", integrator="direct", max_depth=1, width=1920, height=1080, spp=64, cpu_spp=2, cpu_stride=1,
               desc="Cornell box, DirectIntegrator with a TriangleLight (half of the quad), 1920x1080, 64 spp"),
    "c3": dict(scene="cornell", integrator="gi", max_depth=3, width=1920, height=1080, spp=1024, cpu_spp=4, cpu_stride=1,
               desc="Cornell box (34 tris + quad light), GIIntegrator depth 3, 1920x1080, 1024 spp"),
    "c4": dict(scene="mesh1m", integrator="gi", max_depth=3, width=1920, height=1080, spp=64, cpu_spp=1, cpu_stride=30,
               desc="Cornell box + 999698-triangle displaced sphere, GIIntegrator depth 3, 1920x1080, 64 spp"),
    "c5": dict(scene="volume", integrator="volume", max_depth=16, width=1920, height=1080, spp=256, cpu_spp=8, cpu_stride=2,
               desc="VolumePathTracing depth 16 through a procedural 256^3 density grid + quad light, 1920x1080, 256 spp"),
}
INTEGRATORS = ["normal", "furnace", "direct", "indirect", "gi", "whitted", "volume", "volume_nee"]
CORNELL_C2W = [-1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, -1.0, 0, 278, 274.4, -750.0, 1]


def build_scene(kind):
    """Host C++ scene (libxrthost.so) — the product path."""
    from xraytracer_b200 import scenes
    if kind == "cornell":
        return scenes.cornell_box("quad")
    if kind == "cornell_tri":
        return scenes.cornell_box("triangle")
    if kind == "mesh1m":
        return scenes.cornell_mesh_scene(707, 707)
    if kind == "volume":
        return scenes.volume_scene(n=256, abs_color=(0.01, 0.01, 0.01), scat_color=(0.05, 0.05, 0.05))
    raise ValueError(kind)


def build_flat_scene(kind):
    """The same scenes described in pure Python (xraytracer_b200/flatdesc.py): what the CPU reference arm consumes, so that
    its process maps no product library."""
    from xraytracer_b200 import flatdesc
    if kind == "cornell":
        return flatdesc.cornell_box()
    if kind == "cornell_tri":
        return flatdesc.cornell_box("triangle")
    if kind == "mesh1m":
        return flatdesc.cornell_mesh_scene(707, 707)
    if kind == "volume":
        return flatdesc.volume_scene(n=256)
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


def cpu_baseline(wl):
    """The reference's CPU renderer (compiled reference if available, else the oracle port) on ALL host threads over a
    bounded sample of the workload, on a scene description built WITHOUT the product libraries. Returns the cpu_baseline dict
    (+ "seconds")."""
    from xraytracer_b200 import api, capi, flatdesc
    flat = build_flat_scene(wl["scene"])
    desc = flat.desc()
    kind = "reference" if capi.have_reference() else "port"
    cpu = api.ReferenceScene(desc) if kind == "reference" else api.OracleScene(desc)
    # all host cores this process may use — NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    W, H = wl["width"], wl["height"]
    cam = flatdesc.make_camera(W, H, CORNELL_C2W, 60.0)
    stride, spp = wl["cpu_stride"], wl["cpu_spp"]
    nx, ny = (W + stride - 1) // stride, (H + stride - 1) // stride
    integ_id = INTEGRATORS.index(wl["integrator"])
    _, sec, _ = cpu.render(cam, W, H, spp, integ_id, wl["max_depth"], nthreads=cores, pixel_stride=stride)
    samples = nx * ny * spp
    sample = f"{W}x{H} frame, every {stride}th pixel in x and y ({nx}x{ny} pixels), {spp} spp of {wl['spp']}" if stride > 1 else \
        f"{W}x{H}, {spp} spp of {wl['spp']}"
    return {"value": samples / sec / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample, "seconds": sec}


def run_reference(args, wl, rank, world):
    """--impl reference: rank 0 alone times the CPU implementation; other ranks exit 0."""
    if rank != 0:
        return
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        info = cpu_baseline(wl)
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
    loaded = sorted({Path(l.split()[-1]).name for l in open("/proc/self/maps") if "libxrt" in l})
    line = {"impl": "reference", "metric": "Msamples/s, " + wl["desc"], "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / len(vals), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "step": "bounded CPU sample: " + last["sample"],
                       "libraries_mapped": loaded},
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


class Rig:
    """Per-process CUDA / torch.distributed state shared by every workload of the run."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the render path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        if args.gpus != self.world and self.rank == 0 and self.world > 1:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE {self.world}", file=sys.stderr)
        # a CPU-side (gloo) group for waits that must not occupy the GPUs with a spinning NCCL kernel
        self.cpu_group = dist.new_group(backend="gloo") if self.world > 1 else None
        self.stream = torch.cuda.current_stream(self.dev)
        self.flag = torch.zeros(1, dtype=torch.int32, device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def host_barrier(self):
        """Ranks wait on the CPU (gloo): the GPUs stay idle while rank 0 drives all of them through one multi-GPU handle."""
        if self.world > 1:
            self.torch.cuda.synchronize(self.dev)
            self.dist.barrier(group=self.cpu_group)

    def stream_barrier(self):
        """Stream-ordered rendezvous of all ranks (a 4-byte NCCL all-reduce on the render stream): when it completes on a
        rank's stream, every rank's stream has reached it — no host synchronisation."""
        if self.world > 1:
            self.dist.all_reduce(self.flag)

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, arr):
        if self.world == 1:
            return arr
        t = self.torch.tensor(arr, dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()


class FusedReduce:
    """The multi-process form of libxrtgpu's fused reduce + finalize (csrc/multi.cu): every rank renders into its exportable
    SUM buffer (xrtg_exchange_buffer slot 0); CUDA IPC maps every rank's buffer and rank 0's image (slot 1) into every rank;
    after a stream-ordered rendezvous each rank runs ONE kernel over its slice — pull the slice from all ranks over NVLink,
    add in rank order, divide by the total spp, store into rank 0's image. Falls back to ONE NCCL reduce + divide when IPC
    mapping is refused (the mode is reported in the JSON line)."""

    def __init__(self, rig: Rig, scene, W, H):
        self.rig, self.scene, self.n = rig, scene, W * H * 3
        torch, dist = rig.torch, rig.dist
        nbytes = self.n * 4
        self.out_t = torch.empty((H, W, 3), dtype=torch.float32, device=rig.dev)   # world 1 / NCCL mode: render target + result
        self.mode = "single" if rig.world == 1 else "ipc"
        self.target = self.out_t.data_ptr()
        self.opened = []
        if rig.world == 1:
            return
        ok = 1
        try:
            self.partial = scene.exchange_buffer(0, nbytes)
            self.final = scene.exchange_buffer(1, nbytes) if rig.rank == 0 else 0
            mine = scene.ipc_export(self.partial) + (scene.ipc_export(self.final) if rig.rank == 0 else bytes(64))
            hs = torch.tensor(list(mine), dtype=torch.uint8, device=rig.dev)
            allh = [torch.empty_like(hs) for _ in range(rig.world)]
            dist.all_gather(allh, hs)
            allh = [bytes(h.cpu().numpy().tolist()) for h in allh]
            self.parts = []
            for r, h in enumerate(allh):
                if r == rig.rank:
                    self.parts.append(self.partial)
                else:
                    self.parts.append(scene.ipc_open(h[:64]))
                    self.opened.append(self.parts[-1])
            if rig.rank == 0:
                self.final_mapped = self.final
            else:
                self.final_mapped = scene.ipc_open(allh[0][64:])
                self.opened.append(self.final_mapped)
        except Exception as e:   # IPC refused (container / driver policy)
            print(f"[rank {rig.rank}] CUDA IPC unavailable ({e}); using the NCCL reduce", file=sys.stderr)
            ok = 0
        t = torch.tensor([ok], dtype=torch.int32, device=rig.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) == 1:
            self.target = self.partial
            per = ((self.n + rig.world - 1) // rig.world + 3) & ~3   # slice boundaries on 16-byte boundaries
            self.first = min(self.n, per * rig.rank)
            self.count = min(self.n, per * (rig.rank + 1)) - self.first
        else:
            self.mode = "nccl"

    def finish(self, spp_total):
        """After this rank's render into `target` (per-pixel SUM) has been enqueued: produce the mean image on rank 0."""
        rig = self.rig
        if self.mode == "ipc":
            rig.stream_barrier()     # every rank's SUM is complete (stream order, no host sync)
            self.scene.reduce_finalize(self.parts, self.final_mapped, self.first, self.count, float(spp_total), rig.stream.cuda_stream)
            rig.stream_barrier()     # every slice has landed in rank 0's image; the SUM buffers may be overwritten
        elif self.mode == "nccl":
            rig.dist.reduce(self.out_t, dst=0, op=rig.dist.ReduceOp.SUM)
            if rig.rank == 0:
                self.out_t.div_(float(spp_total))

    def result_ptr(self):
        return self.final if self.mode == "ipc" else self.out_t.data_ptr()

    def launches_per_step(self):
        return {"single": 0, "ipc": 1, "nccl": 2}[self.mode]

    def close(self):
        for p in self.opened:
            try:
                self.scene.ipc_close(p)
            except Exception:
                pass
        self.opened = []


STAT_KEYS = ["closest_rays", "shadow_rays", "kernel_launches", "extend_ms", "shade_ms", "connect_ms", "extend_launches", "tracking_steps",
             "primary_hits", "bounce_entries", "bounce_launches", "shade_launches", "rays_traced", "truncated_paths", "other_ms", "render_ms",
             "connect_launches", "dropped_samples", "untraced_closest", "untraced_shadow"]


def roofline_for(name, wl, info, st, cst, paths_r0, peak, peak_src, wave_paths=0):
    """Roofline object of the workload's dominant kernel. `st` = rank 0's stage times and unit counts summed over the steps of
    the instrumented pass; `cst` = one pass with the node / triangle counters on."""
    ms_total = st["render_ms"]
    share = {"extend": st["extend_ms"] / ms_total, "shade": st["shade_ms"] / ms_total, "connect": st["connect_ms"] / ms_total,
             "other": st["other_ms"] / ms_total} if ms_total > 0 else None
    traffic = None
    is_volume = wl["integrator"].startswith("volume")
    tp = ROOT / "profiles" / (f"bounce_traffic_{name}.json" if st["bounce_entries"] > 0 else
                              (f"volume_traffic_{name}.json" if is_volume else f"extend_traffic_{name}.json"))
    traffic_note = None
    if tp.exists():
        try:
            tj = json.loads(tp.read_text())
            traffic = tj.get("dram_bytes_per_launch")
            # the ncu capture renders a smaller wave than the bench does (replaying a 10 GB wave 40 times per kernel is slow): the queue
            # traffic is proportional to the paths of a wave, so the per-launch figure is scaled to this run's wave size
            if traffic is not None and tj.get("wave_paths") and wave_paths:
                traffic = traffic * wave_paths / tj["wave_paths"]
                traffic_note = f"ncu capture of a wave of {tj['wave_paths']} paths, scaled linearly to this run's {wave_paths} paths per wave"
        except Exception:
            traffic = None
    n_cl = max(cst["closest_rays"] - cst.get("untraced_closest", 0), 1)   # closest-hit rays the kernels TRACED in the counter pass
    nodes_per_ray, tris_per_ray = cst["nodes_visited"] / n_cl, cst["tris_tested"] / n_cl
    common = {"bound": "hbm", "peak": peak, "unit": "GB/s", "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "share_of_step": share,
              "timing": "CUDA events around every launch of the kernel (XRTG_FLAG_STAGE_TIMES) in an instrumented pass of the same K steps, "
                        "run right after the timed region; the timed region itself carries no instrumentation"}
    if st["bounce_entries"] > 0:
        # small scene: k_bounce_small. COMPULSORY HBM bytes: per queue entry 64 B in (48 B ray/throughput/path word + 16 B hit
        # record) + 16 B radiance read + 16 B radiance write; per survivor 64 B out (ray + hit record of the next bounce);
        # survivors = entries of bounce >= 1 = entries - primary hits. Triangles come from shared memory (`fetch`).
        entries, launches_b, b_ms = st["bounce_entries"], max(st["bounce_launches"], 1), st["shade_ms"]
        hbm_bytes = entries * 96.0 + (entries - st["primary_hits"]) * 64.0
        achieved = hbm_bytes / (b_ms * 1e-3) / 1e9 if b_ms > 0 else 0.0
        traced = st["closest_rays"] - paths_r0      # closest-hit rays traced inside the bounce kernels
        smem_bytes = 80.0 * (traced * info["small_records_all"] + st["shadow_rays"] * info["small_records_occ"]) if info["small_records_all"] else \
            64.0 * info["n_triangles"] * (traced + st["shadow_rays"])
        return dict(common, kernel="k_bounce_small (per bounce: shade + NEE shadow rays + next closest hit + next Russian roulette, fused)",
                    achieved=achieved, frac=achieved / peak, algorithmic_bytes_per_launch=hbm_bytes / launches_b,
                    bytes_per_entry_hbm=hbm_bytes / max(entries, 1.0), entries_per_launch=entries / launches_b,
                    fetch={"achieved": smem_bytes / (b_ms * 1e-3) / 1e9 if b_ms > 0 else 0.0, "unit": "GB/s",
                           "level": "shared memory (every ray reads the plane-paired triangle block, 80 B per record; upper bound: shadow rays skip "
                                    "planes no lane of the warp can reach)",
                           "records_closest": info["small_records_all"], "records_occluders": info["small_records_occ"]},
                    launches=int(launches_b), avg_launch_ms=b_ms / launches_b,
                    note="latency / dependency bound (ncu: issue-active and lanes in profiles/), not memory bound: a 36-triangle scene cannot "
                         "saturate HBM; the fraction says how far the queue traffic is from the HBM roof")
    if is_volume and st["tracking_steps"] > 0 and st["shade_ms"] > st["extend_ms"]:
        # volume workloads: k_volume_paths. SURVEY §8(d): 8 voxels x 4 B = 32 B per tracking step + per path that enters it 64 B in
        # (ray + hit record) and a 16 B radiance read-modify-write. The gathers of a 64 MiB grid are served by L2.
        v_ms, v_launches = st["shade_ms"], max(st["shade_launches"], 1)
        hbm_bytes = st["tracking_steps"] * 32.0 + st["primary_hits"] * 96.0
        achieved = hbm_bytes / (v_ms * 1e-3) / 1e9 if v_ms > 0 else 0.0
        return dict(common, kernel="k_volume_paths (volume paths run to completion: lockstep delta-tracking walk + inline closest hits)",
                    achieved=achieved, frac=achieved / peak, algorithmic_bytes_per_launch=hbm_bytes / v_launches,
                    tracking_steps_per_launch=st["tracking_steps"] / v_launches, tracking_steps_per_s=st["tracking_steps"] / (v_ms * 1e-3) if v_ms > 0 else 0.0,
                    paths_per_launch=st["primary_hits"] / v_launches, launches=int(v_launches), avg_launch_ms=v_ms / v_launches,
                    truncated_paths=int(st["truncated_paths"]),
                    note="instruction-issue bound on the tracking step (3 draws, log, 8 voxel gathers + trilinear weights per step); the gathers "
                         "hit L2, so the HBM fraction only says how far the walk is from the memory roof")
    # deep BVH: closest-hit traversal stage (k_trace on the wide tree). COMPULSORY HBM bytes: 32 B ray read + 16 B hit write per
    # ray + the tree and triangle arrays once per launch; node / triangle re-fetches are served by L1/L2 and reported as `fetch`
    # = rays x (node bytes x nodes visited + triangle bytes x triangles tested) / time (SURVEY §8(d)'s B_ray).
    # (rays TRACED by the stage: the reference-equivalent count minus the extension rays of Russian-roulette losers / scissored primaries)
    ext_ms, ext_launches, closest = st["extend_ms"], max(st["extend_launches"], 1), st["closest_rays"] - st.get("untraced_closest", 0)
    node_bytes = {8: 80.0, 4: 128.0}.get(info["wide_arity"], 64.0)
    n_nodes = info["n_wide_nodes"] if info["wide_arity"] else info["n_bvh_nodes"]
    bvh_bytes = node_bytes * n_nodes + 64.0 * info["n_triangles"]
    if st["primary_hits"] > 0:   # fused primary kernel in use (shallow BVHs)
        hbm_bytes = paths_r0 * 16.0 + st["primary_hits"] * 64.0 + max(st["closest_rays"] - paths_r0 - st.get("untraced_closest", 0), 0) * 48.0 + ext_launches * bvh_bytes
    else:
        hbm_bytes = closest * 48.0 + ext_launches * bvh_bytes
    fetch_bytes = closest * (node_bytes * nodes_per_ray + 64.0 * tris_per_ray)
    achieved = hbm_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
    return dict(common, kernel="closest-hit stage (k_trace: resumable traversal of the wide BVH, plane-equation triangle records)",
                achieved=achieved, frac=achieved / peak, algorithmic_bytes_per_launch=hbm_bytes / ext_launches,
                bytes_per_ray_hbm=hbm_bytes / max(closest, 1.0), bvh_bytes=bvh_bytes, rays_traced_by_stage=int(closest),
                fetch={"achieved": fetch_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0, "unit": "GB/s",
                       "level": f"L1/L2 (node + triangle fetches, {int(node_bytes)} B and 64 B records)",
                       "bytes_per_ray": node_bytes * nodes_per_ray + 64.0 * tris_per_ray, "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray,
                       "wide_arity": info["wide_arity"]},
                launches=int(ext_launches), avg_launch_ms=ext_ms / ext_launches,
                note="node/triangle counters from an instrumented run (XRTG_FLAG_COUNTERS) of the same kernels on the same scene")


def run_workload(rig: Rig, name, wl, args, headline):
    """Measure one workload on this rank's GPU (all ranks call it together). Returns the result dict on rank 0, else None."""
    import numpy as np
    from xraytracer_b200 import api, capi, scenes
    from xraytracer_b200 import dist as xdist
    torch = rig.torch
    W, H, spp_total = wl["width"], wl["height"], wl["spp"]
    integ_id = capi.INTEGRATOR_NAMES.index(wl["integrator"])
    lo, hi = xdist.sample_range(spp_total, rig.rank, rig.world)   # spp split: rank r renders sample indices [lo, hi)
    my_spp = hi - lo
    host = build_scene(wl["scene"])
    desc = host.flatten()
    cam = scenes.make_camera(W, H)
    t_create = time.perf_counter()
    scene = api.GpuScene(desc, rig.local_rank)
    create_ms = 1e3 * (time.perf_counter() - t_create)   # whole xrtg_scene_create call (the first one of a process also creates the CUDA context)
    info = scene.info()
    # creation has no warm-up: the first one in a process state (right after another scene's multi-GB workspace was released, with
    # the clock sampler's nvidia-smi polling the driver) occasionally takes 10x longer than the steady state, so a device-built
    # scene is created a second time and both times are reported
    build_repeat_ms = None
    if info["bvh_builder"] == 2:
        again = api.GpuScene(desc, rig.local_rank)
        build_repeat_ms = again.info()["build_ms"]
        del again
    red = FusedReduce(rig, scene, W, H)
    stream = rig.stream
    steps, warmup = args.steps, max(args.warmup, 3)

    def render(flags, want_stats, target, spp=my_spp, sample_offset=lo, sum_only=True):
        return scene.render_device(cam, W, H, spp, integ_id, wl["max_depth"], target, stream.cuda_stream,
                                   flags=flags | (capi.FLAG_SUM_ONLY if sum_only else 0), seed=1234, sample_offset=sample_offset, spp_total=spp_total,
                                   want_stats=want_stats)

    def step(flags=0, want_stats=False):
        """One render of this rank's share, then the reduce + finalize that leaves the mean image on rank 0."""
        st = None
        if rig.world == 1:
            st = render(flags, want_stats, red.target, sum_only=False)   # single GPU: k_finalize divides, nothing to reduce
        else:
            if my_spp > 0:
                st = render(flags, want_stats, red.target)
            else:   # more ranks than samples: this rank contributes zeros
                render(flags, False, red.target, spp=1)
                raise SystemExit("bench.py: fewer samples than ranks")
            red.finish(spp_total)
        return st

    # ---- node / triangle counters (one short instrumented render) ----
    cst = render(capi.FLAG_COUNTERS, True, red.out_t.data_ptr(), spp=max(1, min(my_spp, 4)))
    torch.cuda.synchronize(rig.dev)

    for _ in range(warmup):
        step()
    rig.barrier()

    clocks = ClockSampler(rig.local_rank)
    if rig.rank == 0:
        clocks.start()
        time.sleep(0.3)
    rig.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    for _ in range(steps):
        step()              # flags = 0, no statistics: nothing inside the timed region synchronises with the host
    e1.record(stream)
    rig.barrier()
    t1 = time.time()
    ms = rig.max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop(t0, t1) if rig.rank == 0 else None

    # ---- instrumented pass (untimed): stage times + ray counters of the very same steps ----
    agg = np.zeros(len(STAT_KEYS), dtype=np.float64)
    for _ in range(steps):
        st = step(capi.FLAG_STAGE_TIMES, True)
        if st:
            agg += np.array([st[k] for k in STAT_KEYS], dtype=np.float64)
    rig.barrier()
    agg_all = rig.sum_over_ranks(agg)
    st0 = dict(zip(STAT_KEYS, agg))
    st_all = dict(zip(STAT_KEYS, agg_all))

    # ---- e2e: the plug-in call with HOST buffers — scene H2D + render on all N GPUs + reduce + finalize + image D2H ----
    e2e = None
    if not args.no_e2e:
        if rig.rank == 0:
            multi = api.GpuScene(desc, devices=list(range(rig.world))) if rig.world > 1 else scene
            pinned = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
            host_img = pinned.numpy()

            def e2e_step():
                multi.upload()   # flattened scene + BVH from pinned host memory -> HBM of every device
                multi.render(cam, W, H, spp_total, integ_id, wl["max_depth"], seed=1234, out=host_img)   # xrtg_render: host buffer in, mean image out
            e2e_step()
            tw0 = time.perf_counter()
            for _ in range(steps):
                e2e_step()
            tw = time.perf_counter() - tw0
            up = multi.info()
            cam_bytes = 18 * 4 + 10 * 4
            e2e = {"value": W * H * spp_total * steps / tw / 1e6, "unit": "Msamples/s",
                   "h2d_bytes_per_step": int(up["upload_bytes"]) + cam_bytes, "d2h_bytes_per_step": W * H * 3 * 4,
                   "ms_per_step": 1e3 * tw / steps, "devices": multi.device_count(),
                   "call": "xrtg_scene_upload + xrtg_render(host buffer)" + (f" on one xrtg_scene_create_multi handle over {rig.world} GPUs, single "
                           "process: spp split across devices, fused peer-memory reduce + finalize, no NCCL" if rig.world > 1 else "")}
            if multi is not scene:
                del multi
        rig.host_barrier()   # the other ranks wait on the CPU (a NCCL barrier would spin on their GPUs while rank 0 uses them)

    result = None
    if rig.rank == 0:
        samples = W * H * spp_total * steps
        peak, peak_src = measured_peak_hbm()
        paths_r0 = float(W) * H * my_spp * steps
        rays = st_all["closest_rays"] + st_all["shadow_rays"]
        cpu = None   # filled in by main() after ALL GPU measurements (see there)
        result = {
            "metric": "Msamples/s, " + wl["desc"], "value": samples / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "n_gpus": rig.world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name + ": " + wl["desc"], "spp_per_gpu": my_spp, "integrator": wl["integrator"],
                       "max_depth": wl["max_depth"], "rng": "Philox4x32-7 counter RNG keyed (seed,pixel)/(sample,block)",
                       "parallelism": f"spp split over {rig.world} GPU(s), scene replicated; reduce = {red.mode}" +
                                      (" (one fused peer-memory pull + `image /= spp` kernel per rank over CUDA IPC, two 4-byte NCCL all-reduces as stream-ordered rendezvous)"
                                       if red.mode == "ipc" else (" (one NCCL reduce(sum) + divide)" if red.mode == "nccl" else "")),
                       "l2_policy": "per-wave working set (ray + hit queues and the per-path radiance of up to 66M paths = 32 samples of every pixel, several GB) exceeds the 126 MB L2; no flush needed",
                       "triangles": info["n_triangles"], "bvh_nodes": info["n_bvh_nodes"], "wide_arity": info["wide_arity"], "wide_nodes": info["n_wide_nodes"]},
            "mrays_per_s": rays / (ms * 1e-3) / 1e6, "mrays_traced_per_s": st_all["rays_traced"] / (ms * 1e-3) / 1e6,
            "rays_per_sample": rays / samples, "rays_traced": int(st_all["rays_traced"] / steps), "rays_reference_equivalent": int(rays / steps),
            "gpu_launches": int(st_all["kernel_launches"]) + red.launches_per_step() * steps * rig.world,
            "clocks": clk, "roofline": roofline_for(name, wl, info, st0, cst, paths_r0, peak, peak_src, wave_paths=min(my_spp, max(1, (64 << 20) // (W * H))) * W * H), "cpu_baseline": cpu, "e2e": e2e,
            "scene_build_ms": info["build_ms"], "bvh_build_ms": info["bvh_build_ms"], "scene_upload_ms": info["upload_ms"], "scene_create_call_ms": create_ms, "scene_build_ms_repeat": build_repeat_ms,
            "bvh_builder": ["host binned SAH", "GPU LBVH", "GPU ingest + PLOC + sweep-SAH top levels + eight-child collapse (csrc/gpu_build.cu)"][info["bvh_builder"]],
            "truncated_paths": int(st_all["truncated_paths"] / steps), "dropped_samples": int(st_all["dropped_samples"] / steps),
        }
    red.close()
    del scene
    return result


def run_exact_mode(rig: Rig, args):
    """Throughput of the EXACT instantiation on the c3 scene (what the drop-in NormalRenderer / ParallelRenderer names map to:
    per-pixel mt19937 stream, no FMA contraction, one sample per pixel per wave). Rank 0 only, 16 spp."""
    if rig.rank != 0:
        return None
    from xraytracer_b200 import api, capi, scenes
    torch = rig.torch
    wl = WORKLOADS["c3"]
    W, H, spp = wl["width"], wl["height"], 16
    host = build_scene(wl["scene"])
    scene = api.GpuScene(host.flatten(), rig.local_rank)
    cam = scenes.make_camera(W, H)
    out = torch.empty((H, W, 3), dtype=torch.float32, device=rig.dev)
    run = lambda: scene.render_device(cam, W, H, spp, capi.INT_GI, wl["max_depth"], out.data_ptr(), rig.stream.cuda_stream, flags=capi.FLAG_EXACT, want_stats=False)
    run()
    torch.cuda.synchronize(rig.dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = max(2, min(args.steps, 5))
    e0.record(rig.stream)
    for _ in range(n):
        run()
    e1.record(rig.stream)
    torch.cuda.synchronize(rig.dev)
    ms = e0.elapsed_time(e1) / n
    return {"value": W * H * spp / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms, "spp": spp, "steps": n,
            "config": "Cornell GI depth 3, 1920x1080, exact instantiation (XRTG_FLAG_EXACT: per-pixel std::mt19937 stream in HBM, -fmad=false), 1 GPU"}


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS), help="headline workload (top level of the JSON line)")
    ap.add_argument("--side", default="c4,c5,c1,c2,exact", help="comma list of side measurements reported under 'workloads' / 'exact_mode_c3' ('' = none)")
    ap.add_argument("--spp", type=int, default=0, help="override the headline workload's spp")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.spp:
        wl["spp"] = args.spp
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    rig = Rig(args)
    line = run_workload(rig, args.workload, wl, args, True)
    side = [x for x in args.side.split(",") if x]
    extra = {}
    for name in side:
        if name in WORKLOADS and name != args.workload:
            r = run_workload(rig, name, dict(WORKLOADS[name]), args, False)
            if rig.rank == 0:
                extra[name] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "mrays_per_s", "mrays_traced_per_s", "rays_per_sample", "rays_traced",
                                                   "roofline", "cpu_baseline", "e2e", "config", "clocks", "gpu_launches", "scene_build_ms", "bvh_build_ms",
                                                   "scene_upload_ms", "scene_create_call_ms", "scene_build_ms_repeat", "bvh_builder", "truncated_paths", "dropped_samples")}
    exact = run_exact_mode(rig, args) if "exact" in side else None
    # The CPU baselines (the compiled reference on every host core, a few seconds each) run AFTER every GPU measurement of the
    # process: scene creations that followed one inside the same process were occasionally 10-30x slower than on a quiet process
    # (22 ms -> 50-700 ms for the device build of c4), and nothing on the GPU side should be billed for that.
    if rig.rank == 0 and not args.no_cpu_baseline and rig.world == 1:
        for name, res in [(args.workload, line)] + list(extra.items()):
            cpu = cpu_baseline(dict(WORKLOADS[name], **({"spp": args.spp} if (args.spp and name == args.workload) else {})))
            cpu.pop("seconds", None)
            res["cpu_baseline"] = cpu
    if rig.rank == 0:
        if extra:
            line["workloads"] = extra
        if exact:
            line["exact_mode_c3"] = exact
        emit(line)
    rig.host_barrier()
    if rig.world > 1:
        rig.dist.destroy_process_group()


if __name__ == "__main__":
    main()

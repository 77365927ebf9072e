#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (primary+GI Mrays/s and ms/frame at 3840x2160, % of HBM roofline) for the
voxel-rendering hot path, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one frame of the hot path over the synthetic scene: clear + visibility (K1) [+ SVO-GI shading (K3) when the
SVO exists] through libtgb200.so. N=1 workload = BASELINE configs[1] (1,024 objects, 2^21 clusters, 4K); at N>1 every
rank owns one such 1,024-object shard of an N-times larger world (weak scaling), renders the full 4K frame against its
shard and the frames are merged with ncclAllReduce(u64, min). `value` counts the rays all ranks traced per second.
`--impl reference` times the CPU path (the oracle port of the reference's shader logic; the reference itself is
Win32/Vulkan-only and cannot run here) on a bounded scanline sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 3840, 2160
METRIC = "primary+GI Mrays/s at 3840x2160"
CPU_YSTEP = 8          # cpu_baseline sample: every 8th scanline of the same frame
ALG_BYTES_PER_CLUSTER = 72   # 64 B mask + 4 B pointer + 4 B cluster->object
ALG_BYTES_PER_OBJECT = 96
ALG_BYTES_PER_PIXEL_VIS = 16  # clear store + resolved store
ALG_BYTES_PER_PIXEL_GI = 40   # vis + ptr + c2o + LUT-idx word + LUT + RGBA32F out


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def build_scene(rank, n_ranks):
    """Rank's shard: 1,024 objects (grid 32 x 32) out of a 32 x 32N lattice; global object index decides seed and angle."""
    from tg_b200 import scenes
    grid_x, grid_z = 32, 32 * n_ranks
    s = scenes.grid_scene(f"config2_x{n_ranks}", grid_x, grid_z, WIDTH, HEIGHT, k=3, first_object=rank * 1024, n_objects=1024)
    return s


def cpu_baseline_sample(scene, n_ranks, threads=None, repeats=1):
    """Oracle (scalar C port of the reference's shader logic) on every CPU_YSTEP-th scanline. Returns (Mrays/s, cores, seconds, rays)."""
    from oracle import oracle as O
    if threads:
        O.lib().tgo_set_threads(threads)
    cores = O.lib().tgo_max_threads()
    cam = O.camera_from_spec(scene.camera)
    rays = O.camera_rays(cam)
    view = O.SceneView.from_scene(scene, with_lut=False)
    best = None
    n_rows = len(range(0, HEIGHT, CPU_YSTEP))
    for _ in range(repeats):
        t0 = time.perf_counter()
        vis, _ = O.visibility(view, rays, WIDTH, HEIGHT, O.VIS_SCREEN_RECT, 0, HEIGHT, CPU_YSTEP)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    n_rays = n_rows * WIDTH  # primary rays of the sample (GI rays are added by the caller when GI is on)
    return n_rays / best / 1e6, cores, best, n_rays


def run_reference(args, rank, world):
    """The CPU arm: rank 0 only."""
    if rank != 0:
        return
    scene = build_scene(0, 1)
    from oracle import oracle as O
    cores = O.lib().tgo_max_threads()
    cam = O.camera_from_spec(scene.camera)
    rays = O.camera_rays(cam)
    view = O.SceneView.from_scene(scene, with_lut=False)
    n_rows = len(range(0, HEIGHT, CPU_YSTEP))
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.visibility(view, rays, WIDTH, HEIGHT, O.VIS_SCREEN_RECT, 0, HEIGHT, CPU_YSTEP)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    n_rays = n_rows * WIDTH
    total = sum(times)
    value = n_rays * len(times) / total / 1e6
    sample = f"every {CPU_YSTEP}th scanline of the 3840x2160 frame ({n_rows} rows, {n_rays} primary rays) per step, screen-rect pruned oracle, OpenMP"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u64", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: 1,024 objects (2^21 clusters, 1.07e9 voxels), 3840x2160, primary visibility", "sample": sample},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tg_b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import tg_b200
    from tg_b200.raytracer import comm_unique_id, from_scene

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene = build_scene(rank, world)
    rt = from_scene(scene, device=local_rank)
    n_clusters, n_objects = scene.n_clusters, len(scene.objects)
    if world > 1:
        rt.set_shard(rank, world, rank * n_clusters)
        ids = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        rt.comm_init(ids[0], rank, world)

    lib = tg_b200.lib()
    import ctypes as C
    stream = torch.cuda.ExternalStream(lib.tgb200_stream(C.byref(rt._rt)), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")  # > 126 MB L2
    host_rad = torch.empty(WIDTH * HEIGHT * 4, dtype=torch.float32).pin_memory()
    host_rad_np = host_rad.numpy().reshape(HEIGHT, WIDTH, 4)

    # static scene: the SVO is built once, like the reference does on its first frame (tgvk_raytracer.c:1187-1217)
    rt.set_gi(True, 1)
    rt.svo_update(force_full=True)
    rt.synchronize()
    svo_build_ms = rt.timings()["svo_ms"]

    def frame():
        rt.clear()
        rt.render_visibility()
        if world > 1:
            rt.merge_visibility()
        rt.render_shading()

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.fill_(1)

    # ---- warm-up ----
    for _ in range(args.warmup):
        flush_l2()
        frame()
    rt.synchronize()
    first = rt.read_visibility()
    n_hit = int((first != np.uint64(0xFFFFFFFFFFFFFFFF)).sum())
    rays_per_frame = WIDTH * HEIGHT + n_hit  # primary + one secondary (GI) ray per hit pixel

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed: inputs resident in HBM, CUDA events on the library's stream, L2 flushed between steps ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stage = {"clear_ms": 0.0, "cull_ms": 0.0, "visibility_ms": 0.0, "merge_ms": 0.0, "shading_ms": 0.0}
    rt.reset_launch_counter()
    barrier()
    for i in range(args.steps):
        flush_l2()
        with torch.cuda.stream(stream):
            starts[i].record()
        frame()
        with torch.cuda.stream(stream):
            stops[i].record()
        t = rt.timings()  # synchronises the stream; per-stage CUDA events of this frame
        for k in stage:
            stage[k] += t[k]
    barrier()
    launches = rt.timings()["n_kernel_launches"]
    clocks = sampler.stop()
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    if world > 1:
        tt = torch.tensor([dev_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms = float(tt.item())
    ms_per_step = dev_ms / args.steps
    value = world * rays_per_frame / (ms_per_step * 1e-3) / 1e6

    # ---- end to end: the user's calls (clear, render, read the frame into HOST memory), copies inside the timed region ----
    barrier()
    t_e2e = 0.0
    for i in range(args.steps):
        flush_l2()
        rt.synchronize()
        t0 = time.perf_counter()
        frame()
        rt.read_radiance(host_rad_np)   # D2H of the step's result (the RGBA32F frame) into pinned host memory, synchronous
        t_e2e += time.perf_counter() - t0
    barrier()
    if world > 1:
        tt = torch.tensor([t_e2e], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
    e2e_value = world * rays_per_frame * args.steps / t_e2e / 1e6
    assert np.isfinite(host_rad_np).all()

    # ---- roofline of the dominant kernel stage (visibility: clear + cull/sort + K1), per frame ----
    peak, peak_src = measured_peak()
    alg_bytes = n_clusters * ALG_BYTES_PER_CLUSTER + n_objects * ALG_BYTES_PER_OBJECT + WIDTH * HEIGHT * ALG_BYTES_PER_PIXEL_VIS
    vis_stage_ms = (stage["clear_ms"] + stage["cull_ms"] + stage["visibility_ms"]) / args.steps
    achieved = alg_bytes / (vis_stage_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            v, cores, secs, n_rays = cpu_baseline_sample(build_scene(0, 1), 1)
            cpu = {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                   "sample": f"every {CPU_YSTEP}th scanline of the same 3840x2160 frame ({n_rays} primary rays, {secs:.2f} s wall), screen-rect pruned oracle, OpenMP"}
        line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u64", "data": "synthetic",
                "config": {"workload": f"BASELINE configs[1] per GPU: 1,024 objects (2^21 clusters, 1.07e9 voxels), 3840x2160 primary visibility"
                                       + (f"; world = {world} such shards, ncclAllReduce(u64,min) merge" if world > 1 else ""),
                           "rays_per_frame": rays_per_frame, "hit_pixels": n_hit, "gi": True, "svo_build_ms": svo_build_ms, "l2": "flushed between steps (256 MiB write, outside the timed events)",
                           "stage_ms": {k: v / args.steps for k, v in stage.items()}},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 96, "d2h_bytes_per_step": WIDTH * HEIGHT * 16,
                        "note": "clear + render (K1 + K3 GI) + read_radiance (RGBA32F frame) into pinned host memory through the C ABI"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                             "kernel": "visibility stage (k_clear_visibility + k_cull_objects + k_sort_frames + k_visibility)", "algorithmic_bytes": alg_bytes,
                             "stage_ms": vis_stage_ms, "peak_source": peak_src},
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    rt.destroy()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

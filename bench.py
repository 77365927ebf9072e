#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (primary+GI Mrays/s and ms/frame at 3840x2160, % of HBM roofline) for the
voxel-rendering hot path, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one frame of the hot path over the synthetic scene through libtgb200.so: clear + visibility (K1) + [N > 1:
ncclAllReduce(u64, min) merge + owner-resolved materials, reduce-scattered by screen tile] + GI / shading (K3: one
secondary ray per hit pixel through the replicated 1-bit SVO). The SVO (K2) is built once before the timed region, like
the reference builds it on its first frame (tgvk_raytracer.c:1187-1217); its build time is reported beside the frame.
Headline workload (c2) = BASELINE configs[1]/[2]: 1,024 objects, 2^21 clusters (1.07e9 voxels), 4K. At N>1 the SAME scene is dealt
out over the ranks by object (strong scaling: the frame's work is fixed), every rank traces the full 4K frame against its
objects and shades 1/N of the rows (16-row bands dealt out to the ranks). `value` = SURVEY 8(d) rays per second at every N:
W*H primary + one GI ray per hit pixel (that every rank casts its own W*H primaries against its shard is an implementation
detail reported as `rank_primary_rays`, not counted). The same invocation also times configs[3] (c4, N=1), configs[4] (c5: one
12,288-object shard per GPU, the 1e11-voxel world at N=8) and the dense-view stress c2far briefly and attaches them under
`also`; at N>1 every workload first proves, inside its warm-up, that the sharded frame equals the single-GPU frame
(`parity_check`).
`--impl reference` times the CPU path (the oracle port of the reference's shader logic; the reference itself is
Win32/Vulkan-only and cannot run here) on a bounded scanline sample of the same workload, rank 0 only.
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
CPU_YSTEP = 4          # CPU sample: every 4th scanline of the same frame (540 rows)
# SURVEY.md section 8d algorithmic bytes
ALG_BYTES_PER_CLUSTER = 72    # 64 B mask + 4 B pointer + 4 B cluster->object
ALG_BYTES_PER_OBJECT = 96
ALG_BYTES_PER_PIXEL_VIS = 16  # clear store + resolved store
ALG_BYTES_PER_PIXEL_GI = 40   # vis + ptr + c2o + LUT-idx word + LUT + RGBA32F out
CLEAR = np.uint64(0xFFFFFFFFFFFFFFFF)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture, or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name))).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.device)],
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
        time.sleep(0.12)
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


WORKLOADS = {
    "c2": "BASELINE configs[1]+[2]: 1,024 objects (2^21 clusters, 1.07e9 voxels), 3840x2160, primary visibility + 1-bounce SVO GI (1 spp)",
    "c4": "BASELINE configs[3]: the configs[1] scene with 64 objects (the nearest to the player) moving every frame (translation.x += 0.5, angle += 1 degree "
          "per frame, 120-frame cycle): 64 transform uploads + incremental SVO update + primary visibility + 1-bounce SVO GI (1 spp), 3840x2160, 1 GPU",
    "c4m8": "configs[3] with 8 instead of 64 movers (the 8 objects nearest to the player; same motion): what the incremental SVO update buys when only part of the "
            "+-512 box changes -- with 64 movers every object inside the box moves and every leaf is re-sampled",
    "c5": "BASELINE configs[4] per GPU: 12,288 objects (25,165,824 clusters, 1.29e10 voxels; 8 GPUs = 1.03e11 voxels) generated on the device, "
          "3840x2160, primary visibility + 1-bounce SVO GI (1 spp)",
    "c2far": "dense-view stress, beside the headline: the configs[1] scene (1,024 objects, 2^21 clusters) with the far plane at 4000 so that several hundred "
             "objects survive the cull instead of ~33, 3840x2160, primary visibility + 1-bounce SVO GI (1 spp)",
}
SCALING = {"c2": "strong", "c4": "strong", "c4m8": "strong", "c2far": "strong", "c5": "weak"}
N_MOVERS = {"c4": 64, "c4m8": 8}
C2FAR_FAR = 4000.0


def build_scene(rank, n_ranks, workload="c2", host_bits=True):
    """Rank's shard. c2 / c4 / c2far: the 32 x 32 lattice of configs[1] (1,024 objects), dealt out over the ranks -- the world is the
    same at every N (strong scaling). c5: 12,288 objects out of a 384 x 32N lattice per rank, masks generated on the device (weak
    scaling: the resident world grows with N). The lattice cells are dealt out in order of their distance from the camera, boustrophedon
    (scenes.deal_by_distance): every rank holds the same number of objects and an equal share of the near (large on screen) and far ones.
    The global lattice index decides seed, angle and position."""
    from tg_b200 import scenes
    if workload == "c5":
        return scenes.config5_shard(rank, n_ranks, WIDTH, HEIGHT)
    s = scenes.grid_scene(f"config2_{rank}of{n_ranks}", 32, 32, WIDTH, HEIGHT, k=3, with_bits=host_bits, owner=(rank, n_ranks) if n_ranks > 1 else None)
    if workload == "c2far":
        s.camera.far = C2FAR_FAR
    return s


def union_scene(n_ranks, workload):
    """The single-GPU scene whose frame the sharded frame must equal bit for bit: every rank's objects in rank-major order, so that
    a cluster's pointer is its global pointer (base of rank r = r * clusters per rank). c5: the whole world does not fit one GPU;
    only the objects that can write a pixel (within the far plane) are kept, so pointers differ and only depth / voxel / radiance
    are comparable (union_pointers_match = False)."""
    from tg_b200 import scenes
    shards = [build_scene(r, n_ranks, workload, host_bits=False) for r in range(n_ranks)]
    u = shards[0]
    objects = [o for sh in shards for o in sh.objects]
    pointers_match = True
    if workload == "c5":
        cam, far = u.camera.position, u.camera.far
        objects = [o for o in objects if ((o.center[0] - cam[0]) ** 2 + (o.center[2] - cam[2]) ** 2) ** 0.5 < far + 160.0]
        pointers_match = False
    return scenes.SceneSpec(name=f"union_{workload}", width=WIDTH, height=HEIGHT, camera=u.camera, objects=objects, lut=u.lut, n_luts=u.n_luts), pointers_match


def c4_movers(scene, n=64):
    """configs[3]: the n = 64 objects nearest to the player (objects 0..63 of SURVEY 8d lie outside the +-512 SVO box and would not
    exercise the update) and their pose at frame f: translation.x += 0.5 f, angle += 1 degree * f (float32, like the tests)."""
    from tg_b200 import scenes
    order = np.argsort([o.center[0] ** 2 + o.center[2] ** 2 for o in scene.objects], kind="stable")[:n]
    movers = [int(i) for i in order]

    def pose(i, f):
        o = scene.objects[i]
        return (o.center[0] + 0.5 * f, o.center[1], o.center[2]), float(np.float32(o.angle) + scenes.deg2rad(1.0) * np.float32(f))
    return movers, pose


def cpu_scene(workload):
    """The N=1 scene for the CPU arm. c5: only the objects that can write a pixel (within the far plane of the camera; the rest
    fail depth <= 1, visibility.frag:194) get host-side masks -- 1.6 GB of masks for the others would only be skipped."""
    from tg_b200 import scenes
    s = build_scene(0, 1, workload)
    if workload == "c5":
        cam, far = s.camera.position, s.camera.far
        near = [o for o in s.objects if ((o.center[0] - cam[0]) ** 2 + (o.center[2] - cam[2]) ** 2) ** 0.5 < far + 160.0]
        for o in near:
            o.bits = scenes.random_solid_bits(o.seed, o.n_clusters, o.k)
        s.objects = near
    return s


class CpuArm:
    """The CPU path on a bounded sample of the N=1 frame: every CPU_YSTEP-th scanline through the oracle's visibility pass
    (screen-rect pruned, OpenMP), then GI + shading of the same rows from that buffer with the oracle's SVO (built once,
    outside the timed region, like the GPU arm)."""

    def __init__(self, scene, dynamic=False, ystep=CPU_YSTEP, n_movers=64):
        from oracle import oracle as O
        self.O = O
        self.dynamic, self.scene, self.frame_idx, self.ystep, self.n_movers = dynamic, scene, 0, ystep, n_movers
        # all the host threads this process may use (torch.distributed.run exports OMP_NUM_THREADS=1 to its workers)
        O.lib().tgo_set_threads(len(os.sched_getaffinity(0)))
        self.cores = O.lib().tgo_max_threads()
        self.rays = O.camera_rays(O.camera_from_spec(scene.camera))
        self.view = O.SceneView.from_scene(scene, with_lut=True)
        # the SVO only sees objects that can touch the +-512 box (the others fail the SAT against every root child)
        self.svo = self._build_svo(scene)
        self.rows = np.arange(0, HEIGHT, ystep)
        self.frame = np.zeros((HEIGHT, WIDTH, 4), dtype=np.float32)

    def _build_svo(self, scene):
        from tg_b200 import scenes
        O = self.O
        near = [o for o in scene.objects if max(abs(o.center[0]), abs(o.center[2])) < 512 + 160]
        return O.svo_create(O.SceneView.from_scene(scenes.SceneSpec(name="near", width=WIDTH, height=HEIGHT, camera=scene.camera, objects=near), with_lut=False),
                            capacities=(1 << 25, 1 << 15, 1 << 16))

    def sample(self):
        """(seconds, rays) of one sample. dynamic (configs[3]): the movers take their next pose and the SVO is rebuilt inside the
        sample -- the reference has no incremental path (tg_svo_create from scratch, tgvk_raytracer.c:1187-1217)."""
        O = self.O
        t0 = time.perf_counter()
        if self.dynamic:
            import copy
            self.frame_idx = self.frame_idx % 120 + 1
            movers, pose = c4_movers(self.scene, self.n_movers)
            moved = copy.copy(self.scene)
            moved.objects = list(self.scene.objects)
            for i in movers:
                o = copy.copy(self.scene.objects[i])
                o.center, o.angle = pose(i, self.frame_idx)
                moved.objects[i] = o
            self.view = O.SceneView.from_scene(moved, with_lut=True)
            O.svo_destroy(self.svo)
            self.svo = self._build_svo(moved)
        vis, _ = O.visibility(self.view, self.rays, WIDTH, HEIGHT, O.VIS_SCREEN_RECT, 0, HEIGHT, self.ystep)
        O.shade(self.view, self.rays, WIDTH, HEIGHT, vis, self.svo, gi=True, frame_seed=1, y0=0, y1=HEIGHT, ystep=self.ystep, out=self.frame)
        dt = time.perf_counter() - t0
        return dt, len(self.rows) * WIDTH + int((vis[self.rows] != CLEAR).sum())

    def text(self, n_rays):
        return ((f"{self.n_movers} objects moved + oracle tg_svo_create from scratch (whole tree, not sampled) + " if self.dynamic else "")
                + f"every {self.ystep}th scanline of the 3840x2160 frame ({len(self.rows)} rows): oracle visibility (screen-rect pruned) + oracle GI/shading of "
                f"those rows, {n_rays} rays per sample, OpenMP")

    def close(self):
        self.O.svo_destroy(self.svo)


def run_reference(args, rank):
    """The CPU arm (rank 0 only; the other ranks exit without work)."""
    if rank != 0:
        return
    arm = CpuArm(cpu_scene(args.workload), dynamic=args.workload in N_MOVERS, ystep=CPU_YSTEP * (4 if args.workload == "c2far" else 1), n_movers=N_MOVERS.get(args.workload, 0))
    times, n_rays = [], 0
    for i in range(args.warmup + args.steps):
        secs, n_rays = arm.sample()
        if i >= args.warmup:
            times.append(secs)
    text = arm.text(n_rays)
    arm.close()
    total = sum(times)
    value = n_rays * len(times) / total / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": SCALING[args.workload], "vs_baseline": None, "dtype": "f32+u64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "sample": text},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": arm.cores, "kind": "port", "sample": text},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


class Ctx:
    """One process of the job: rank / world / device, and the few collectives bench.py itself needs."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x, op="max"):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN, "sum": self.dist.ReduceOp.SUM}[op])
        return float(t.item())

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def parity_check(ctx, rt, frame, workload, merge):
    """N > 1, inside the warm-up: (1) the frame merged by the one-kernel peer-memory exchange equals, word for word and radiance bit
    for radiance bit, the frame merged by the NCCL collectives; (2) the sharded frame equals the frame ONE GPU renders from the
    union of the shards (every rank renders that union on its own GPU and compares the whole visibility buffer and the whole
    gathered radiance frame). Returns the dict attached to the bench line; every entry is the AND over all ranks."""
    from tg_b200.raytracer import from_scene

    def sharded(kind):
        frame(kind)
        rt.synchronize()
        ctx.barrier()
        vis = rt.read_visibility()
        rt.gather_radiance()
        rt.synchronize()
        rad = rt.read_radiance().view(np.uint32).copy()
        ctx.barrier()
        return vis, rad

    rt.set_merge_kind(0)          # the peer-memory exchange may run (mapped on first use, collectively) whatever --merge says for the timed frames
    vis_p, rad_p = sharded("peer")
    vis_n, rad_n = sharded("nccl")
    rt.set_merge_kind(0 if merge == "peer" else 1)
    out = {"merged_eq_nccl": bool(np.array_equal(vis_p, vis_n) and np.array_equal(rad_p, rad_n)), "peer_memory_path_ran": bool(rt.timings()["merge_ms"] > 0)}
    union, pointers_match = union_scene(ctx.world, workload)
    ref = from_scene(union, device=ctx.local_rank)
    try:
        ref.set_gi(True, 1)
        ref.clear()
        ref.render()
        ref.synchronize()
        want_vis, want_rad = ref.read_visibility(), ref.read_radiance().view(np.uint32)
    finally:
        ref.destroy()
    if pointers_match:
        out["sharded_eq_single"] = bool(np.array_equal(vis_p, want_vis) and np.array_equal(rad_p, want_rad))
        out["compared"] = "every u64 visibility word and every radiance bit of the 3840x2160 frame, on every rank, against one GPU rendering the union of the shards"
    else:
        # same depth24 and voxel for every pixel; the pointer field numbers clusters differently in the reduced union
        field = np.uint64(0xFFFFFF00000001FF)
        same_words = np.array_equal(vis_p & field, want_vis & field)
        n_rad = int((rad_p != want_rad).any(axis=-1).sum())
        out["sharded_eq_single"] = bool(same_words and n_rad == 0)
        out["radiance_pixels_differing"] = n_rad
        out["compared"] = ("depth24 + voxel9 of every visibility word and every radiance bit of the 3840x2160 frame, on every rank, against one GPU rendering "
                           f"the {len(union.objects)} objects of the world that lie within the far plane (the whole world does not fit one GPU; pointers number "
                           "clusters differently there)")
    for k in ("merged_eq_nccl", "peer_memory_path_ran", "sharded_eq_single"):
        out[k] = bool(ctx.reduce(1.0 if out[k] else 0.0, "min") > 0.5)
    return out


def gpu_arm(ctx, args, workload, steps, warmup, headline):
    """One workload through libtgb200.so on this process's GPU; rank 0 returns the result dict, the others None."""
    import ctypes as C
    import tg_b200
    from tg_b200.raytracer import comm_unique_id, from_scene
    torch, rank, world, local_rank, dev = ctx.torch, ctx.rank, ctx.world, ctx.local_rank, ctx.dev

    scene = build_scene(rank, world, workload, host_bits=False)  # masks generated on the device (same seeds as the host definition)
    rt = from_scene(scene, device=local_rank)
    n_clusters, n_objects = scene.n_clusters, len(scene.objects)
    if world > 1:
        rt.set_shard(rank, world, rank * n_clusters)
        ids = [comm_unique_id() if rank == 0 else None]
        ctx.dist.broadcast_object_list(ids, src=0)
        rt.comm_init(ids[0], rank, world)
        rt.set_merge_kind(0 if args.merge == "peer" else 1)
    y0, y1 = rt.tile_rows()

    lib = tg_b200.lib()
    stream = torch.cuda.ExternalStream(lib.tgb200_stream(C.byref(rt._rt)), device=dev)
    flush = ctx.flush

    # static scene: the SVO is built before the timed frames (collective when sharded); second build = warm number
    rt.set_gi(True, 1)
    rt.svo_update(force_full=True)
    rt.synchronize()
    rt.svo_update(force_full=True)
    rt.synchronize()
    svo_build_ms = rt.timings()["svo_ms"]

    dynamic = workload in N_MOVERS
    assert not (dynamic and world > 1), "configs[3] is a 1-GPU configuration"
    movers, pose = c4_movers(scene, N_MOVERS[workload]) if dynamic else ([], None)
    frame_idx = [0]

    def move_objects():
        """configs[3]: 64 x tg_raytracer_set_object_transform (96-byte record upload each); marks the SVO for an incremental update."""
        frame_idx[0] = frame_idx[0] % 120 + 1
        for i in movers:
            center, angle = pose(i, frame_idx[0])
            rt.set_object_transform(i, center, angle)

    def frame(merge=None):
        merge = merge or args.merge
        if dynamic:
            move_objects()
        rt.clear()
        rt.render_visibility()
        if world > 1 and merge == "nccl":
            rt.merge_visibility()   # else the shading stage merges this rank's tile straight from the peers' buffers
        if dynamic:
            rt.svo_update()   # incremental: only the leaves the moved objects touch are re-sampled
        rt.render_shading()

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.fill_(1)

    # ---- warm-up ----
    for _ in range(warmup):
        flush_l2()
        frame()
    rt.synchronize()
    parity = parity_check(ctx, rt, frame, workload, args.merge) if world > 1 else None
    if parity is not None:   # the check changed the exchange kind twice: one more warm frame on the timed path
        flush_l2()
        frame()
        rt.synchronize()
    n_hit = int((rt.read_visibility() != CLEAR).sum())   # merged buffer: identical on every rank
    rays_per_frame = WIDTH * HEIGHT + n_hit               # W*H primary + one GI ray per hit pixel, at every N

    # ---- device-timed: inputs resident in HBM, CUDA events on the library's stream, L2 flushed between steps ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    stage = {"clear_ms": 0.0, "cull_ms": 0.0, "visibility_ms": 0.0, "merge_ms": 0.0, "shading_ms": 0.0}
    if world > 1 and args.merge == "peer":
        stage.update({"merge_resolve_ms": 0.0, "merge_gather_ms": 0.0, "merge_kernel_ms": 0.0})  # parts of merge_ms (rank 0)
    if dynamic:
        stage["svo_ms"] = 0.0
    leaves_resampled = 0
    rt.reset_launch_counter()
    ctx.barrier()
    for i in range(steps):
        flush_l2()
        with torch.cuda.stream(stream):
            starts[i].record()
        frame()
        with torch.cuda.stream(stream):
            stops[i].record()
        t = rt.timings()  # synchronises the stream; per-stage CUDA events of this frame
        for k in stage:
            stage[k] += t[k]
        if dynamic:
            leaves_resampled += rt.svo_leaves_resampled()
    ctx.barrier()
    if dynamic:  # hit pixels drift as the objects move: mean of the count before and after the timed frames
        n_hit = (n_hit + int((rt.read_visibility() != CLEAR).sum())) // 2
        rays_per_frame = WIDTH * HEIGHT + n_hit
    last = rt.timings()
    launches = last["n_kernel_launches"]
    clocks = sampler.stop()
    dev_ms = ctx.reduce(sum(s.elapsed_time(e) for s, e in zip(starts, stops)), "max")
    ms_per_step = dev_ms / steps
    value = rays_per_frame / (ms_per_step * 1e-3) / 1e6

    # ---- end to end: the user's calls with HOST buffers. Every frame: tg_raytracer_clear + tg_raytracer_render with a frame sink
    # (tgb200_set_frame_sink): the shaded RGBA32F rows are copied into pinned host memory on a copy stream while the next frame
    # renders; two host buffers alternate, frame i is awaited (tgb200_wait_frame) after frame i+1 has been submitted. All
    # copies and the L2 flush of every frame are inside the timed region.
    rows = max(y1 - y0, 1)
    host_tiles = [torch.empty(rows * WIDTH * 4, dtype=torch.float32).pin_memory().numpy().reshape(rows, WIDTH, 4)[:y1 - y0] for _ in range(2)]
    present_tiles = [torch.empty(rows * WIDTH, dtype=torch.int32).pin_memory().numpy().view(np.uint32).reshape(rows, WIDTH)[:y1 - y0] for _ in range(2)]
    bands = 1          # throughput: whole-frame copies behind the next frame's rendering (double-buffered radiance on the device)
    latency_bands = 4  # latency: four row bands, each copied while the next is shaded

    def e2e_loop(n, tiles):
        tickets = []
        for i in range(n):
            flush_l2()
            rt.set_frame_sink(tiles[i % 2], bands)
            if dynamic:
                move_objects()                        # render() then updates the SVO incrementally before shading
            rt.clear()
            rt.render()                               # tg_raytracer_render: K1 [+ merge + resolve] + K3 in bands, each band copied D2H as it completes
            tickets.append(rt.frame_ticket())
            if i >= 1:
                rt.wait_frame(tickets[i - 1])
        rt.wait_frame(tickets[-1])

    def e2e_timed(tiles):
        e2e_loop(2, tiles)
        rt.synchronize()
        ctx.barrier()
        t0 = time.perf_counter()
        e2e_loop(steps, tiles)
        dt = time.perf_counter() - t0
        ctx.barrier()
        return ctx.reduce(dt, "max")

    t_e2e = e2e_timed(host_tiles)
    e2e_value = rays_per_frame * steps / t_e2e / 1e6
    assert np.isfinite(host_tiles[0]).all() and np.isfinite(host_tiles[1]).all()
    # the same loop with the reference's own end of frame: the present pass (present.frag -> B8G8R8A8_UNORM swapchain image), 4 B / pixel
    t_present = e2e_timed(present_tiles)
    assert present_tiles[0].any() and present_tiles[1].any()
    frame_latency_ms = None
    if headline:
        # latency of ONE frame from the first call to the last byte in host memory (band overlap only, nothing in flight before it)
        lat = []
        for i in range(5):
            flush_l2(); rt.synchronize()
            t0 = time.perf_counter()
            if dynamic:
                move_objects()
            rt.set_frame_sink(host_tiles[0], latency_bands); rt.clear(); rt.render(); rt.wait_frame(rt.frame_ticket())
            lat.append(time.perf_counter() - t0)
        frame_latency_ms = 1e3 * float(np.median(lat))
    rt.set_frame_sink(None)
    rt.synchronize()

    # ---- roofline: the dominant stage is K3 (GI + shading); the visibility stage is reported beside it ----
    peak, peak_src = measured_peak()
    svo_bytes = 0
    try:
        svo, nodes, leaf, vox = rt.svo_download()
        svo_bytes = nodes.nbytes + leaf.nbytes + vox.nbytes
        rt.svo_free(svo)
    except Exception:
        pass
    tile_px = WIDTH * int((rt.tile_physical_rows() >= 0).sum())  # the rows this rank shades (16-row bands dealt out to the ranks)
    n_objects_world = int(ctx.reduce(float(n_objects), "sum"))
    gi_bytes = tile_px * ALG_BYTES_PER_PIXEL_GI + n_objects_world * ALG_BYTES_PER_OBJECT + svo_bytes
    gi_ms = stage["shading_ms"] / steps
    vis_bytes = n_clusters * ALG_BYTES_PER_CLUSTER + n_objects * ALG_BYTES_PER_OBJECT + WIDTH * HEIGHT * ALG_BYTES_PER_PIXEL_VIS
    vis_ms = (stage["clear_ms"] + stage["cull_ms"] + stage["visibility_ms"]) / steps
    gi_achieved, vis_achieved = gi_bytes / (gi_ms * 1e-3) / 1e9, vis_bytes / (vis_ms * 1e-3) / 1e9

    def roofline(kernel, alg_bytes, achieved, stage_ms, traffic_file):
        # `frac` follows SURVEY 8(d)'s accounting (every submitted cluster counts, touched or not); `frac_touched` is what the kernel really
        # moves through DRAM (ncu dram__bytes_read + write of the committed capture, c2 at N=1) over the same time: the honest HBM figure
        traffic = profiled_traffic(traffic_file) if workload in ("c2", "c4", "c4m8") and world == 1 else None
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "frac_touched": (traffic / (stage_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "kernel": kernel, "algorithmic_bytes": alg_bytes, "stage_ms": stage_ms, "peak_source": peak_src,
                "note": "issue-bound under divergence, not bandwidth-bound: frac counts SURVEY 8(d) bytes whether or not the kernel touches them; "
                        "frac_touched = measured DRAM traffic / stage time / peak"}

    result = None
    if rank == 0:
        cpu = None
        if headline and not args.no_cpu_baseline and world == 1:  # the CPU baseline is an N=1 figure
            arm = CpuArm(cpu_scene(workload), dynamic=dynamic, n_movers=N_MOVERS.get(workload, 0))
            samples = [arm.sample() for _ in range(12)]  # ~10-30 s of CPU work on the box host cores
            secs = sum(t for t, _ in samples)
            cpu = {"value": sum(n for _, n in samples) / secs / 1e6, "unit": "Mrays/s", "cores": arm.cores, "kind": "port",
                   "sample": f"12 x ({arm.text(samples[0][1])}); {secs:.1f} s wall"}
            arm.close()
        e2e_note = ("per frame: tg_raytracer_clear + tg_raytracer_render through the C ABI with a frame sink: the shaded RGBA32F rows go to pinned host memory "
                    "on a copy stream while the next frame renders (radiance double-buffered on the device, two host buffers alternate, frame i is awaited "
                    "after frame i+1 was submitted; every rank receives its own tile). All copies and the per-frame L2 flush are inside the timed region. "
                    "The scene arrays stay resident in HBM like the reference's SSBOs, the camera block is the per-frame input")
        if world == 1:
            gbs = WIDTH * HEIGHT * 16 / (t_e2e / steps) / 1e9
            e2e_note += (f". PCIe-bound: the 133 MB HDR frame reaches host memory at {gbs:.0f} GB/s (the link's practical ceiling is ~55), which is what separates this "
                         "figure from the device-timed frame; with the reference's own end of frame (presented_bgra8, 33 MB) the loop follows the device time")
        if world > 1:
            e2e_note += (f". Host-ingest bound at N = {world}: the {world} ranks together deliver the 133 MB HDR frame ({133 // world} MB each) into ONE host's memory "
                         "every frame; the gap between e2e and the device-timed frame is that PCIe / host-memory ingest (it shrinks 4x with the 4 B / pixel "
                         "presented frame, see presented_bgra8), not GPU work")
        result = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": SCALING[workload], "vs_baseline": None, "dtype": "f32+u64", "data": "synthetic",
            "config": {"workload": WORKLOADS[workload]
                                   + (f"; the objects are dealt out over {world} ranks (in order of distance from the camera, boustrophedon), GI split by screen tile, merge = "
                                      + ("one kernel over peer memory (min + winner's material per tile, NVLink)" if args.merge == "peer" else "ncclAllReduce(u64,min) + material reduce-scatter")
                                      if world > 1 else ""),
                       "rays_per_frame": rays_per_frame, "primary_rays": WIDTH * HEIGHT, "gi_rays": n_hit, "rank_primary_rays": world * WIDTH * HEIGHT,
                       "objects_world": n_objects_world, "clusters_per_rank": n_clusters,
                       "svo_build_ms": svo_build_ms, "svo_bytes": svo_bytes, "visible_objects": last["n_visible_objects"],
                       **({"svo_leaves_resampled_per_frame": leaves_resampled / steps, "moved_objects_per_frame": len(movers)} if dynamic else {}),
                       "gi_work_rank0": {"rays_traced_through_svo": last["n_gi_rays"], "of_those_by_the_exact_kernel": last["n_gi_rays_exact"], "node_visits": last["n_gi_node_visits"], "dda_steps": last["n_gi_dda_steps"], "advances": last["n_gi_advances"]},
                       "l2": "flushed between steps (256 MiB write, outside the timed events)", "stage_ms": {k: v / steps for k, v in stage.items()}},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 96 * world + 96 * len(movers), "d2h_bytes_per_step": WIDTH * HEIGHT * 16,
                    "ms_per_step": 1e3 * t_e2e / steps, "frame_latency_ms": frame_latency_ms,
                    "presented_bgra8": {"value": rays_per_frame * steps / t_present / 1e6, "unit": "Mrays/s", "ms_per_step": 1e3 * t_present / steps,
                                        "d2h_bytes_per_step": WIDTH * HEIGHT * 4,
                                        "note": "same loop, the frame delivered as the reference delivers it (present.frag into the B8G8R8A8_UNORM swapchain image): "
                                                "k_present per band + 4 B / pixel over PCIe instead of the 16 B / pixel HDR rows"},
                    "note": e2e_note},
            "gpu_launches": int(launches),
            "roofline": roofline("GI + shading stage (k_object_frames + k_shade + k_gi_trace_fast + k_gi_trace_list" + (" + k_resolve_material + exchange" if world > 1 else "") + ")",
                                 gi_bytes, gi_achieved, gi_ms, "k3_traffic.json"),
            "roofline_visibility": roofline("visibility stage (k_clear_visibility + k_cull_objects + k_sort_frames + k_visibility)", vis_bytes, vis_achieved, vis_ms, "k1_traffic.json"),
            "cpu_baseline": cpu}
        if parity is not None:
            result["parity_check"] = parity
    rt.comm_destroy()
    rt.destroy()
    ctx.barrier()
    return result


def compact(r):
    """What an `also` workload contributes to the headline line."""
    keep = {k: r[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "scaling", "gpu_launches")}
    keep["workload"] = r["config"]["workload"]
    for k in ("rays_per_frame", "gi_rays", "visible_objects", "stage_ms", "svo_build_ms", "svo_leaves_resampled_per_frame", "moved_objects_per_frame", "objects_world", "clusters_per_rank"):
        if k in r["config"]:
            keep[k] = r["config"][k]
    keep["fps"] = 1e3 / r["ms_per_step"]
    keep["e2e"] = {k: r["e2e"][k] for k in ("value", "ms_per_step")}
    keep["e2e"]["presented_bgra8_ms_per_step"] = r["e2e"]["presented_bgra8"]["ms_per_step"]
    keep["roofline_frac"] = r["roofline"]["frac"]
    keep["roofline_visibility_frac"] = r["roofline_visibility"]["frac"]
    if "parity_check" in r:
        keep["parity_check"] = r["parity_check"]
    return keep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tg_b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--merge", default="peer", choices=["peer", "nccl"], help="N > 1: peer = one kernel over peer memory (NVLink, falls back to nccl if unmappable); "
                    "nccl = ncclAllReduce(u64, min) + materials + ncclReduceScatter")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="c2 = the headline configuration (default); c4 = configs[3]; c5 = one 12,288-object shard of the "
                    "1e11-voxel world per GPU; c2far = dense-view stress")
    ap.add_argument("--also", default=None, help="comma-separated workloads timed briefly in the same invocation and attached under `also` "
                    "(default with --workload c2: c4, c4m8 and c2far at N=1, c5 at every N; 'none' disables)")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
        return
    args.warmup = max(args.warmup, 3)

    ctx = Ctx()
    line = gpu_arm(ctx, args, args.workload, args.steps, args.warmup, headline=True)
    if args.also is None:
        also = (["c4", "c4m8", "c2far", "c5"] if ctx.world == 1 else ["c5"]) if args.workload == "c2" else []
    else:
        also = [w for w in args.also.split(",") if w and w != "none"]
    extra = {}
    for w in also:
        if w in N_MOVERS and ctx.world > 1:
            continue
        r = gpu_arm(ctx, args, w, max(5, min(args.steps, 10)), 3, headline=False)
        if r is not None:
            extra[w] = compact(r)
    if ctx.rank == 0:
        if extra:
            line["also"] = extra
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Knob sweeps on ONE resident scene (GPU box): every configuration = a set of TGB_* environment knobs (the library reads
them at every frame), timed like bench.py times a frame (CUDA events of the library's stages, L2 flushed between frames)
and checked bit for bit against the frame of the first configuration (visibility words and radiance bits).

    python tools/sweep.py --workload c2 --what gi|k1|all [--frames 12] [--rows H] > gpurun_out/sweep.jsonl
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GI_SWEEP = [
    {"TGB_GI_KERNEL": 1},                                   # round-1 kernel: one ray per lane
    {"TGB_GI_KERNEL": 0, "TGB_GI_TRAVERSAL_STACK": 1},      # the shader's stack machine (the exactness witness)
    {"TGB_GI_KERNEL": 2},                                   # pool, defaults (K = 4)
    {"TGB_GI_KERNEL": 2, "TGB_GI_RAYS_PER_LANE": 1},
    {"TGB_GI_KERNEL": 2, "TGB_GI_RAYS_PER_LANE": 2},
    {"TGB_GI_KERNEL": 2, "TGB_GI_RAYS_PER_LANE": 3},
    {"TGB_GI_KERNEL": 2, "TGB_GI_RAYS_PER_LANE": 6},
    {"TGB_GI_KERNEL": 2, "TGB_GI_RAYS_PER_LANE": 8},
    {"TGB_GI_KERNEL": 2, "TGB_GI_RAYS_PER_LANE": 4, "TGB_GI_POOL_DDA_STEPS": 8},
    {"TGB_GI_KERNEL": 2, "TGB_GI_RAYS_PER_LANE": 4, "TGB_GI_POOL_DDA_STEPS": 32},
    {"TGB_GI_KERNEL": 2, "TGB_GI_RAYS_PER_LANE": 4, "TGB_GI_POOL_TREE_REPS": 2},
    {"TGB_GI_KERNEL": 2, "TGB_GI_RAYS_PER_LANE": 4, "TGB_GI_POOL_TREE_REPS": 8},
    {"TGB_GI_KERNEL": 2, "TGB_GI_RAYS_PER_LANE": 4, "TGB_GI_POOL_SERVICE_SLOTS": 16},
    {"TGB_GI_KERNEL": 2, "TGB_GI_RAYS_PER_LANE": 4, "TGB_GI_POOL_SERVICE_SLOTS": 64},
    {"TGB_GI_KERNEL": 2, "TGB_GI_RAYS_PER_LANE": 4, "TGB_GI_POOL_CTAS_PER_SM": 4},
    {"TGB_GI_KERNEL": 2, "TGB_GI_RAYS_PER_LANE": 2, "TGB_GI_POOL_CTAS_PER_SM": 12},
    {"TGB_GI_KERNEL": 2, "TGB_GI_RAYS_PER_LANE": 2, "TGB_GI_POOL_CTAS_PER_SM": 16},
]
FAST_SWEEP = [
    {"TGB_GI_KERNEL": 2},                                   # the exact kernel on every ray (the frame every other line must equal)
    {},                                                     # default: certified walk over the coarser tiling + exact list kernel (TGB_GI_KERNEL=4)
    {"TGB_GI_KERNEL": 3},                                   # the certified walk's first form, over the octree's cells
    {"TGB_GI_SHADE_STEPS": 0},                              # every ray queued (k_shade does not enter the first cell)
    {"TGB_GI_SHADE_STEPS": 2},
    {"TGB_SHADE_MIN_CTAS": 4},
    {"TGB_GI_LIST_KERNEL": 0},                              # handed-over rays through the pool kernel's list mode
    {"TGB_GI_LIST_TMA": 0},                                 # leaf blocks staged by 32 loads per lane instead of one bulk copy
    {"TGB_GI_FAST_CAREFUL": 1},                             # careful second pass (cube check) in front of the exact list kernel
    {"TGB_GI_FAST_STEPS": 2},
    {"TGB_GI_FAST_STEPS": 8},
    {"TGB_GI_FAST_SERVICE_LANES": 4},
    {"TGB_GI_FAST_SERVICE_LANES": 16},
    {"TGB_GI_FAST_CTAS_PER_SM": 6},
    {"TGB_GI_FAST_MAX_STEPS": 64, "TGB_GI_FAST_MAX_STEPS_UNCERTAIN": 16},
    {"TGB_GI_FAST_DELTA_PERCENT": 25},                      # a quarter of the certificate's margin: the frame must still be the exact kernel's
    {"TGB_GI_FAST_DELTA_PERCENT": 10},
]
K1_SWEEP = [
    {"TGB_K1_KERNEL": 1},                                   # round-1 kernel: one pixel per lane
    {"TGB_K1_KERNEL": 2},                                   # pool, defaults (K = 2, 4 CTAs / SM)
    {"TGB_K1_STAGE_MASKS": 1},                              # cluster masks staged in shared memory by cp.async (north_star (a) as written)
    {"TGB_K1_KERNEL": 2, "TGB_K1_PIXELS_PER_LANE": 1},
    {"TGB_K1_KERNEL": 2, "TGB_K1_PIXELS_PER_LANE": 1, "TGB_K1_POOL_MIN_CTAS": 3},
    {"TGB_K1_KERNEL": 2, "TGB_K1_PIXELS_PER_LANE": 2, "TGB_K1_POOL_MIN_CTAS": 3},
    {"TGB_K1_KERNEL": 2, "TGB_K1_PIXELS_PER_LANE": 3},
    {"TGB_K1_KERNEL": 2, "TGB_K1_PIXELS_PER_LANE": 3, "TGB_K1_POOL_MIN_CTAS": 4},
    {"TGB_K1_KERNEL": 2, "TGB_K1_PIXELS_PER_LANE": 4},
    {"TGB_K1_KERNEL": 2, "TGB_K1_PIXELS_PER_LANE": 4, "TGB_K1_POOL_MIN_CTAS": 2},
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--what", default="gi")
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--rows", type=int, default=0, help="shade only ROWS rows starting at --row0 (emulates the screen tile of a sharded frame)")
    ap.add_argument("--row0", type=int, default=0)
    ap.add_argument("--shard", default="", help="r,n: the scene of rank r of n (its 1/n of the objects) on this one GPU: K1 as a rank of a sharded frame sees it")
    ap.add_argument("--configs", default="", help="JSON list of knob dicts (overrides --what)")
    args = ap.parse_args()

    import torch
    import bench
    import tg_b200
    from tg_b200.raytracer import from_scene

    shard = tuple(int(v) for v in args.shard.split(",")) if args.shard else (0, 1)
    scene = bench.build_scene(shard[0], shard[1], args.workload, host_bits=False)
    rt = from_scene(scene, device=0)
    lib = tg_b200.lib()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(lib.tgb200_stream(C.byref(rt._rt)), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rt.set_gi(True, 1)
    rt.svo_update(force_full=True)
    rt.synchronize()

    if args.configs:
        configs = json.loads(args.configs)
    else:
        configs = {"gi": GI_SWEEP, "fast": FAST_SWEEP, "k1": K1_SWEEP, "all": GI_SWEEP + K1_SWEEP}[args.what]
    touched = sorted({k for c in configs for k in c})
    want_vis = want_rad = None
    for cfg in configs:
        for k in touched:
            os.environ.pop(k, None)
        for k, v in cfg.items():
            os.environ[k] = str(v)
        rt.set_gi_traversal(1 if cfg.get("TGB_GI_TRAVERSAL_STACK") else 0)
        stage = {"visibility_ms": [], "shading_ms": [], "cull_ms": [], "clear_ms": []}
        for i in range(args.frames + 3):
            with torch.cuda.stream(stream):
                flush.fill_(1)
            rt.clear()
            rt.render_visibility()
            if args.rows:
                lib.tgb200_render_shading_rows(C.byref(rt._rt), args.row0, args.row0 + args.rows)
            else:
                rt.render_shading()
            t = rt.timings()
            if i >= 3:
                for k in stage:
                    stage[k].append(t[k])
        vis, rad = rt.read_visibility(), rt.read_radiance().view(np.uint32)
        if args.rows:
            rad = rad[args.row0:args.row0 + args.rows]
        if want_vis is None:
            want_vis, want_rad = vis.copy(), rad.copy()
        out = {"config": cfg, **{k: float(np.median(v)) for k, v in stage.items()}, "shading_min_ms": float(np.min(stage["shading_ms"])),
               "vis_equal": bool(np.array_equal(vis, want_vis)), "radiance_bits_equal": bool(np.array_equal(rad, want_rad)),
               "radiance_close_1e-3": bool(np.allclose(rad.view(np.float32), want_rad.view(np.float32), rtol=1e-3, atol=1e-6)),
               "gi": {k: t[k] for k in ("n_gi_rays", "n_gi_rays_exact", "n_gi_node_visits", "n_gi_dda_steps", "n_gi_advances")}}
        print(json.dumps(out), flush=True)
    rt.destroy()


if __name__ == "__main__":
    main()

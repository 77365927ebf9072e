"""How much margin does the certified fast GI walk (tg_b200/csrc/tgb_gi_fast.cuh) have?

Runs the host builds of the fast walk and of the exact state machine (tests/cpu_sim; the latter is held against the oracle's
transcription of svo_functions.inc by tests/test_gi_walk_cpu.py) on the same rays and reports, per DELTA: the share of rays handed to
the exact kernel and the number of rays the fast walk DECIDED differently from the exact one. The product uses DELTA0 = 2e-4 (growing by 3e-5 per box entered); the sweep
goes down until disagreements appear, which shows how far below that the real sideways displacement of the shader's walk stays
(below ~3e-5 DELTA0 no longer covers the half ulp(1024) by which the walk's own origin is rounded: disagreements there are expected).

    python tools/gi_fast_margin.py [n_rays] [scene]      scene: g6 (6x6 objects, default) | small | c1
"""
import sys
import os
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tg_b200 import scenes           # noqa: E402
from tests import cpu_sim            # noqa: E402
from oracle import oracle as ORACLE   # noqa: E402


def gi_like_rays(rng, n, view_scene, spread):
    """origins on / just above / inside the objects' volume (where k_shade's secondary rays start), uniform directions"""
    o = np.empty((n, 3), dtype=np.float32)
    objs = view_scene.objects
    idx = rng.integers(0, len(objs), size=n)
    for i, ob in enumerate(objs):
        sel = np.flatnonzero(idx == i)
        if len(sel) == 0:
            continue
        half = np.array([ob.dims[0] * 4.0, ob.dims[1] * 4.0, ob.dims[2] * 4.0], dtype=np.float32)
        local = rng.uniform(-1.0, 1.0, size=(len(sel), 3)).astype(np.float32) * (half + 3.0)
        o[sel] = local + np.array(ob.center, dtype=np.float32)
    far = rng.uniform(-spread, spread, size=(n // 8, 3)).astype(np.float32)   # an eighth anywhere in / around the box
    o[: n // 8] = far
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.sqrt((d.astype(np.float32) ** 2).sum(axis=1, dtype=np.float32))[:, None]
    return o, d.astype(np.float32)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    which = sys.argv[2] if len(sys.argv) > 2 else "g6"
    oracle = ORACLE
    oracle.lib()
    s = {"g6": lambda: scenes.grid_scene("g6", 6, 6, 64, 36, k=3), "small": lambda: scenes.small_grid(), "c1": lambda: scenes.config1(k=3, width=64, height=36)}[which]()
    view = oracle.SceneView.from_scene(s, with_lut=False)
    svo = oracle.svo_create(view)
    nodes, leaf, vox = oracle.svo_arrays(svo)
    grid = cpu_sim.flatten(nodes, leaf.view(np.uint32).ravel())
    voxels = vox.view(np.uint32).ravel()
    rng = np.random.default_rng(2024)
    o, d = gi_like_rays(rng, n, s, 600.0)
    far = np.float32(s.camera.far)
    bmin, bmax = (-512.0,) * 3, (512.0,) * 3
    t0 = time.time()
    exact, capped, work = cpu_sim.gi_trace(bmin, bmax, far, grid, voxels, o, d)
    t1 = time.time()
    print(f"{which}: {n} rays, exact walk {t1 - t0:.1f} s: occluded {exact.mean():.3f}, per ray {work[0] / n:.1f} look-ups {work[1] / n:.1f} DDA steps; capped {capped}")
    tiling = cpu_sim.fast_tiling(grid, voxels)
    for name, walk in (("octree cells (TGB_GI_KERNEL=3)", lambda delta: cpu_sim.gi_fast(bmin, bmax, far, grid, voxels, o, d, delta=delta)),
                       ("coarser tiling (TGB_GI_KERNEL=4, default)", lambda delta: cpu_sim.gi_fast_tiled(bmin, bmax, far, grid, voxels, tiling, o, d, delta=delta)),
                       ("coarser tiling, careful (cube check)", lambda delta: cpu_sim.gi_fast_tiled(bmin, bmax, far, grid, voxels, tiling, o, d, delta=delta, cube=True))):
        print(f" {name}")
        for delta in (1e-3, 2e-4, 1e-4, 5e-5, 2e-5, 1e-5, 3e-6, 1e-6, 0.0):
            res, w = walk(delta)
            decided = res != 2
            bad = np.flatnonzero(decided & ((res == 1) != exact))
            print(f"  delta {delta:8.1e}: handed over {1.0 - decided.mean():7.4f}  decided differently {len(bad):6d}  per ray {w[0] / n:.2f} boxes {w[1] / n:.2f} cells of leaf blocks"
                  + (f"   first: o={o[bad[0]].tolist()} d={d[bad[0]].tolist()}" if len(bad) else ""))


if __name__ == "__main__":
    main()

"""CPU: the certified fast walk of the GI pass (tg_b200/csrc/tgb_gi_fast.cuh: what k_gi_trace_fast runs per ray) compiled for the host
(tests/cpu_sim). Every ray it DECIDES must be decided like the shader's traversal (svo_functions.inc:1-329) decides it; the rays it
hands over are traced by the exact kernel and are not its responsibility. Two references: the oracle's transcription of the shader
(slow, a few thousand rays) and the host build of the exact state machine (tgb_gi_walk.cuh, itself held against the oracle by
tests/test_gi_walk_cpu.py; fast enough for a million rays and for a DELTA ten times below the product's)."""
import ctypes as C

import numpy as np
import pytest

from tg_b200 import ctypes_defs as T
from tg_b200 import scenes
from tests import cpu_sim
from tests.test_gi_walk_cpu import _rays

BMIN, BMAX = (-512.0,) * 3, (512.0,) * 3


def _svo(oracle, s):
    view = oracle.SceneView.from_scene(s, with_lut=False)
    svo = oracle.svo_create(view)
    nodes, leaf, vox = oracle.svo_arrays(svo)
    grid = cpu_sim.flatten(nodes, leaf.view(np.uint32).ravel())
    assert grid[-1] != 0
    return svo, grid, vox.view(np.uint32).ravel().copy()


def _surface_rays(rng, s, n):
    """origins in and just around the objects' volumes (where k_shade's secondary rays start), uniform directions"""
    o = np.empty((n, 3), dtype=np.float32)
    idx = rng.integers(0, len(s.objects), size=n)
    for i, ob in enumerate(s.objects):
        sel = np.flatnonzero(idx == i)
        half = np.array(ob.extent, dtype=np.float32) * 0.5
        o[sel] = rng.uniform(-1.0, 1.0, size=(len(sel), 3)).astype(np.float32) * (half + 3.0) + np.array(ob.center, dtype=np.float32)
    o[: n // 8] = rng.uniform(-600.0, 600.0, size=(n // 8, 3)).astype(np.float32)   # an eighth anywhere in / around the box
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.sqrt((d.astype(np.float32) ** 2).sum(axis=1, dtype=np.float32))[:, None]
    return o, d.astype(np.float32)


@pytest.mark.parametrize("make,spread", [(lambda: scenes.small_grid(), 300.0), (lambda: scenes.config1(k=3, width=64, height=36), 200.0),
                                         (lambda: scenes.grid_scene("g6", 6, 6, 64, 36, k=3), 600.0)])
@pytest.mark.parametrize("budgets", [(8,), (1,), (5,)])
def test_every_decided_ray_is_decided_like_the_shader(oracle, make, spread, budgets):
    s = make()
    svo, grid, voxels = _svo(oracle, s)
    try:
        rng = np.random.default_rng(4321)
        o, d = _rays(rng, 4000, spread)   # includes axis-parallel directions and origins on cell / voxel borders
        far = np.float32(s.camera.far)
        got, work = cpu_sim.gi_fast(BMIN, BMAX, far, grid, voxels, o, d, *budgets)
        L = oracle.lib()
        hp, hn, ni, vi = T.v3(), T.v3(), T.u32(), T.u32()
        want = np.zeros(len(o), dtype=bool)
        for i in range(len(o)):
            want[i] = L.tgo_svo_traverse_glsl(C.byref(svo), far, T.v3(*map(float, o[i])), T.v3(*map(float, d[i])), C.byref(hp), C.byref(hn), C.byref(ni), C.byref(vi)) < 1.0
        decided = got != 2
        bad = np.flatnonzero(decided & ((got == 1) != want))
        assert len(bad) == 0, f"{len(bad)} rays decided differently from the shader, first: o={o[bad[:2]].tolist()} d={d[bad[:2]].tolist()}"
        # axis-parallel directions (a tenth of the rays) are handed over by rule; of the others most must be decided here
        generic = np.abs(d).min(axis=1) >= 1.0e-3
        assert not (got[~generic] == 1).any()   # (a shallow ray that misses the root box is never queued: unoccluded)
        assert decided[generic].mean() > 0.6, decided[generic].mean()
        assert (got == 1).any() and (got == 0).any() and work[2] == (~decided).sum()
        # the step budget only suspends and resumes the walk (the step caps are looked at when a budget ends, so WHICH long rays are handed over
        # may depend on it; a decision may not)
        ref, _ = cpu_sim.gi_fast(BMIN, BMAX, far, grid, voxels, o, d, 64)
        both = (ref != 2) & decided
        assert np.array_equal(ref[both], got[both]) and both.sum() > 0.9 * decided.sum()
    finally:
        oracle.svo_destroy(svo)


@pytest.mark.parametrize("make", [lambda: scenes.grid_scene("g6", 6, 6, 64, 36, k=3), lambda: scenes.small_grid()])
def test_a_million_rays_with_a_tenth_of_the_margin(oracle, make):
    """DELTA0 = 2e-5 instead of the product's 2e-4: still no ray decided differently from the exact walk (tools/gi_fast_margin.py: the
    first disagreement out of 2e7 rays appears below 3e-6), and the product's DELTA hands over only a few per cent."""
    s = make()
    svo, grid, voxels = _svo(oracle, s)
    try:
        rng = np.random.default_rng(99)
        o, d = _surface_rays(rng, s, 1_000_000)
        far = np.float32(s.camera.far)
        exact, capped, _ = cpu_sim.gi_trace(BMIN, BMAX, far, grid, voxels, o, d)
        assert capped == 0
        for delta, most in ((2.0e-4, 0.12), (2.0e-5, 0.08)):
            got, work = cpu_sim.gi_fast(BMIN, BMAX, far, grid, voxels, o, d, delta=delta)
            decided = got != 2
            bad = np.flatnonzero(decided & ((got == 1) != exact))
            assert len(bad) == 0, f"delta {delta}: {len(bad)} rays decided differently, first: o={o[bad[:2]].tolist()} d={d[bad[:2]].tolist()}"
            assert 1.0 - decided.mean() < most, (delta, 1.0 - decided.mean())
    finally:
        oracle.svo_destroy(svo)

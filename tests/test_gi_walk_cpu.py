"""CPU: the per-ray traversal state machine the GI kernel runs (tg_b200/csrc/tgb_gi_walk.cuh: flattened tree, resumable
tree / DDA phases, 16-word ray state) compiled for the host (tests/cpu_sim) and held against the oracle's transcription of
svo_functions.inc:1-329 on random rays: the occluded / unoccluded decision must be the same for every ray, whatever the
phase budgets (the kernel's scheduler may suspend a ray after any number of steps)."""
import ctypes as C

import numpy as np
import pytest

from tg_b200 import ctypes_defs as T
from tg_b200 import scenes
from tests import cpu_sim


def _rays(rng, n, spread):
    """origins around and inside the SVO box, directions on the sphere plus axis-parallel / face-grazing / lattice-aligned ones"""
    o = rng.uniform(-spread, spread, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.sqrt((d.astype(np.float32) ** 2).sum(axis=1, dtype=np.float32))[:, None]
    # axis-parallel directions (zero components: t_delta = F32_MAX, slab quotients skipped)
    k = n // 10
    axis = rng.integers(0, 3, size=k)
    d[:k] = 0
    d[np.arange(k), axis] = rng.choice([-1.0, 1.0], size=k)
    # origins exactly on cell borders (multiples of 32) and on voxel borders (integers): the p == mid tie rules
    o[k:2 * k] = np.round(o[k:2 * k] / 32.0) * 32.0
    o[2 * k:3 * k] = np.round(o[2 * k:3 * k])
    return o, d.astype(np.float32)


@pytest.mark.parametrize("make,spread", [(lambda: scenes.small_grid(), 300.0), (lambda: scenes.config1(k=3, width=64, height=36), 200.0),
                                         (lambda: scenes.grid_scene("g6", 6, 6, 64, 36, k=3), 600.0)])
@pytest.mark.parametrize("budgets", [(4, 16), (1, 1), (3, 5), (64, 64)])
def test_state_machine_decides_like_the_shader(oracle, make, spread, budgets):
    s = make()
    view = oracle.SceneView.from_scene(s, with_lut=False)
    svo = oracle.svo_create(view)
    try:
        nodes, leaf, vox = oracle.svo_arrays(svo)
        grid = cpu_sim.flatten(nodes, leaf.view(np.uint32).ravel())
        assert grid[-1] != 0, "the tree must be tabulated completely"
        assert (grid[:-1] & 0x80000000).any(), "some cell must hold a leaf with data"
        rng = np.random.default_rng(1234)
        o, d = _rays(rng, 6000, spread)
        far = np.float32(s.camera.far)
        bmin, bmax = (-512.0, -512.0, -512.0), (512.0, 512.0, 512.0)
        got, capped, work = cpu_sim.gi_trace(bmin, bmax, far, grid, vox.view(np.uint32).ravel(), o, d, *budgets)
        assert capped == 0
        got16, _, work16 = cpu_sim.gi_trace(bmin, bmax, far, grid, vox.view(np.uint32).ravel(), o, d, *budgets, grid16=True)
        assert np.array_equal(got16, got) and np.array_equal(work16, work), "the 16-bit form of the flattened tree must describe the same walk"
        got_u, _, work_u = cpu_sim.gi_trace(bmin, bmax, far, grid, vox.view(np.uint32).ravel(), o, d, *budgets, uniform_dda=True)
        assert np.array_equal(got_u, got) and np.array_equal(work_u, work), "the branch form of the leaf DDA (k_gi_trace_list) must take the same steps"
        L = oracle.lib()
        want = np.zeros(len(o), dtype=bool)
        hp, hn, ni, vi = T.v3(), T.v3(), T.u32(), T.u32()
        for i in range(len(o)):
            depth = L.tgo_svo_traverse_glsl(C.byref(svo), far, T.v3(*map(float, o[i])), T.v3(*map(float, d[i])), C.byref(hp), C.byref(hn), C.byref(ni), C.byref(vi))
            want[i] = depth < 1.0
        assert want.any() and (~want).any()
        bad = np.flatnonzero(got != want)
        assert len(bad) == 0, f"{len(bad)} of {len(o)} rays decided differently, first: {bad[:5].tolist()} o={o[bad[:2]].tolist()} d={d[bad[:2]].tolist()}"
        assert work[0] > 0 and work[1] > 0 and work[2] > 0
    finally:
        oracle.svo_destroy(svo)


def test_blocks_view_traversal_equals_the_oracle_bit_for_bit(oracle):
    """tgb_svo_traverse_stack (tg_b200/csrc/tgb_svo_traverse.cuh: what the BLOCKS view's primary-ray kernel runs per pixel) on the host
    against the oracle's transcription of svo_functions.inc: same result bits, leaf node index and voxel index for every ray --
    un-normalised directions included, like debug_visibility_svo.frag shoots them."""
    s = scenes.small_grid()
    view = oracle.SceneView.from_scene(s, with_lut=False)
    svo = oracle.svo_create(view)
    try:
        nodes, leaf, vox = oracle.svo_arrays(svo)
        rng = np.random.default_rng(7)
        o, d = _rays(rng, 4000, 90.0)
        aim = -o[2000:] + rng.uniform(-25.0, 25.0, size=(2000, 3)).astype(np.float32)   # half of the rays point at the objects
        d[2000:] = (aim / np.linalg.norm(aim, axis=1, keepdims=True)).astype(np.float32)
        d[1000:3000] *= rng.uniform(0.3, 1.0, size=(2000, 1)).astype(np.float32)   # un-normalised
        far = np.float32(s.camera.far)
        res, node, voxel, word = cpu_sim.svo_traverse(nodes, leaf.view(np.uint32).ravel(), vox.view(np.uint32).ravel(), (-512.0,) * 3, (512.0,) * 3, far, o, d)
        L = oracle.lib()
        hp, hn, ni, vi = T.v3(), T.v3(), T.u32(), T.u32()
        n_hit = 0
        for i in range(len(o)):
            depth = L.tgo_svo_traverse_glsl(C.byref(svo), far, T.v3(*map(float, o[i])), T.v3(*map(float, d[i])), C.byref(hp), C.byref(hn), C.byref(ni), C.byref(vi))
            assert np.float32(depth).view(np.uint32) == res[i:i + 1].view(np.uint32)[0] and ni.value == node[i] and vi.value == voxel[i], (i, depth, res[i], ni.value, node[i])
            n_hit += depth < 1.0
        assert 100 < n_hit < len(o) - 100, n_hit
        assert (word[node == 0xFFFFFFFF] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()   # a miss writes the clear value
    finally:
        oracle.svo_destroy(svo)

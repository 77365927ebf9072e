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


# ---- the same walk over the coarser tiling (tgb_gi_fast.cuh, second half; TGB_GI_KERNEL=4) ------------------------------------------------

def _boxes_of(cells):
    """(x0, y0, z0, x1, y1, z1) of every free table cell's box, and the leaf flag"""
    leaf = (cells >> 31) != 0
    x0, y0, z0 = cells & 31, (cells >> 5) & 31, (cells >> 10) & 31
    x1, y1, z1 = x0 + ((cells >> 15) & 31), y0 + ((cells >> 20) & 31), z0 + ((cells >> 25) & 31)
    return leaf, np.stack([x0, y0, z0, x1, y1, z1], axis=1).astype(np.int64)


@pytest.mark.parametrize("make", [lambda: scenes.small_grid(), lambda: scenes.grid_scene("g6", 6, 6, 64, 36, k=3), lambda: scenes.config1(k=3, width=64, height=36)])
def test_the_coarser_cells_tile_the_root(oracle, make):
    """What the certificate needs of the tiling: every free table cell lies in its box, all cells of a box name the same box, no box
    contains a leaf block with data (so boxes neither overlap nor hide a solid voxel); the same for the boxes of empty bricks inside
    every leaf block, and a brick is marked solid exactly when one of its 512 voxels is."""
    s = make()
    svo, grid, voxels = _svo(oracle, s)
    oracle.svo_destroy(svo)
    cells, bricks, columns = cpu_sim.fast_tiling(grid, voxels)
    has_data = (grid[:-1] >> 31) != 0
    leaf, box = _boxes_of(cells)
    assert np.array_equal(leaf, has_data) and np.array_equal(cells[leaf] & 0x0FFFFFFF, grid[:-1][leaf] & 0x0FFFFFFF)
    c = np.arange(32 ** 3)
    cx, cy, cz = c & 31, (c >> 5) & 31, c >> 10
    free = ~leaf
    inside = (box[:, 0] <= cx) & (cx <= box[:, 3]) & (box[:, 1] <= cy) & (cy <= box[:, 4]) & (box[:, 2] <= cz) & (cz <= box[:, 5])
    assert inside[free].all()
    occ = has_data.reshape(32, 32, 32)   # [z, y, x]
    label = np.where(free, cells, 0xFFFFFFFF).reshape(32, 32, 32)
    n_cells_in_boxes = 0
    for e in np.unique(cells[free]):
        b = _boxes_of(np.array([e], dtype=np.uint32))[1][0]
        sub = (slice(b[2], b[5] + 1), slice(b[1], b[4] + 1), slice(b[0], b[3] + 1))
        assert not occ[sub].any() and (label[sub] == e).all(), f"box {b.tolist()} is not a cell of the tiling"
        n_cells_in_boxes += label[sub].size
    assert n_cells_in_boxes == int(free.sum())   # boxes are disjoint and cover every free cell
    assert len(np.unique(cells[free])) < 0.2 * free.sum()   # and they are coarse: that is their point
    # bricks
    n_leaves = len(voxels) // 1024
    rows = voxels.reshape(n_leaves, 32, 32)   # [leaf, z, y] words, bit x
    bits = ((rows[..., None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)   # [leaf, z, y, x]
    solid = bits.reshape(n_leaves, 4, 8, 4, 8, 4, 8).any(axis=(2, 4, 6))           # [leaf, bz, by, bx]
    br = bricks[: 64 * n_leaves].reshape(n_leaves, 4, 4, 4).astype(np.int64)
    assert np.array_equal((br >> 31) != 0, solid)
    for leaf_idx in range(n_leaves):
        e = br[leaf_idx]
        for entry in np.unique(e[~solid[leaf_idx]]):
            # corner and sides - 1 in voxels of the block: whole bricks
            vx0, vy0, vz0 = entry & 31, (entry >> 5) & 31, (entry >> 10) & 31
            vsx, vsy, vsz = ((entry >> 15) & 31) + 1, ((entry >> 20) & 31) + 1, ((entry >> 25) & 31) + 1
            assert all(v % 8 == 0 for v in (vx0, vy0, vz0, vsx, vsy, vsz))
            x0, y0, z0 = vx0 // 8, vy0 // 8, vz0 // 8
            x1, y1, z1 = x0 + vsx // 8 - 1, y0 + vsy // 8 - 1, z0 + vsz // 8 - 1
            sub = (slice(z0, z1 + 1), slice(y0, y1 + 1), slice(x0, x1 + 1))
            assert not solid[leaf_idx][sub].any() and (e[sub] == entry).all()
        assert sum(int((e == entry).sum()) for entry in np.unique(e[~solid[leaf_idx]])) == int((~solid[leaf_idx]).sum())


@pytest.mark.parametrize("make,spread", [(lambda: scenes.small_grid(), 300.0), (lambda: scenes.config1(k=3, width=64, height=36), 200.0),
                                         (lambda: scenes.grid_scene("g6", 6, 6, 64, 36, k=3), 600.0)])
@pytest.mark.parametrize("budget", [8, 1, 5])
def test_tiled_walk_decides_like_the_shader(oracle, make, spread, budget):
    s = make()
    svo, grid, voxels = _svo(oracle, s)
    try:
        tiling = cpu_sim.fast_tiling(grid, voxels)
        rng = np.random.default_rng(8765)
        o, d = _rays(rng, 4000, spread)
        far = np.float32(s.camera.far)
        got, work = cpu_sim.gi_fast_tiled(BMIN, BMAX, far, grid, voxels, tiling, o, d, budget)
        L = oracle.lib()
        hp, hn, ni, vi = T.v3(), T.v3(), T.u32(), T.u32()
        want = np.zeros(len(o), dtype=bool)
        for i in range(len(o)):
            want[i] = L.tgo_svo_traverse_glsl(C.byref(svo), far, T.v3(*map(float, o[i])), T.v3(*map(float, d[i])), C.byref(hp), C.byref(hn), C.byref(ni), C.byref(vi)) < 1.0
        decided = got != 2
        bad = np.flatnonzero(decided & ((got == 1) != want))
        assert len(bad) == 0, f"{len(bad)} rays decided differently from the shader, first: o={o[bad[:2]].tolist()} d={d[bad[:2]].tolist()}"
        generic = np.abs(d).min(axis=1) >= 1.0e-3
        assert decided[generic].mean() > 0.6 and (got == 1).any() and (got == 0).any() and work[2] == (~decided).sum()
    finally:
        oracle.svo_destroy(svo)


@pytest.mark.parametrize("make", [lambda: scenes.grid_scene("g6", 6, 6, 64, 36, k=3), lambda: scenes.small_grid()])
def test_tiled_walk_a_million_rays_and_fewer_steps(oracle, make):
    """Against the exact state machine on a million rays, with the product's margin and a fifth of it (DELTA0 below ~3e-5 no longer covers
    the half ulp(1024) by which the walk's own origin is rounded: disagreements are expected there and do appear, tools/gi_fast_margin.py);
    the coarser cells must hand over no more rays than the octree cells and enter far fewer cells."""
    s = make()
    svo, grid, voxels = _svo(oracle, s)
    try:
        tiling = cpu_sim.fast_tiling(grid, voxels)
        rng = np.random.default_rng(99)
        o, d = _surface_rays(rng, s, 1_000_000)
        far = np.float32(s.camera.far)
        exact, capped, _ = cpu_sim.gi_trace(BMIN, BMAX, far, grid, voxels, o, d)
        assert capped == 0
        plain, work_plain = cpu_sim.gi_fast(BMIN, BMAX, far, grid, voxels, o, d, delta=2.0e-4)
        for delta in (2.0e-4, 4.0e-5):
            got, work = cpu_sim.gi_fast_tiled(BMIN, BMAX, far, grid, voxels, tiling, o, d, delta=delta)
            decided = got != 2
            bad = np.flatnonzero(decided & ((got == 1) != exact))
            assert len(bad) == 0, f"delta {delta}: {len(bad)} rays decided differently, first: o={o[bad[:2]].tolist()} d={d[bad[:2]].tolist()}"
        got, work = cpu_sim.gi_fast_tiled(BMIN, BMAX, far, grid, voxels, tiling, o, d, delta=2.0e-4)
        assert (got == 2).sum() <= (plain == 2).sum()
        assert work[0] + work[1] < 0.7 * (work_plain[0] + work_plain[1])
        # the second stage: the handed-over rays once more with the cube check of near-edge steps (k_gi_trace_fast<.., CUBE>): still never a
        # wrong decision, and most of them decided
        for delta in (2.0e-4, 4.0e-5):
            first, _ = cpu_sim.gi_fast_tiled(BMIN, BMAX, far, grid, voxels, tiling, o, d, delta=delta)
            h = np.flatnonzero(first == 2)
            second, _ = cpu_sim.gi_fast_tiled(BMIN, BMAX, far, grid, voxels, tiling, o[h], d[h], delta=delta, cube=True)
            bad = np.flatnonzero((second != 2) & ((second == 1) != exact[h]))
            assert len(bad) == 0, f"delta {delta}: careful pass decided {len(bad)} rays differently, first: o={o[h][bad[:2]].tolist()} d={d[h][bad[:2]].tolist()}"
            assert (second != 2).mean() > 0.1, (delta, (second != 2).mean())
    finally:
        oracle.svo_destroy(svo)

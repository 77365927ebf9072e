"""GPU: K2 (tg_b200/csrc/tgb_svo.cu) through the C ABI against the oracle's literal restatement of the reference's CPU
builder (graphics/tg_sparse_voxel_octree.c:23-542), bit for bit: node array (DFS allocation order, relative u16 child
pointers, valid / leaf masks), leaf records (n + cluster indices in ascending pointer order) and the 32^3 voxel blocks."""
import ctypes as C
import os

import numpy as np
import pytest

from tg_b200 import ctypes_defs as T
from tg_b200 import scenes
from tg_b200.raytracer import from_scene

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
BIG = (1 << 25, 1 << 15, 1 << 16)  # oracle capacities for scenes beyond the reference's 2048-leaf reserve


def oracle_svo(O, scene, capacities=None):
    view = O.SceneView.from_scene(scene, with_lut=False)
    svo = O.svo_create(view, capacities=capacities)
    arrays = O.svo_arrays(svo)
    O.svo_destroy(svo)
    return arrays


def gpu_svo(scene):
    rt = from_scene(scene)
    try:
        rt.svo_update(force_full=True)
        rt.synchronize()
        svo, nodes, leaf, vox = rt.svo_download()
        rt.svo_free(svo)
        return nodes, leaf, vox, rt.timings()
    finally:
        rt.destroy()


def compare(got, want, what=""):
    gn, gl, gv = got[:3]
    wn, wl, wv = want
    assert gn.shape == wn.shape and np.array_equal(gn, wn), f"{what}: node arrays differ ({gn.shape} vs {wn.shape}); first bad {np.argwhere(gn[:min(len(gn), len(wn))] != wn[:min(len(gn), len(wn))])[:4].ravel()}"
    assert gl.shape == wl.shape and np.array_equal(gl, wl), f"{what}: leaf records differ: rows {np.unique(np.argwhere(gl != wl)[:, 0])[:8]}"
    assert gv.shape == wv.shape and np.array_equal(gv, wv), f"{what}: {int((gv != wv).sum())} of {gv.size} voxel words differ"


def test_small_grid_bit_exact(gpu, oracle):
    s = scenes.small_grid()
    got = gpu_svo(s)
    compare(got, oracle_svo(oracle, s), "small_grid")
    assert got[0].size > 1 and got[2].any()


def test_config1_object_bit_exact(gpu, oracle):
    """BASELINE configs[0]'s object: 16^3 clusters rotated 15 degrees about +Y, both densities."""
    for k in (3, 1):
        s = scenes.config1(k=k)
        compare(gpu_svo(s), oracle_svo(oracle, s), f"config1 k={k}")


@pytest.mark.parametrize("name", ["config1_k3_320x180", "config1_k1_320x180", "small_grid3_320x180"])
def test_against_committed_golden_fixtures(gpu, name):
    from tests.golden.make_golden import CASES
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    nodes, leaf, vox, _ = gpu_svo(CASES[name]())
    assert np.array_equal(nodes, g["svo_nodes"])
    assert np.array_equal(leaf, g["svo_leaf"])
    nz = np.nonzero(vox)[0].astype(np.uint32)
    assert np.array_equal(nz, g["svo_voxels_nonzero_idx"]) and np.array_equal(vox[nz], g["svo_voxels_nonzero"])


@pytest.mark.parametrize("name", ["small_grid", "grid4_tall", "config1_small", "dense_k1", "sparse_k5"])
def test_against_reference_made_fixtures(gpu, name):
    """K2 against arrays THE REFERENCE'S OWN tg_svo_create produced (tests/golden/make_reference_golden.py ran the reference's
    graphics/tg_sparse_voxel_octree.c, compiled from /root/reference, on these scenes)."""
    from tests.golden.make_reference_golden import SVO_CASES
    g = np.load(os.path.join(GOLDEN, f"ref_svo_{name}.npz"))
    nodes, leaf, vox, _ = gpu_svo(SVO_CASES[name]())
    assert np.array_equal(nodes, g["nodes"]), "node array differs from the reference's"
    assert np.array_equal(leaf, g["leaf"]), "leaf records differ from the reference's"
    nz = np.nonzero(vox)[0].astype(np.uint32)
    assert vox.size == int(g["n_voxel_words"]) and np.array_equal(nz, g["voxels_nonzero_idx"]) and np.array_equal(vox[nz], g["voxels_nonzero"])


def test_known_answer_single_axis_aligned_cluster(gpu):
    """SURVEY 8c KAT: one solid 8^3 cluster, axis aligned, centred at (20,20,20): blocks [0,32)^3 only, exactly 512 bits."""
    s = scenes.config1(k=1, dims=(1, 1, 1), width=64, height=36)
    s.objects[0].angle = 0.0
    s.objects[0].center = (20.0, 20.0, 20.0)
    s.objects[0].bits = np.full((1, 16), 0xFFFFFFFF, dtype=np.uint32)
    nodes, leaf, vox, _ = gpu_svo(s)
    # root -> 5 inner levels, one child each (octant 7 first: +x+y+z of the origin-centred box), then the leaf
    assert nodes.size == 6 and leaf.shape[0] == 1 and vox.size == 1024
    assert int(np.unpackbits(vox.view(np.uint8)).sum()) == 512
    assert leaf[0, 0] == 1 and leaf[0, 1] == 0
    assert (nodes[0] >> 16) & 0xFF == 0x80 and nodes[0] & 0xFFFF == 1            # child 7, first child right behind the root
    for lvl in (1, 2, 3, 4):
        assert (nodes[lvl] >> 16) & 0xFF == 0x01 and nodes[lvl] & 0xFFFF == 1     # then octant 0 all the way down
    assert nodes[4] >> 24 == 0x01 and nodes[3] >> 24 == 0                          # only level-4 nodes carry leaf bits
    assert nodes[5] == 0                                                           # leaf: data_pointer 0
    z, y, x = np.unravel_index(np.nonzero(np.unpackbits(vox.view(np.uint8), bitorder="little"))[0], (32, 32, 32))
    assert (x.min(), x.max(), y.min(), y.max(), z.min(), z.max()) == (16, 23, 16, 23, 16, 23)


def test_edge_cases_bit_exact(gpu, oracle):
    cases = []
    # object straddling the centre of the box (all 8 root children valid), arbitrary axis
    s = scenes.config1(k=3, dims=(5, 3, 4), width=64, height=36)
    s.objects[0].axis = (0.6, 0.0, 0.8)
    s.objects[0].angle = 1.234
    s.objects[0].center = (3.0, -2.0, 1.0)
    cases.append(s)
    # object partly outside the +-512 box, and one entirely outside (root stays empty)
    s = scenes.config1(k=1, dims=(6, 2, 6), width=64, height=36)
    s.objects[0].center = (500.0, -505.0, 0.0)
    cases.append(s)
    s = scenes.config1(k=1, dims=(2, 2, 2), width=64, height=36)
    s.objects[0].center = (900.0, 0.0, 0.0)
    cases.append(s)
    # empty masks only (no candidates: :515-531), and a mix of empty / single-voxel clusters
    s = scenes.config1(k=3, dims=(2, 2, 2), width=64, height=36)
    s.objects[0].bits = np.zeros_like(s.objects[0].bits)
    cases.append(s)
    s = scenes.config1(k=3, dims=(3, 2, 2), width=64, height=36)
    s.objects[0].bits = np.zeros_like(s.objects[0].bits)
    s.objects[0].bits[7, 3] = 1 << 21
    s.objects[0].bits[2, 15] = 1 << 31
    cases.append(s)
    # overlapping twins: two objects in the same space (leaf lists interleave by pointer order)
    s = scenes.config1(k=3, dims=(3, 3, 3), width=64, height=36)
    s.objects.append(scenes.ObjectSpec(center=s.objects[0].center, extent=s.objects[0].extent, angle=0.3, bits=s.objects[0].bits.copy()))
    cases.append(s)
    # dense fully solid rotated object: leaves with more than 64 contributing clusters keep the first 64
    s = scenes.config1(k=1, dims=(8, 8, 8), width=64, height=36)
    s.objects[0].bits = np.full_like(s.objects[0].bits, 0xFFFFFFFF)
    s.objects[0].center = (16.0, 16.0, 16.0)
    cases.append(s)
    for i, s in enumerate(cases):
        compare(gpu_svo(s), oracle_svo(oracle, s), f"edge case {i}")


def test_config2_box_region_bit_exact(gpu, oracle):
    """BASELINE configs[2]'s SVO: the +-512 box around the player inside the 1,024-object scene. The oracle builds from
    the objects that can touch the box (the others fail the SAT against every root child); the GPU sees all 2^21 clusters."""
    full = scenes.config2()
    rt = from_scene(full)
    try:
        rt.svo_update(force_full=True)
        rt.synchronize()
        svo, nodes, leaf, vox = rt.svo_download()
        rt.svo_free(svo)
        t = rt.timings()
    finally:
        rt.destroy()
    # oracle on the sub-scene of nearby objects, with pointers / cluster indices remapped to the full scene's
    near = [i for i, o in enumerate(full.objects) if max(abs(o.center[0]), abs(o.center[2])) < 512 + 160]
    sub = scenes.SceneSpec(name="near", width=64, height=36, camera=full.camera, objects=[full.objects[i] for i in near])
    wn, wl, wv = oracle_svo(oracle, sub, BIG)
    assert np.array_equal(nodes, wn)
    assert np.array_equal(vox, wv)
    # leaf records hold cluster INDICES (== pointers in a fresh scene): remap sub-scene indices to full-scene ones
    n_per = full.objects[0].n_clusters
    remap = np.concatenate([np.arange(n_per, dtype=np.uint32) + np.uint32(i * n_per) for i in near])
    wl2 = wl.copy()
    for r in range(wl.shape[0]):
        n = wl[r, 0]
        wl2[r, 1:1 + n] = remap[wl[r, 1:1 + n]]
    assert np.array_equal(leaf, wl2)
    assert leaf.shape[0] > 100 and t["svo_ms"] > 0


def test_tg_svo_create_entry_point(gpu, oracle):
    """The reference's own entry point (tg_sparse_voxel_octree.h:51): tg_svo_create(min, max, &scene, &svo) -> host arrays."""
    import tg_b200
    s = scenes.small_grid()
    rt = from_scene(s)
    try:
        L = tg_b200.lib()
        svo = T.tg_svo()
        L.tg_svo_create(T.v3(-512, -512, -512), T.v3(512, 512, 512), C.byref(rt._rt.scene), C.byref(svo))
        tg_b200.check()
        assert svo.node_buffer_capacity >= 1 << 14 and svo.leaf_node_data_buffer_capacity >= 1 << 13
        nodes = np.ctypeslib.as_array(svo.p_node_buffer, shape=(svo.node_buffer_count,)).copy()
        wn, wl, wv = oracle_svo(oracle, s)
        assert np.array_equal(nodes, wn) and svo.leaf_node_data_buffer_count == wl.shape[0] and svo.voxel_buffer_count_in_u32 == wv.size
        L.tg_svo_destroy(C.byref(svo))
        # a box that is not 1024^3 is refused (tg_sparse_voxel_octree.c:468-472)
        L.tg_svo_create(T.v3(0, 0, 0), T.v3(512, 512, 512), C.byref(rt._rt.scene), C.byref(svo))
        assert b"1024" in L.tgb200_last_error()
        L.tgb200_clear_error()
    finally:
        rt.destroy()


def _moved(scene, frame, which):
    """BASELINE configs[3] motion rule: translation.x += 0.5 * frame, angle += 1 degree * frame."""
    import copy
    s = copy.copy(scene)
    s.objects = list(scene.objects)
    for i in which:
        o = copy.copy(scene.objects[i])
        o.center = (o.center[0] + 0.5 * frame, o.center[1], o.center[2])
        o.angle = float(np.float32(o.angle) + scenes.deg2rad(1.0) * np.float32(frame))
        s.objects[i] = o
    return s


def test_incremental_update_equals_full_rebuild_and_oracle(gpu, oracle):
    """Config 4 in miniature: objects move every frame; the incremental update must leave exactly the arrays a full
    rebuild (and the oracle) produce, while re-sampling only the leaves the moved objects touch."""
    base = scenes.small_grid(grid=4, dims=(4, 2, 4))
    movers = [0, 5, 10]
    rt = from_scene(base, max_n_objects=len(base.objects) + 1, max_n_clusters=base.n_clusters + 1)
    try:
        rt.svo_update(force_full=True)
        rt.synchronize()
        n_full = rt.svo_leaves_resampled()
        for frame in range(1, 6):
            cur = _moved(base, frame, movers)
            for i in movers:
                rt.set_object_transform(i, cur.objects[i].center, cur.objects[i].angle)
            rt.svo_update()                      # incremental: only transforms changed
            rt.synchronize()
            resampled = rt.svo_leaves_resampled()
            svo, n1, l1, v1 = rt.svo_download(); rt.svo_free(svo)
            rt.svo_update(force_full=True)
            rt.synchronize()
            svo, n2, l2, v2 = rt.svo_download(); rt.svo_free(svo)
            compare((n1, l1, v1), (n2, l2, v2), f"frame {frame}: incremental vs full rebuild")
            assert 0 < resampled < len(l2), (resampled, len(l2))
            if frame in (1, 5):
                compare((n1, l1, v1), oracle_svo(oracle, cur), f"frame {frame}: incremental vs oracle")
        assert n_full > 0
        # an update with nothing moved re-samples nothing and changes nothing
        rt.set_object_transform(1, base.objects[1].center, base.objects[1].angle)
        rt.svo_update(); rt.synchronize()
        svo, n3, l3, v3 = rt.svo_download(); rt.svo_free(svo)
        compare((n3, l3, v3), (n2, l2, v2), "no-op move")
        # creating an object afterwards forces a full rebuild (pointer table changed)
        rt.create_object_from_data((0.0, 40.0, 0.0), (8, 8, 8), 0.2, (0.0, 1.0, 0.0), np.full((1, 16), 0xFFFFFFFF, dtype=np.uint32))
        rt.svo_update(); rt.synchronize()
        assert rt.svo_leaves_resampled() >= len(l2)
    finally:
        rt.destroy()


def test_config4_dynamic_scene_full_size(gpu):
    """BASELINE configs[3]: the 10^9-voxel scene, 64 objects moving per frame. Objects 0..63 (the rule of SURVEY 8d) lie
    outside the +-512 box; the 64 objects nearest to the player are the ones that exercise the update."""
    full = scenes.config2()
    order = np.argsort([o.center[0] ** 2 + o.center[2] ** 2 for o in full.objects])
    for movers in (list(range(64)), [int(i) for i in order[:64]]):
        rt = from_scene(full)
        try:
            rt.svo_update(force_full=True); rt.synchronize()
            for frame in (1, 2, 3):
                cur = _moved(full, frame, movers)
                for i in movers:
                    rt.set_object_transform(i, cur.objects[i].center, cur.objects[i].angle)
                rt.svo_update(); rt.synchronize()
                t_inc = rt.timings()["svo_ms"]
                resampled = rt.svo_leaves_resampled()
                svo, n1, l1, v1 = rt.svo_download(); rt.svo_free(svo)
                rt.svo_update(force_full=True); rt.synchronize()
                t_full = rt.timings()["svo_ms"]
                svo, n2, l2, v2 = rt.svo_download(); rt.svo_free(svo)
                compare((n1, l1, v1), (n2, l2, v2), f"movers[0]={movers[0]} frame {frame}")
                assert resampled <= len(l2)
            print(f"config4 movers[0]={movers[0]}: {resampled}/{len(l2)} leaves re-sampled, incremental {t_inc:.3f} ms vs full {t_full:.3f} ms")
        finally:
            rt.destroy()


def test_config4_full_size_incremental_update_against_the_oracle(gpu, oracle):
    """BASELINE configs[3] at full size, against the ORACLE (the test above only compares the incremental update with the GPU's own full
    rebuild): the 10^9-voxel scene, the 64 objects nearest to the player moved for three frames, updated incrementally; the three SVO
    arrays must equal the oracle's from-scratch tg_svo_create of the moved scene, and the frame rendered with that tree (visibility +
    GI radiance, every 90th scanline) the oracle's frame. The oracle only gets the objects that can touch the +-512 box."""
    full = scenes.config2()
    order = np.argsort([o.center[0] ** 2 + o.center[2] ** 2 for o in full.objects], kind="stable")
    movers = [int(i) for i in order[:64]]
    rt = from_scene(full)
    try:
        rt.set_gi(True, 1)
        rt.svo_update(force_full=True); rt.synchronize()
        for frame in (1, 2, 3):
            cur = _moved(full, frame, movers)
            for i in movers:
                rt.set_object_transform(i, cur.objects[i].center, cur.objects[i].angle)
            rt.clear(); rt.render(); rt.synchronize()   # render() updates the SVO incrementally before shading
        assert 0 < rt.svo_leaves_resampled()
        svo, n1, l1, v1 = rt.svo_download(); rt.svo_free(svo)
        vis, rad = rt.read_visibility(), rt.read_radiance()
    finally:
        rt.destroy()
    in_box = [o for o in cur.objects if max(abs(o.center[0]), abs(o.center[2])) < 512 + 160 + 2]
    box_scene = scenes.SceneSpec(name="c4_box", width=cur.width, height=cur.height, camera=cur.camera, objects=in_box)
    osvo = oracle.svo_create(oracle.SceneView.from_scene(box_scene, with_lut=False), capacities=(1 << 25, 1 << 15, 1 << 16))
    try:
        wn, wl, wv = oracle.svo_arrays(osvo)
        # the oracle numbers the clusters of its sub-scene from 0: leaf records hold cluster indices -> compare nodes and voxels bit for bit, and the
        # leaf records through the object-order-preserving pointer map (every object has 2,048 clusters)
        assert np.array_equal(n1, wn), "node arrays differ"
        assert np.array_equal(v1, wv), f"{int((v1 != wv).sum())} voxel words differ"
        keep = np.asarray([i for i, o in enumerate(cur.objects) if any(o is b for b in in_box)], dtype=np.uint32)
        gl, ol = l1.view(np.uint32).reshape(-1, 65), wl.view(np.uint32).reshape(-1, 65)
        assert np.array_equal(gl[:, 0], ol[:, 0]), "per-leaf cluster counts differ"
        n = np.minimum(ol[:, 0], 64)
        mask = np.arange(64)[None, :] < n[:, None]
        mapped = keep[(ol[:, 1:] // 2048) % len(keep)] * 2048 + ol[:, 1:] % 2048
        assert np.array_equal(gl[:, 1:][mask], mapped[mask]), "leaf cluster lists differ"
        rows = np.arange(13, cur.height, 90)
        rays = oracle.camera_rays(oracle.camera_from_spec(cur.camera))
        view = oracle.SceneView.from_scene(cur, with_lut=True)
        want_vis, _ = oracle.visibility(view, rays, cur.width, cur.height, oracle.VIS_SCREEN_RECT, 13, cur.height, 90)
        assert np.array_equal(vis[rows], want_vis[rows]), f"{int((vis[rows] != want_vis[rows]).sum())} visibility words differ"
        want_rad = np.zeros((cur.height, cur.width, 4), dtype=np.float32)
        oracle.shade(view, rays, cur.width, cur.height, want_vis, osvo, gi=True, frame_seed=1, y0=13, y1=cur.height, ystep=90, out=want_rad)
        bad = ~np.isclose(rad[rows], want_rad[rows], rtol=1e-3, atol=1e-6)
        assert not bad.any(), f"{int(bad.any(axis=-1).sum())} pixels beyond 1e-3"
    finally:
        oracle.svo_destroy(osvo)

"""GPU: K1 (tg_b200/csrc/tgb_visibility.cu) through the C ABI against the oracle, bit for bit.

The visibility buffer is integer work: the bar is bit-exact (np.array_equal on the u64 words), at sizes the oracle
finishes in seconds, against the committed golden fixtures, and at BASELINE config 2's full size through a scanline
subset plus size-independent properties (idempotence of re-rendering, clear value, pointer-base linearity)."""
import os

import numpy as np
import pytest

from tg_b200 import scenes
from tests.helpers import CLEAR, describe_mismatch, gpu_visibility, oracle_visibility

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def check(O, scene, mode=None):
    got, timings = gpu_visibility(scene)
    want = oracle_visibility(O, scene, mode)
    assert np.array_equal(got, want), describe_mismatch(got, want)
    return got, timings


@pytest.mark.parametrize("k", [1, 3])
def test_config1_full_size_bit_exact(gpu, oracle, k):
    """BASELINE configs[0]: one object of 16^3 clusters, 1280x720."""
    got, _ = check(oracle, scenes.config1(k=k))
    assert (got != CLEAR).mean() > 0.05


def test_config1_against_brute_force_oracle(gpu, oracle):
    check(oracle, scenes.config1(k=3, width=320, height=180), oracle.VIS_BRUTE_FORCE)


@pytest.mark.parametrize("name", ["config1_k3_320x180", "config1_k1_320x180", "small_grid3_320x180"])
def test_against_committed_golden_fixtures(gpu, name):
    from tests.golden.make_golden import CASES
    want = np.load(os.path.join(GOLDEN, name + ".npz"))["vis"]
    got, _ = gpu_visibility(CASES[name]())
    assert np.array_equal(got, want), describe_mismatch(got, want)


@pytest.mark.parametrize("make", [lambda: scenes.config1(k=3), lambda: scenes.config1(k=1, width=640, height=360), lambda: scenes.small_grid(),
                                  lambda: scenes.small_grid(grid=5, width=333, height=177, dims=(3, 5, 2))])
def test_masks_staged_in_shared_memory_variant_is_bit_exact(gpu, oracle, make):
    """TGB_K1_STAGE_MASKS=1 (cp.async copies of the cluster masks into shared memory, one per group of lanes that march the same cluster;
    north_star (a) as written, not the default): pure data movement, the words must be those of the default kernel and of the oracle."""
    s = make()
    want = oracle_visibility(oracle, s)
    os.environ["TGB_K1_STAGE_MASKS"] = "1"
    try:
        got, _ = gpu_visibility(s)
    finally:
        os.environ.pop("TGB_K1_STAGE_MASKS", None)
    assert np.array_equal(got, want), describe_mismatch(got, want)


def test_rotated_objects_grid_brute_force(gpu, oracle):
    check(oracle, scenes.small_grid(), oracle.VIS_BRUTE_FORCE)
    check(oracle, scenes.small_grid(grid=5, width=333, height=177, dims=(3, 5, 2)), oracle.VIS_BRUTE_FORCE)  # ragged resolution


def test_edge_cases(gpu, oracle):
    base = dict(width=192, height=108)
    cases = []
    # camera inside the object, looking around
    for pos, pitch, yaw in [((0.0, 0.0, 0.0), 0.0, 0.0), ((3.3, -2.1, 7.7), -0.9, 2.0), ((0.0, 40.0, 0.0), -1.5, 0.0), ((60.0, 0.5, 0.5), 0.0, 1.5707964)]:
        s = scenes.config1(k=3, dims=(6, 4, 6), **base)
        s.camera = scenes.CameraSpec(pos, pitch, yaw, 0.0, aspect=192 / 108)
        cases.append(s)
    # axis-aligned object (exact zeros in the rotation) and a grazing, axis-parallel view
    s = scenes.config1(k=1, dims=(4, 4, 4), **base)
    s.objects[0].angle = 0.0
    s.camera = scenes.CameraSpec((0.0, 16.0, 80.0), 0.0, 0.0, 0.0, aspect=192 / 108)  # eye exactly in the top face plane
    cases.append(s)
    # single cluster, dense
    s = scenes.config1(k=1, dims=(1, 1, 1), **base)
    s.camera = scenes.CameraSpec((0.0, 0.0, 20.0), 0.0, 0.0, 0.0, aspect=192 / 108)
    cases.append(s)
    # far plane cuts through the object (d <= 1 test) and object entirely beyond it
    s = scenes.config1(k=3, dims=(8, 8, 8), **base)
    s.camera = scenes.CameraSpec((0.0, 0.0, 160.0), 0.0, 0.0, 0.0, aspect=192 / 108, far=150.0)
    cases.append(s)
    s = scenes.config1(k=3, dims=(4, 4, 4), **base)
    s.camera = scenes.CameraSpec((0.0, 0.0, 160.0), 0.0, 0.0, 0.0, aspect=192 / 108, far=100.0)
    cases.append(s)
    # object behind the camera
    s = scenes.config1(k=3, dims=(4, 4, 4), **base)
    s.camera = scenes.CameraSpec((0.0, 0.0, -160.0), 0.0, 0.0, 0.0, aspect=192 / 108)
    cases.append(s)
    # empty masks, and one with a single voxel
    s = scenes.config1(k=3, dims=(2, 2, 2), **base)
    s.objects[0].bits = np.zeros_like(s.objects[0].bits)
    cases.append(s)
    s = scenes.config1(k=3, dims=(2, 2, 2), **base)
    s.objects[0].bits = np.zeros_like(s.objects[0].bits)
    s.objects[0].bits[5, 9] = 1 << 13
    s.camera = scenes.CameraSpec((2.0, 3.0, 40.0), 0.0, 0.0, 0.0, aspect=192 / 108)
    cases.append(s)
    # large translations (float precision of the hoisted chain), arbitrary axis
    s = scenes.config1(k=3, dims=(5, 3, 4), **base)
    s.objects[0].center = (2900.0, -1300.0, 2100.0)
    s.objects[0].axis = (0.6, 0.0, 0.8)
    s.objects[0].angle = 1.234
    s.camera = scenes.CameraSpec((2900.0, -1290.0, 2200.0), -0.1, 0.0, 0.0, aspect=192 / 108)
    cases.append(s)
    for i, s in enumerate(cases):
        got, want = gpu_visibility(s)[0], oracle_visibility(oracle, s, oracle.VIS_BRUTE_FORCE)
        assert np.array_equal(got, want), f"edge case {i}: " + describe_mismatch(got, want)


def test_overlapping_objects_and_ties(gpu, oracle):
    """Two objects occupying the same space with identical voxels: equal depths, the lower cluster pointer must win."""
    a = scenes.config1(k=3, width=192, height=108, dims=(3, 3, 3))
    twin = scenes.ObjectSpec(center=a.objects[0].center, extent=a.objects[0].extent, angle=a.objects[0].angle, bits=a.objects[0].bits.copy())
    a.objects.append(twin)
    a.camera = scenes.CameraSpec((0.0, 5.0, 60.0), -0.1, 0.0, 0.0, aspect=192 / 108)
    got, _ = check(oracle, a, oracle.VIS_BRUTE_FORCE)
    hit = got != CLEAR
    assert hit.any() and (((got[hit] >> np.uint64(9)) & np.uint64(0x7FFFFFFF)) < np.uint64(27)).all()


def test_render_is_idempotent_and_clear_resets(gpu):
    from tg_b200.raytracer import from_scene
    s = scenes.small_grid()
    rt = from_scene(s)
    try:
        rt.clear(); rt.render_visibility(); rt.synchronize()
        a = rt.read_visibility()
        rt.render_visibility(); rt.synchronize()  # atomicMin onto an already resolved buffer: unchanged
        assert np.array_equal(a, rt.read_visibility())
        rt.clear(); rt.synchronize()
        assert (rt.read_visibility() == CLEAR).all()
        hit, depth, cluster, voxel = rt.get_hovered_voxel(0, 0)
        assert not hit and cluster == 0xFFFFFFFF and voxel == 0xFFFFFFFF
        rt.render_visibility(); rt.synchronize()
        ys, xs = np.nonzero(a != CLEAR)
        y, x = int(ys[len(ys) // 2]), int(xs[len(xs) // 2])
        hit, depth, cluster, voxel = rt.get_hovered_voxel(x, y)  # tgvk_raytracer.c:1605-1655
        w = int(a[y, x])
        assert hit and cluster == (w >> 9) & 0x7FFFFFFF and voxel == w & 511
        assert depth == np.float32(w >> 40) / np.float32(16777215.0)
    finally:
        rt.destroy()


def test_pointer_base_linearity(gpu):
    """Multi-GPU contract: a shard's words carry local pointer + base, nothing else changes."""
    s = scenes.small_grid()
    a, _ = gpu_visibility(s, base=0)
    b, _ = gpu_visibility(s, base=1 << 20)
    hit = a != CLEAR
    assert np.array_equal(hit, b != CLEAR)
    assert ((b[hit] - a[hit]) == np.uint64((1 << 20) << 9)).all()


def test_destroy_object_then_render(gpu, oracle):
    """Pointer-table compaction (tgvk_raytracer.c:1009-1066) mirrored on the device: cluster idx != pointer afterwards."""
    from tg_b200.raytracer import from_scene
    s = scenes.small_grid()
    rt = from_scene(s)
    try:
        rt.destroy_object(3)
        rt.destroy_object(0)
        rt.clear(); rt.render_visibility(); rt.synchronize()
        got = rt.read_visibility()
        sc = rt.scene
        n = sc.n_cluster_pointers
        cap = sc.cluster_pointer_capacity
        pointers = np.ctypeslib.as_array(sc.p_cluster_pointers, shape=(cap,))[:n].copy()
        assert not np.array_equal(pointers, np.arange(n))
        c2o = np.ctypeslib.as_array(sc.p_cluster_idx_to_object_idx, shape=(cap,)).copy()
        masks = np.ctypeslib.as_array(sc.p_voxel_cluster_data, shape=(cap, 16)).copy()
        from tg_b200 import ctypes_defs as T
        import ctypes
        objects = np.frombuffer(ctypes.string_at(sc.p_objects, sc.object_capacity * 44), dtype=T.VOXEL_OBJECT_DTYPE).copy()
        view = oracle.SceneView(objects, pointers, c2o, masks)
        rays = oracle.camera_rays(oracle.camera_from_spec(s.camera))
        want, _ = oracle.visibility(view, rays, s.width, s.height, oracle.VIS_BRUTE_FORCE)
        assert np.array_equal(got, want), describe_mismatch(got, want)
    finally:
        rt.destroy()


def test_config2_full_size_scanline_subset_and_properties(gpu, oracle):
    """BASELINE configs[1]: 1,024 objects, 2^21 clusters, 3840x2160. The oracle evaluates every 40th scanline."""
    s = scenes.config2()
    from tg_b200.raytracer import from_scene
    rt = from_scene(s)
    try:
        rt.clear(); rt.render_visibility(); rt.synchronize()
        got = rt.read_visibility()
        t = rt.timings()
        want = oracle_visibility(oracle, s, None, 7, None, 40)
        rows = np.arange(7, s.height, 40)
        assert np.array_equal(got[rows], want[rows]), describe_mismatch(got[rows], want[rows])
        hit = got != CLEAR
        assert 0.3 < hit.mean() < 0.9
        # properties: pointers in range, voxel bit really set, depth < 2^24
        ptr = ((got[hit] >> np.uint64(9)) & np.uint64(0x7FFFFFFF)).astype(np.int64)
        vox = (got[hit] & np.uint64(511)).astype(np.int64)
        assert ptr.max() < s.n_clusters
        masks = np.concatenate([o.bits for o in s.objects])
        assert ((masks[ptr, vox // 32] >> (vox % 32).astype(np.uint32)) & 1).all()
        # idempotence at full size
        rt.render_visibility(); rt.synchronize()
        assert np.array_equal(got, rt.read_visibility())
        assert 0 < t["n_visible_objects"] < 400
    finally:
        rt.destroy()


def test_config5_one_shard_full_size_scanline_subset(gpu, oracle):
    """BASELINE configs[4], one GPU's share: 12,288 objects = 25,165,824 clusters = 1.29e10 voxels generated on the device
    (tg_raytracer_create_object_synthetic), 3840x2160, visibility + SVO + GI. The oracle evaluates every 60th scanline over the
    objects that can write a pixel (those within the far plane; the others fail depth <= 1, visibility.frag:194) with host-made
    masks of the same seeds; its cluster pointers are mapped into the full scene's (every object has 2,048 clusters and the
    subset keeps the object order, so ties resolve identically)."""
    from tg_b200.raytracer import from_scene
    s = scenes.config5_shard(0, 1)
    assert len(s.objects) == 12288 and s.n_clusters == 25165824
    rt = from_scene(s)
    try:
        rt.set_gi(True, 1)
        rt.clear(); rt.render(); rt.synchronize()
        got = rt.read_visibility()
        rad = rt.read_radiance()
        t = rt.timings()
        hit = got != CLEAR
        assert 0.3 < hit.mean() < 0.9 and 0 < t["n_visible_objects"] < 400
        ptr = ((got[hit] >> np.uint64(9)) & np.uint64(0x7FFFFFFF)).astype(np.int64)
        assert ptr.max() < s.n_clusters
        # the voxel every word names is solid in the CPU mirror the library keeps (scene.p_voxel_cluster_data)
        vox = (got[hit] & np.uint64(511)).astype(np.int64)
        mirror = np.ctypeslib.as_array(rt.scene.p_voxel_cluster_data, shape=(s.n_clusters * 16,))
        assert ((mirror[ptr * 16 + vox // 32] >> (vox % 32).astype(np.uint32)) & 1).all()
    finally:
        rt.destroy()
    cam, far = s.camera.position, s.camera.far
    keep = [i for i, o in enumerate(s.objects) if ((o.center[0] - cam[0]) ** 2 + (o.center[2] - cam[2]) ** 2) ** 0.5 < far + 160.0]
    assert 20 < len(keep) < 400
    sub = scenes.SceneSpec(name="c5_near", width=s.width, height=s.height, camera=s.camera, objects=[s.objects[i] for i in keep])
    for o in sub.objects:
        o.bits = scenes.random_solid_bits(o.seed, o.n_clusters, o.k)
    rows = np.arange(11, s.height, 60)
    want = oracle_visibility(oracle, sub, None, 11, None, 60)[rows]
    w_hit = want != CLEAR
    sub_ptr = (want >> np.uint64(9)) & np.uint64(0x7FFFFFFF)
    full_ptr = np.asarray(keep, dtype=np.uint64)[(sub_ptr // np.uint64(2048)).astype(np.int64) % len(keep)] * np.uint64(2048) + sub_ptr % np.uint64(2048)
    remapped = np.where(w_hit, (want & ~(np.uint64(0x7FFFFFFF) << np.uint64(9))) | (full_ptr << np.uint64(9)), want)
    assert np.array_equal(got[rows], remapped), describe_mismatch(got[rows], remapped)
    # GI + shading of those rows: the oracle shades its own (subset-pointer) buffer with its own SVO of the +-512 box
    rays = oracle.camera_rays(oracle.camera_from_spec(s.camera))
    view = oracle.SceneView.from_scene(sub, with_lut=True)
    in_box = [o for o in sub.objects if max(abs(o.center[0]), abs(o.center[2])) < 512 + 160]
    svo = oracle.svo_create(oracle.SceneView.from_scene(scenes.SceneSpec(name="c5_box", width=s.width, height=s.height, camera=s.camera, objects=in_box), with_lut=False),
                            capacities=(1 << 25, 1 << 15, 1 << 16))
    want_vis = np.full((s.height, s.width), CLEAR, dtype=np.uint64)
    want_vis[rows] = want
    want_rad = np.zeros((s.height, s.width, 4), dtype=np.float32)
    oracle.shade(view, rays, s.width, s.height, want_vis, svo, gi=True, frame_seed=1, y0=11, y1=s.height, ystep=60, out=want_rad)
    oracle.svo_destroy(svo)
    assert np.allclose(rad[rows], want_rad[rows], rtol=1e-3, atol=1e-6), f"{int((~np.isclose(rad[rows], want_rad[rows], rtol=1e-3, atol=1e-6)).any(axis=-1).sum())} pixels beyond 1e-3"


def test_scene_dump_and_load_round_trip(gpu, tmp_path):
    """tgb200_scene_save / tgb200_scene_load: a scene with per-voxel materials, two LUTs and a destroyed object, written to disc and
    loaded into a fresh raytracer, renders the identical frame (visibility words, radiance) and writes the identical file again."""
    from tg_b200.raytracer import Raytracer, from_scene
    from tg_b200 import ctypes_defs as T
    s = scenes.small_grid(grid=3, dims=(3, 2, 4))
    for i, o in enumerate(s.objects):
        o.lut_indices = scenes.random_lut_indices(o.seed, o.n_clusters)
        o.lut_idx = i % 2
    s.n_luts = 2
    rt = from_scene(s)
    path, path2 = tmp_path / "scene.tgb", tmp_path / "scene2.tgb"
    try:
        for i in range(8):
            rt.color_lut_set(i, 0.1 * i, 1.0 - 0.1 * i, 0.5, lut_idx=1)
        rt.destroy_object(4)
        # a new object takes over the freed cluster indices (LIFO free-list, tgvk_raytracer.c:855: not one ascending run any more)
        bits = scenes.random_solid_bits(77, 3 * 2 * 4, 2)
        rt.create_object_from_data((20.0, 30.0, -10.0), (24, 16, 32), 0.4, (0.0, 1.0, 0.0), bits, scenes.random_lut_indices(77, 3 * 2 * 4), lut_idx=1)
        rt.set_gi(True, 3)
        rt.clear(); rt.render(); rt.synchronize()
        vis, rad = rt.read_visibility(), rt.read_radiance()
        assert rt.scene_save(path)
    finally:
        rt.destroy()
    cam = T.make_camera(s.camera.position, s.camera.pitch, s.camera.yaw, s.camera.roll, s.camera.fov_y_deg, s.camera.aspect, s.camera.near, s.camera.far)
    rt2 = Raytracer(cam, len(s.objects), s.n_clusters, s.width, s.height)
    try:
        assert rt2.scene_load(path)
        assert rt2.scene.n_objects == len(s.objects)
        rt2.set_gi(True, 3)
        rt2.clear(); rt2.render(); rt2.synchronize()
        vis2 = rt2.read_visibility()
        # the file lists objects in index order (the new object took index 4 but the LAST pointer range): pointers differ, depth and voxel fields are identical
        field = ~(np.uint64(0x7FFFFFFF) << np.uint64(9))
        assert np.array_equal(vis & field, vis2 & field)
        assert np.array_equal(rad, rt2.read_radiance())
        assert rt2.scene_save(path2)
        assert open(path, "rb").read() == open(path2, "rb").read()
    finally:
        rt2.destroy()
    with open(tmp_path / "bad.tgb", "wb") as f:
        f.write(open(path, "rb").read()[:1000])
    rt3 = Raytracer(cam, len(s.objects), s.n_clusters, s.width, s.height)
    try:
        with pytest.raises(Exception):
            rt3.scene_load(tmp_path / "bad.tgb")
    finally:
        import tg_b200
        tg_b200.lib().tgb200_clear_error()
        rt3.destroy()


def _objects_in_window_cone(scene, rays, x0, x1, y0, y1, radius):
    """Indices of the objects whose bounding sphere meets the cone around the window's rays -- a float64 statement of "can this object
    touch a ray of the window" that shares no code with the screen-rectangle pruning of either path."""
    def ray(px, py):
        fx, fy = (px + 0.5) / scene.width, 1.0 - (py + 0.5) / scene.height
        c = lambda v: np.array([v.x, v.y, v.z], dtype=np.float64)
        bl, br, tr, tl = c(rays.ray_bl), c(rays.ray_br), c(rays.ray_tr), c(rays.ray_tl)
        left, right = bl * (1 - fy) + tl * fy, br * (1 - fy) + tr * fy
        d = left * (1 - fx) + right * fx
        return d / np.linalg.norm(d)
    axis = ray(0.5 * (x0 + x1 - 1), 0.5 * (y0 + y1 - 1))
    half = max(np.arccos(np.clip(np.dot(axis, ray(px, py)), -1, 1)) for px in (x0, x1 - 1) for py in (y0, y1 - 1)) + 1e-3
    cam = np.array([rays.camera.x, rays.camera.y, rays.camera.z], dtype=np.float64)
    keep = []
    for i, o in enumerate(scene.objects):
        v = np.asarray(o.center, dtype=np.float64) - cam
        dist = np.linalg.norm(v)
        if dist - radius > scene.camera.far * 1.001:
            continue
        if dist <= radius or np.arccos(np.clip(np.dot(v / dist, axis), -1, 1)) <= half + np.arcsin(min(1.0, radius / dist)):
            keep.append(i)
    return keep


@pytest.mark.parametrize("window", [(1890, 1938, 640, 688), (300, 348, 1700, 1748), (600, 648, 600, 648), (2300, 2348, 500, 548), (1912, 1960, 520, 568)])
def test_config2_full_size_windows_against_unpruned_brute_force(gpu, oracle, window):
    """BASELINE configs[1] at 3840x2160: five 48x48-pixel windows (near objects, object silhouettes near the horizon, two objects overlapping in depth) of the frame against the oracle's UNPRUNED brute force (every pixel x
    every cluster of every object that can touch the window's cone of rays, chosen by a float64 sphere-vs-cone test written here).
    The scanline test above leans on the oracle's screen-rectangle pruning, which is the same design as k_cull_objects; this one does
    not share it."""
    from tg_b200.raytracer import from_scene
    x0, x1, y0, y1 = window
    s = scenes.grid_scene("config2_devicebits", 32, 32, 3840, 2160, k=3, with_bits=False)
    rays = oracle.camera_rays(oracle.camera_from_spec(s.camera))
    rt = from_scene(s)
    try:
        rt.clear(); rt.render_visibility(); rt.synchronize()
        got = rt.read_visibility()[y0:y1, x0:x1]
    finally:
        rt.destroy()
    keep = _objects_in_window_cone(s, rays, x0, x1, y0, y1, radius=0.5 * float(np.linalg.norm(s.objects[0].extent)) + 1.0)
    assert 1 <= len(keep) <= 40, len(keep)
    sub = scenes.SceneSpec(name="c2_window", width=s.width, height=s.height, camera=s.camera, objects=[s.objects[i] for i in keep])
    for o in sub.objects:
        o.bits = scenes.random_solid_bits(o.seed, o.n_clusters, o.k)
    view = oracle.SceneView.from_scene(sub, with_lut=False)
    want, n_fragments = oracle.visibility_window(view, rays, s.width, s.height, x0, x1, y0, y1)
    want = want[y0:y1, x0:x1]
    assert n_fragments == 48 * 48 * sub.n_clusters
    # the sub-scene numbers its clusters from 0: map its pointers to the full scene's (every object has 2,048 clusters, order kept)
    w_hit = want != CLEAR
    sub_ptr = (want >> np.uint64(9)) & np.uint64(0x7FFFFFFF)
    full_ptr = np.asarray(keep, dtype=np.uint64)[(sub_ptr // np.uint64(2048)).astype(np.int64) % len(keep)] * np.uint64(2048) + sub_ptr % np.uint64(2048)
    remapped = np.where(w_hit, (want & ~(np.uint64(0x7FFFFFFF) << np.uint64(9))) | (full_ptr << np.uint64(9)), want)
    assert np.array_equal(got, remapped), describe_mismatch(got, remapped)
    assert w_hit.sum() > 100


def test_dense_view_far_plane_4000_scanline_subset(gpu, oracle):
    """The dense-view stress of bench.py (`c2far`): the configs[1] scene with the far plane at 4000, so that several hundred objects
    survive the cull, the front-to-back sort and the per-tile object windows carry real load, and distant objects project to a few
    pixels. Every 120th scanline against the oracle; visibility + GI radiance."""
    from tg_b200.raytracer import from_scene
    s = scenes.config2()
    s.camera.far = 4000.0
    rt = from_scene(s)
    try:
        rt.set_gi(True, 1)
        rt.clear(); rt.render(); rt.synchronize()
        got = rt.read_visibility()
        rad = rt.read_radiance()
        t = rt.timings()
    finally:
        rt.destroy()
    assert t["n_visible_objects"] > 250, t["n_visible_objects"]
    rows = np.arange(5, s.height, 120)
    want = oracle_visibility(oracle, s, None, 5, None, 120)
    assert np.array_equal(got[rows], want[rows]), describe_mismatch(got[rows], want[rows])
    rays = oracle.camera_rays(oracle.camera_from_spec(s.camera))
    view = oracle.SceneView.from_scene(s, with_lut=True)
    in_box = [o for o in s.objects if max(abs(o.center[0]), abs(o.center[2])) < 512 + 160]
    svo = oracle.svo_create(oracle.SceneView.from_scene(scenes.SceneSpec(name="far_box", width=s.width, height=s.height, camera=s.camera, objects=in_box), with_lut=False),
                            capacities=(1 << 25, 1 << 15, 1 << 16))
    want_rad = np.zeros((s.height, s.width, 4), dtype=np.float32)
    oracle.shade(view, rays, s.width, s.height, want, svo, gi=True, frame_seed=1, y0=5, y1=s.height, ystep=120, out=want_rad)
    oracle.svo_destroy(svo)
    assert np.allclose(rad[rows], want_rad[rows], rtol=1e-3, atol=1e-6), f"{int((~np.isclose(rad[rows], want_rad[rows], rtol=1e-3, atol=1e-6)).any(axis=-1).sum())} pixels beyond 1e-3"

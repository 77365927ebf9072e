"""CPU: the oracle's restatement against THE REFERENCE'S OWN CODE. oracle/Makefile compiles the reference's portable C
files from /root/reference/tg/src -- math/tg_math.c, physics/tg_physics.c, util/tg_amanatides_woo.c and
graphics/tg_sparse_voxel_octree.c (the CPU SVO builder and traversal) -- into oracle/_ref/libtg_ref.so; every test here
feeds the same inputs to a reference function and to its restatement in oracle/ and demands identical bits. This is
what pins the oracle (and through it the CUDA path) to reference-run outputs for the SVO build, the C traversal and
the whole math layer under the shader transcriptions. Skipped only where the library could not be built."""
import ctypes as C

import numpy as np
import pytest

from tg_b200 import ctypes_defs as T
from tg_b200 import scenes


@pytest.fixture(scope="module")
def ref(oracle):
    R = oracle.ref()
    if R is None:
        pytest.skip("oracle/_ref/libtg_ref.so was not built (no reference tree at build time)")
    return R


def bits(x):
    return bytes(x)


def rand_m4(rng, kind):
    m = T.m4()
    a = np.ctypeslib.as_array((T.f32 * 16).from_buffer(m))
    if kind == 0:
        a[:] = rng.normal(size=16).astype(np.float32)
    elif kind == 1:  # rigid-ish: what the path inverts (rotation and camera matrices)
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        full = np.eye(4, dtype=np.float32)
        full[:3, :3] = q.astype(np.float32)
        full[:3, 3] = rng.uniform(-500, 500, 3).astype(np.float32)
        a[:] = full.T.reshape(16)  # column-major
    else:
        a[:] = (rng.integers(-4, 5, 16)).astype(np.float32)
        a[0] += 9.0; a[5] += 9.0; a[10] += 9.0; a[15] += 9.0
    return m


def test_matrix_and_vector_routines_bit_exact(oracle, ref):
    """tgm_m4_mul / inverse / angle_axis / euler / perspective / translate / mulv4, tgm_v3_normalized / lerp
    (math/tg_math.c:1091-1098,1175-1184,1870-1910,2020-2028,2185-2234,2374-2438,2469-2499,2669-2694)."""
    L = oracle.lib()
    rng = np.random.default_rng(11)
    for trial in range(400):
        a, b = rand_m4(rng, trial % 3), rand_m4(rng, (trial + 1) % 3)
        assert bits(L.tgo_pin_m4_mul(a, b)) == bits(ref.tgm_m4_mul(a, b))
        assert bits(L.tgo_pin_m4_inverse(a)) == bits(ref.tgm_m4_inverse(a))
        v = T.v4(*rng.normal(size=4).astype(np.float32))
        assert bits(L.tgo_pin_m4_mulv4(a, v)) == bits(ref.tgm_m4_mulv4(a, v))
        angle = np.float32(rng.uniform(-7, 7))
        axis = rng.normal(size=3).astype(np.float32)
        axis = axis / np.linalg.norm(axis) if trial % 4 else np.array([0, 1, 0], dtype=np.float32)
        assert bits(L.tgo_pin_m4_angle_axis(angle, T.v3(*axis))) == bits(ref.tgm_m4_angle_axis(angle, T.v3(*axis)))
        e = rng.uniform(-3.2, 3.2, 3).astype(np.float32)
        assert bits(L.tgo_pin_m4_euler(*e)) == bits(ref.tgm_m4_euler(*e))
        fov, aspect = np.float32(rng.uniform(0.3, 2.5)), np.float32(rng.uniform(0.5, 3.0))
        near, far = np.float32(rng.uniform(0.01, 1.0)), np.float32(rng.uniform(10.0, 5000.0))
        assert bits(L.tgo_pin_m4_perspective(fov, aspect, near, far)) == bits(ref.tgm_m4_perspective(fov, aspect, near, far))
        t = T.v3(*rng.uniform(-600, 600, 3).astype(np.float32))
        assert bits(L.tgo_pin_m4_translate(t)) == bits(ref.tgm_m4_translate(t))
        p, q = T.v3(*rng.normal(size=3).astype(np.float32)), T.v3(*rng.normal(size=3).astype(np.float32))
        assert bits(L.tgo_pin_v3_normalized(p)) == bits(ref.tgm_v3_normalized(p))
        f = np.float32(rng.uniform(0, 1))
        assert bits(L.tgo_pin_v3_lerp(p, q, f)) == bits(ref.tgm_v3_lerp(p, q, f))


def test_camera_rays_from_reference_matrices(oracle, ref):
    """tgo_camera_rays (tgvk_core.c:382-444 restated) rebuilt here from the REFERENCE's matrix routines: same four corner rays."""
    for pos, pitch, yaw, roll, fov, aspect in [((65.14, -30.74, 73.03), -0.173, 0.710, 0.0, 70.0, 16 / 9), ((0, 200, 0), -0.5235988, 0.0, 0.0, 70.0, 16 / 9),
                                                ((3, 4, 5), 0.3, -2.0, 0.4, 55.0, 4 / 3)]:
        cam = T.make_camera(pos, pitch, yaw, roll, fov, aspect, 0.1, 1000.0)
        got = oracle.camera_rays(cam)
        # tg_camera_rotation = inverse(euler), tg_camera_projection = perspective, corner = normalize((inverse(p * r) * (sx, sy, 1, 1)).xyz)
        r = ref.tgm_m4_inverse(ref.tgm_m4_euler(np.float32(pitch), np.float32(yaw), np.float32(roll)))
        p = ref.tgm_m4_perspective(np.float32(np.float32(fov) * np.float32(np.pi) / np.float32(180.0)), np.float32(aspect), np.float32(0.1), np.float32(1000.0))
        ipr = ref.tgm_m4_inverse(ref.tgm_m4_mul(p, r))
        for name, sx, sy in (("ray_bl", -1, 1), ("ray_br", 1, 1), ("ray_tr", 1, -1), ("ray_tl", -1, -1)):
            v = ref.tgm_m4_mulv4(ipr, T.v4(sx, sy, 1, 1))
            n = ref.tgm_v3_normalized(T.v3(v.x, v.y, v.z))
            g = getattr(got, name)
            assert np.allclose([g.x, g.y, g.z], [n.x, n.y, n.z], rtol=0, atol=2e-7), name  # fov degrees->radians rounding is the only slack
        assert (got.camera.x, got.camera.y, got.camera.z) == tuple(np.float32(c) for c in pos)


def test_simplex_noise_bit_exact(oracle, ref):
    """tgm_simplex_noise (math/tg_math.c:182-302) on the arguments the procedural fill feeds it and on random points."""
    L = oracle.lib()
    rng = np.random.default_rng(5)
    pts = [rng.uniform(-300, 300, 3) for _ in range(20000)]
    for obj in (0, 1, 7):
        for _ in range(4000):
            x, y, z = rng.integers(0, 128), rng.integers(0, 128), rng.integers(0, 128)
            xf = np.float32(x) + np.float32(obj) * np.float32(1024.0)
            pts.append((xf * np.float32(0.008), 0.0, np.float32(z) * np.float32(0.008)))
            pts.append((xf * np.float32(0.2), 0.0, np.float32(z) * np.float32(0.2)))
            pts.append((np.float32(0.06) * xf, np.float32(0.06) * np.float32(y), np.float32(0.06) * np.float32(z)))
    bad = 0
    for p in pts:
        a = np.float32(L.tgo_simplex_noise(*map(np.float32, p)))
        b = np.float32(ref.tgm_simplex_noise(*map(np.float32, p)))
        bad += a.tobytes() != b.tobytes()
    assert bad == 0, f"{bad} of {len(pts)} noise values differ from the reference"


def test_xorshift_bit_exact(oracle, ref):
    L = oracle.lib()
    for seed in (1, 2, 0xDEADBEEF, 12345):
        a, b = T.u32(seed), T.u32(seed)
        for i in range(2000):
            if i % 3 == 0:
                assert L.tgo_pin_xorshift32_next(C.byref(a)) == ref.tgm_rand_xorshift32_next_u32(C.byref(b))
            elif i % 3 == 1:
                x, y = np.float32(L.tgo_pin_xorshift32_next_f32(C.byref(a))), np.float32(ref.tgm_rand_xorshift32_next_f32(C.byref(b)))
                assert x.tobytes() == y.tobytes()
            else:
                x = np.float32(L.tgo_pin_xorshift32_next_f32_range(C.byref(a), -1.0, 1.0))
                y = np.float32(ref.tgm_rand_xorshift32_next_f32_inclusive_range(C.byref(b), -1.0, 1.0))
                assert x.tobytes() == y.tobytes()
            assert a.value == b.value


def test_ray_aabb_and_sat_bit_exact(oracle, ref):
    """tg_intersect_ray_aabb (tg_physics.c:394-406) and tg_intersect_aabb_obb_ignore_contact (:226-392)."""
    L = oracle.lib()
    L.tgo_pin_intersect_ray_aabb_c.argtypes = [T.v3, T.v3, T.v3, T.v3, C.POINTER(T.f32), C.POINTER(T.f32)]
    L.tgo_pin_intersect_ray_aabb_c.restype = T.b32
    rng = np.random.default_rng(3)
    for trial in range(3000):
        o = rng.uniform(-20, 20, 3).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        if trial % 5 == 0:
            d[int(rng.integers(0, 3))] = 0.0
        if not d.any():
            d[1] = 1.0
        lo = rng.uniform(-10, 5, 3).astype(np.float32)
        hi = lo + rng.uniform(0.5, 12, 3).astype(np.float32)
        e0, x0, e1, x1 = T.f32(), T.f32(), T.f32(), T.f32()
        r0 = L.tgo_pin_intersect_ray_aabb_c(T.v3(*o), T.v3(*d), T.v3(*lo), T.v3(*hi), C.byref(e0), C.byref(x0))
        r1 = ref.tg_intersect_ray_aabb(T.v3(*o), T.v3(*d), T.v3(*lo), T.v3(*hi), C.byref(e1), C.byref(x1))
        assert bool(r0) == bool(r1)
        assert np.float32(e0.value).tobytes() == np.float32(e1.value).tobytes() and np.float32(x0.value).tobytes() == np.float32(x1.value).tobytes()
    for trial in range(3000):
        lo = rng.uniform(-6, 6, 3).astype(np.float32)
        hi = lo + rng.uniform(0.5, 9, 3).astype(np.float32)
        # an oriented box: rotate + translate the corners of an axis-aligned one (corner k: bit 0 -> x, bit 1 -> y, bit 2 -> z)
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        ext = rng.uniform(0.5, 10, 3)
        c = rng.uniform(-12, 12, 3)
        if trial % 6 == 0:
            q = np.eye(3)  # touching / axis-aligned cases exercise the <= / >= contact rules
            c = np.round(c); ext = np.round(ext) + 1; lo = np.round(lo); hi = lo + np.round(hi - lo) + 1
            lo, hi = lo.astype(np.float32), hi.astype(np.float32)
        corners = (T.v3 * 8)()
        for k in range(8):
            p = c + q @ (np.array([(k & 1), (k >> 1) & 1, (k >> 2) & 1]) * ext)
            corners[k] = T.v3(*p.astype(np.float32))
        r0 = L.tgo_intersect_aabb_obb_ignore_contact(T.v3(*lo), T.v3(*hi), corners)
        r1 = ref.tg_intersect_aabb_obb_ignore_contact(T.v3(*lo), T.v3(*hi), corners)
        assert bool(r0) == bool(r1), trial


def svo_arrays_of(svo):
    nodes = np.ctypeslib.as_array(svo.p_node_buffer, shape=(svo.node_buffer_count,)).copy()
    leaf = np.ctypeslib.as_array(C.cast(svo.p_leaf_node_data_buffer, C.POINTER(T.u32)), shape=(svo.leaf_node_data_buffer_count, 65)).copy()
    vox = np.ctypeslib.as_array(svo.p_voxels_buffer, shape=(svo.voxel_buffer_count_in_u32,)).copy()
    return nodes, leaf, vox


SVO_CASES = {
    "small_grid": lambda: scenes.small_grid(),
    "grid4_tall": lambda: scenes.small_grid(grid=4, dims=(3, 5, 2)),
    "config1_small": lambda: scenes.config1(k=3, width=64, height=36, dims=(6, 4, 6)),
    "dense_k1": lambda: scenes.config1(k=1, width=64, height=36, dims=(4, 4, 4)),
    "sparse_k5": lambda: scenes.small_grid(grid=3, k=5),
}


@pytest.mark.parametrize("name", sorted(SVO_CASES))
def test_svo_builder_equals_the_reference_builder(oracle, ref, name):
    """THE REFERENCE'S tg_svo_create (tg_sparse_voxel_octree.c:466-542, recursive, CPU) and the oracle's restatement on the
    same scene: node, leaf-record and voxel arrays identical. (The CUDA builder is compared with the oracle in
    tests/test_svo_gpu.py, so this closes the chain reference -> oracle -> K2.)"""
    s = SVO_CASES[name]()
    view = oracle.SceneView.from_scene(s, with_lut=False)
    want = T.tg_svo()
    scene = oracle.ref_scene(view)
    ref.tg_svo_create(T.v3(-512, -512, -512), T.v3(512, 512, 512), C.byref(scene), C.byref(want))
    wn, wl, wv = svo_arrays_of(want)
    got = oracle.svo_create(view, capacities=(1 << 21, 1 << 13, 1 << 14))  # the reference's capacities (:479-484)
    gn, gl, gv = oracle.svo_arrays(got)
    try:
        assert wl[:, 0].max() <= 64, "scene exceeds the 64 clusters per leaf the reference can hold without writing out of bounds"
        assert np.array_equal(gn, wn), f"{name}: node arrays differ"
        assert np.array_equal(gv, wv), f"{name}: voxel blocks differ"
        assert np.array_equal(gl, wl), f"{name}: leaf records differ"
        assert len(wn) > 5 and wv.any()
    finally:
        oracle.svo_destroy(got)
        ref.tg_svo_destroy(C.byref(want))


def test_svo_c_traversal_equals_the_reference_traversal(oracle, ref):
    """tg_svo_traverse (tg_sparse_voxel_octree.c:558-740) on the reference-built SVO against tgo_svo_traverse_c on the oracle-built one."""
    s = scenes.small_grid()
    view = oracle.SceneView.from_scene(s, with_lut=False)
    want = T.tg_svo()
    scene = oracle.ref_scene(view)
    ref.tg_svo_create(T.v3(-512, -512, -512), T.v3(512, 512, 512), C.byref(scene), C.byref(want))
    got = oracle.svo_create(view, capacities=(1 << 21, 1 << 13, 1 << 14))
    L = oracle.lib()
    rng = np.random.default_rng(9)
    n_hits = 0
    try:
        for trial in range(4000):
            o = rng.uniform(-90, 90, 3).astype(np.float32)
            o[1] = np.float32(rng.uniform(-30, 120))
            target = rng.uniform(-40, 40, 3).astype(np.float32)
            d = target - o
            if trial % 9 == 0:
                d[int(rng.integers(0, 3))] = 0.0
            if not d.any():
                d[1] = -1.0
            d = (d / np.linalg.norm(d)).astype(np.float32)
            d0, n0, v0, d1, n1, v1 = T.f32(), T.u32(), T.u32(), T.f32(), T.u32(), T.u32()
            r0 = L.tgo_svo_traverse_c(C.byref(got), T.v3(*o), T.v3(*d), C.byref(d0), C.byref(n0), C.byref(v0))
            r1 = ref.tg_svo_traverse(C.byref(want), T.v3(*o), T.v3(*d), C.byref(d1), C.byref(n1), C.byref(v1))
            assert bool(r0) == bool(r1), (trial, o, d)
            if r1:
                n_hits += 1
                assert np.float32(d0.value).tobytes() == np.float32(d1.value).tobytes() and n0.value == n1.value and v0.value == v1.value, (trial, o, d)
        assert n_hits > 300
    finally:
        oracle.svo_destroy(got)
        ref.tg_svo_destroy(C.byref(want))


# ---- the same pins from the committed reference-made fixtures (work without /root/reference and without oracle/_ref) ----
import os  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(SVO_CASES))
def test_oracle_svo_equals_reference_made_fixture(oracle, name):
    g = np.load(os.path.join(GOLDEN, f"ref_svo_{name}.npz"))
    view = oracle.SceneView.from_scene(SVO_CASES[name](), with_lut=False)
    got = oracle.svo_create(view, capacities=(1 << 21, 1 << 13, 1 << 14))
    gn, gl, gv = oracle.svo_arrays(got)
    oracle.svo_destroy(got)
    assert np.array_equal(gn, g["nodes"]) and np.array_equal(gl, g["leaf"])
    nz = np.nonzero(gv)[0].astype(np.uint32)
    assert gv.size == int(g["n_voxel_words"]) and np.array_equal(nz, g["voxels_nonzero_idx"]) and np.array_equal(gv[nz], g["voxels_nonzero"])


def test_oracle_simplex_noise_equals_reference_made_fixture(oracle):
    g = np.load(os.path.join(GOLDEN, "ref_simplex_noise.npz"))
    L = oracle.lib()
    got = np.array([L.tgo_simplex_noise(*map(np.float32, p)) for p in g["points"]], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), g["values"].view(np.uint32))


def test_boundary_tg_svo_traverse_equals_the_reference(oracle, ref):
    """The product's own host entry point tg_svo_traverse (include/tg_raytracer.h, tg_sparse_voxel_octree.h:53) on a
    reference-built SVO against the reference's tg_svo_traverse: same hit flag, distance bits, node and voxel index."""
    import tg_b200
    P = tg_b200.lib()
    s = scenes.small_grid(grid=4, dims=(3, 5, 2))
    view = oracle.SceneView.from_scene(s, with_lut=False)
    svo = T.tg_svo()
    scene = oracle.ref_scene(view)
    ref.tg_svo_create(T.v3(-512, -512, -512), T.v3(512, 512, 512), C.byref(scene), C.byref(svo))
    rng = np.random.default_rng(21)
    n_hits = 0
    try:
        for trial in range(3000):
            o = rng.uniform(-100, 100, 3).astype(np.float32)
            d = (rng.uniform(-40, 40, 3).astype(np.float32) - o)
            if trial % 11 == 0:
                d[int(rng.integers(0, 3))] = 0.0
            if not d.any():
                d[2] = 1.0
            d = (d / np.linalg.norm(d)).astype(np.float32)
            d0, n0, v0, d1, n1, v1 = T.f32(), T.u32(), T.u32(), T.f32(), T.u32(), T.u32()
            r0 = P.tg_svo_traverse(C.byref(svo), T.v3(*o), T.v3(*d), C.byref(d0), C.byref(n0), C.byref(v0))
            r1 = ref.tg_svo_traverse(C.byref(svo), T.v3(*o), T.v3(*d), C.byref(d1), C.byref(n1), C.byref(v1))
            assert bool(r0) == bool(r1), (trial, o, d)
            if r1:
                n_hits += 1
                assert np.float32(d0.value).tobytes() == np.float32(d1.value).tobytes() and n0.value == n1.value and v0.value == v1.value, (trial, o, d)
        assert n_hits > 200
    finally:
        ref.tg_svo_destroy(C.byref(svo))


@pytest.mark.parametrize("name", ["small_grid", "grid4_tall"])
def test_glsl_traversal_transcription_agrees_with_the_reference_c_traversal(oracle, ref, name):
    """G3: the oracle's transcription of the GLSL tg_svo_traverse (svo_functions.inc:1-329) -- the variant the GI kernels follow,
    which cannot be executed here -- against THE REFERENCE'S OWN C traversal (tg_sparse_voxel_octree.c:558-740) run on a
    reference-built SVO. The two variants differ in bookkeeping (Q7: how the distance is accumulated, the DDA start clamp), so the
    distance is compared to 1e-4 relative, but on 20,000 random rays per scene they must make the same hit / miss decision and
    name the same leaf node and the same voxel (the C variant reports leaf data_pointer * 32768 + voxel)."""
    s = SVO_CASES[name]()
    view = oracle.SceneView.from_scene(s, with_lut=False)
    want = T.tg_svo()
    scene = oracle.ref_scene(view)
    ref.tg_svo_create(T.v3(-512, -512, -512), T.v3(512, 512, 512), C.byref(scene), C.byref(want))
    nodes = np.ctypeslib.as_array(want.p_node_buffer, shape=(want.node_buffer_count,))
    L = oracle.lib()
    rng = np.random.default_rng(21)
    n_hits = 0
    try:
        for trial in range(20000):
            o = rng.uniform(-90, 90, 3).astype(np.float32)
            o[1] = np.float32(rng.uniform(-30, 120))
            d = rng.uniform(-40, 40, 3).astype(np.float32) - o
            if trial % 9 == 0:
                d[int(rng.integers(0, 3))] = 0.0
            if not d.any():
                d[1] = -1.0
            d = (d / np.linalg.norm(d)).astype(np.float32)
            hp, hn, node, voxel = T.v3(), T.v3(), T.u32(), T.u32()
            depth = L.tgo_svo_traverse_glsl(C.byref(want), 1000.0, T.v3(*o), T.v3(*d), C.byref(hp), C.byref(hn), C.byref(node), C.byref(voxel))
            d1, n1, v1 = T.f32(), T.u32(), T.u32()
            r1 = ref.tg_svo_traverse(C.byref(want), T.v3(*o), T.v3(*d), C.byref(d1), C.byref(n1), C.byref(v1))
            assert (depth < 1.0) == bool(r1), (trial, o, d, depth, d1.value)
            if r1:
                n_hits += 1
                assert node.value == n1.value and voxel.value == v1.value - int(nodes[n1.value]) * 32768, (trial, o, d)
                assert abs(depth * 1000.0 - d1.value) <= 1e-4 * max(1.0, abs(d1.value)), (trial, depth * 1000.0, d1.value)
        assert n_hits > 5000
    finally:
        ref.tg_svo_destroy(C.byref(want))


def test_visibility_transcription_sees_the_voxels_the_reference_traversal_sees(oracle, ref):
    """V3-V7: visibility.frag cannot run here and has no CPU twin, but for axis-aligned objects on the integer lattice a cluster
    voxel IS a world unit cell IS a voxel of the reference's SVO, so the reference's own code answers the same question by another
    route: tg_svo_create + tg_svo_traverse (tg_sparse_voxel_octree.c) along the primary ray of every pixel. The oracle's
    visibility buffer must make the same hit / miss decision, name the same world voxel (pointer -> object, cluster, voxel) and a
    depth24 within 2 LSB of the traversal's distance (different arithmetic, same geometry) on every pixel."""
    specs = [((0.0, 0.0, 0.0), (32, 16, 32)), ((48.0, 8.0, -16.0), (16, 32, 16)), ((-40.0, -8.0, 24.0), (24, 16, 40)), ((8.0, 40.0, 8.0), (16, 8, 16))]
    objs = []
    for i, (c, e) in enumerate(specs):
        n = (e[0] // 8) * (e[1] // 8) * (e[2] // 8)
        objs.append(scenes.ObjectSpec(center=c, extent=e, angle=0.0, bits=scenes.random_solid_bits(100 + i, n, 3)))
    w, h = 160, 90
    cam = scenes.CameraSpec(position=(6.0, 34.0, 66.0), pitch=float(scenes.deg2rad(-28.0)), yaw=float(scenes.deg2rad(6.0)), roll=0.0, aspect=w / h)
    s = scenes.SceneSpec("aligned", w, h, cam, objs)
    view = oracle.SceneView.from_scene(s, with_lut=False)
    rays = oracle.camera_rays(oracle.camera_from_spec(cam))
    vis, _ = oracle.visibility(view, rays, w, h, oracle.VIS_BRUTE_FORCE)
    svo = T.tg_svo()
    scene = oracle.ref_scene(view)
    ref.tg_svo_create(T.v3(-512, -512, -512), T.v3(512, 512, 512), C.byref(scene), C.byref(svo))
    nodes = np.ctypeslib.as_array(svo.p_node_buffer, shape=(svo.node_buffer_count,))
    firsts = np.cumsum([0] + [o.n_clusters for o in objs])
    L = oracle.lib()
    o = np.array(cam.position, dtype=np.float32)
    n_hits = 0
    try:
        for py in range(h):
            for px in range(w):
                dv = L.tgo_pixel_ray_direction_nn(C.byref(rays), w, h, px, py)
                d = np.array([dv.x, dv.y, dv.z], dtype=np.float32)
                d = (d / np.float32(np.sqrt(np.float32(d @ d)))).astype(np.float32)
                dist, node, voxel = T.f32(), T.u32(), T.u32()
                r = ref.tg_svo_traverse(C.byref(svo), T.v3(*o), T.v3(*d), C.byref(dist), C.byref(node), C.byref(voxel))
                word = int(vis[py, px])
                assert (word != 0xFFFFFFFFFFFFFFFF) == bool(r), (px, py)
                if not r:
                    continue
                n_hits += 1
                pointer, vox, depth24 = (word >> 9) & 0x7FFFFFFF, word & 511, word >> 40
                oi = int(np.searchsorted(firsts, pointer, side="right") - 1)
                nx, ny, _ = objs[oi].dims
                rel = pointer - firsts[oi]
                seen = (np.array(objs[oi].center) - np.array(objs[oi].extent) / 2 + 8 * np.array([rel % nx, (rel // nx) % ny, rel // (nx * ny)])
                        + np.array([vox % 8, (vox // 8) % 8, vox // 64]))
                rel_voxel = voxel.value - int(nodes[node.value]) * 32768   # the C traversal reports data_pointer * 32768 + voxel
                p = o.astype(np.float64) + d.astype(np.float64) * (dist.value + 1e-3)
                traversed = np.floor((p + 512) / 32) * 32 - 512 + np.array([rel_voxel % 32, (rel_voxel // 32) % 32, rel_voxel // 1024])
                assert np.array_equal(seen, traversed), (px, py, seen, traversed)
                assert abs(depth24 - int(np.float32(dist.value) / np.float32(1000.0) * np.float32(16777215.0))) <= 2, (px, py)
        assert n_hits > 2500, n_hits
    finally:
        ref.tg_svo_destroy(C.byref(svo))

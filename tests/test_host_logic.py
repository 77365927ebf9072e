"""CPU: the C-ABI library loads, exports every symbol include/tg_raytracer.h declares, its pure-host logic
(scene bookkeeping of tgvk_raytracer.c:662-712,805-866,994-1077; camera maths of tgvk_core.c:382-444; the per-object
factorisation the kernels rely on) matches hand-derived expectations and the oracle. No compute call is made."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import tg_b200
from tg_b200 import ctypes_defs as T
from tg_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = tg_b200.lib()
    header = open(os.path.join(ROOT, "include", "tg_raytracer.h")).read()
    declared = set(re.findall(r"TG_EXPORT\s+[\w\s\*]+?\b(tg\w+|tgb200_\w+)\s*\(", header))
    assert len(declared) >= 45
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/tg_raytracer.h but not exported"
    assert declared == set(tg_b200.SYMBOLS), declared ^ set(tg_b200.SYMBOLS)


def test_struct_sizes_match_the_reference_layouts():
    # 44, 96, 4, 260
    assert C.sizeof(T.tg_voxel_object) == 44 and C.sizeof(T.tg_object_data) == 96 and C.sizeof(T.tg_svo_leaf_node_data) == 260
    assert C.sizeof(T.m4) == 64 and C.sizeof(T.v3) == 12 and C.sizeof(T.tg_camera) == 52


def test_free_lists_are_lifo_ascending_and_pointer_ranges_contiguous():
    L = tg_b200.lib()
    s = T.tg_scene()
    L.tgb200_scene_init(C.byref(s), 4, 100)
    assert s.n_available_object_indices == 4 and s.n_available_cluster_indices == 100
    ids = [L.tgb200_scene_alloc_object(C.byref(s), T.v3(0, 0, 0), T.v3u(16, 8, 24), 0.0, T.v3(0, 1, 0)),
           L.tgb200_scene_alloc_object(C.byref(s), T.v3(1, 2, 3), T.v3u(8, 8, 8), 0.5, T.v3(0, 1, 0)),
           L.tgb200_scene_alloc_object(C.byref(s), T.v3(4, 5, 6), T.v3u(8, 16, 8), 0.7, T.v3(0, 1, 0))]
    assert ids == [0, 1, 2]  # tgvk_raytracer.c:704-712: pops yield 0, 1, 2, ...
    assert s.n_objects == 3 and s.n_cluster_pointers == 6 + 1 + 2
    assert [s.p_objects[i].first_cluster_pointer for i in range(3)] == [0, 6, 7]
    assert s.p_objects[0].n_cluster_pointers_per_dim.tuple() == (2, 1, 3)
    assert [s.p_cluster_pointers[i] for i in range(9)] == list(range(9))  # fresh scene: idx == pointer
    assert [s.p_cluster_idx_to_object_idx[i] for i in range(9)] == [0] * 6 + [1] + [2] * 2
    assert L.tg_object_is_initialized(C.byref(s), 1) and not L.tg_object_is_initialized(C.byref(s), 3)

    # destroy the middle object: tgvk_raytracer.c:1009-1066
    first, n = T.u32(), T.u32()
    L.tgb200_scene_free_object(C.byref(s), 1, C.byref(first), C.byref(n))
    assert (first.value, n.value) == (6, 2)
    assert s.n_objects == 2 and s.n_cluster_pointers == 8
    assert [s.p_cluster_pointers[i] for i in range(8)] == [0, 1, 2, 3, 4, 5, 7, 8]   # compacted downward
    assert s.p_objects[2].first_cluster_pointer == 6                                  # later object's first pointer decremented
    assert not L.tg_object_is_initialized(C.byref(s), 1)
    # the freed ids come back first (LIFO): object 1, cluster 6
    idx = L.tgb200_scene_alloc_object(C.byref(s), T.v3(0, 0, 0), T.v3u(8, 8, 16), 0.0, T.v3(0, 1, 0))
    assert idx == 1 and s.p_objects[1].first_cluster_pointer == 8
    assert [s.p_cluster_pointers[i] for i in (8, 9)] == [6, 9]
    assert s.p_cluster_idx_to_object_idx[6] == 1 and s.p_cluster_idx_to_object_idx[9] == 1
    assert tg_b200.lib().tgb200_last_error() is None
    L.tgb200_scene_free(C.byref(s))


def test_preconditions_are_recorded_errors():
    L = tg_b200.lib()
    s = T.tg_scene()
    L.tgb200_scene_init(C.byref(s), 1, 4)
    L.tgb200_clear_error()
    assert L.tgb200_scene_alloc_object(C.byref(s), T.v3(0, 0, 0), T.v3u(12, 8, 8), 0.0, T.v3(0, 1, 0)) == T.TG_U32_MAX  # extent % 8 (tgvk_raytracer.c:808-810)
    assert b"multiples of 8" in L.tgb200_last_error()
    L.tgb200_clear_error()
    assert L.tgb200_scene_alloc_object(C.byref(s), T.v3(0, 0, 0), T.v3u(8, 8, 64), 0.0, T.v3(0, 1, 0)) == T.TG_U32_MAX  # 8 clusters > capacity 4
    assert b"do not fit" in L.tgb200_last_error()
    L.tgb200_clear_error()
    L.tgb200_scene_free(C.byref(s))


def test_camera_rays_match_oracle_bitwise(oracle):
    L = tg_b200.lib()
    specs = [scenes.config1().camera, scenes.config2(grid=1, width=64, height=36).camera,
             scenes.CameraSpec((65.1368790, -30.7384720, 73.0285263), -0.173136666, 0.710419059, 0.0)]  # tg_application.c:51-59
    rng = np.random.default_rng(3)
    specs += [scenes.CameraSpec(tuple(rng.uniform(-900, 900, 3)), float(rng.uniform(-1.5, 1.5)), float(rng.uniform(-3, 3)), float(rng.uniform(-0.5, 0.5)),
                                fov_y_deg=float(rng.uniform(30, 110)), aspect=float(rng.uniform(0.5, 2.5))) for _ in range(20)]
    for spec in specs:
        cam = oracle.camera_from_spec(spec)
        a = oracle.camera_rays(cam)
        b = T.tg_camera_rays()
        L.tgb200_camera_rays(C.byref(cam), C.byref(b))
        assert bytes(a) == bytes(b)


def test_pack_color_truncates(oracle):
    L = tg_b200.lib()
    for rgb in [(1.0, 0.0, 0.0), (0.5, 0.25, 0.999), (0.2, 0.4, 0.6)] + [c for c in scenes.reference_lut_ramp(256)[::17]]:
        want = (int(np.float32(rgb[0]) * np.float32(255)) << 24) | (int(np.float32(rgb[1]) * np.float32(255)) << 16) | (int(np.float32(rgb[2]) * np.float32(255)) << 8) | 255
        assert L.tgb200_pack_color(*rgb) == want == oracle.lib().tgo_pack_color(*rgb) == scenes.pack_color(*rgb)


def test_per_object_factorisation_equals_the_full_chain(oracle):
    """tgb_hoist.h (what K1/K3 evaluate per cluster) against the oracle's literal ws2ms3*ws2ms2*ws2ms1*ws2ms0 chain."""
    L = tg_b200.lib()
    OL = oracle.lib()
    rng = np.random.default_rng(0)
    f32 = np.float32
    n_checked = 0
    for trial in range(120):
        vo = T.tg_voxel_object()
        dims = [int(x) for x in rng.integers(1, 20, 3)]
        vo.n_cluster_pointers_per_dim = T.v3u(*dims)
        vo.first_cluster_pointer = int(rng.integers(0, 1000))
        vo.translation = T.v3(*(rng.uniform(-3000, 3000, 3) if trial % 3 else np.zeros(3)))
        vo.angle_in_radians = float(rng.uniform(-7, 7)) if trial % 5 else 0.0
        ax = rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        if trial % 2 == 0:
            ax = np.array([0.0, 1.0, 0.0])
        vo.axis = T.v3(*ax)
        od = T.tg_object_data()
        OL.tgo_object_data(C.byref(vo), 0, C.byref(od))
        spec = scenes.CameraSpec(tuple(rng.uniform(-500, 500, 3)), float(rng.uniform(-1, 1)), float(rng.uniform(-3, 3)), 0.0)
        if trial % 11 == 0:
            spec = scenes.CameraSpec((0.0, 0.0, 0.0), 0.0, 0.0, 0.0)
        rays = oracle.camera_rays(oracle.camera_from_spec(spec))
        for _ in range(10):
            cp = vo.first_cluster_pointer + int(rng.integers(0, dims[0] * dims[1] * dims[2]))
            px, py = int(rng.integers(0, 640)), int(rng.integers(0, 360))
            m = OL.tgo_ws2ms(C.byref(od), cp)
            mm = np.array([getattr(m, f[0]) for f in T.m4._fields_], dtype=f32).reshape(4, 4).T
            d_ws = OL.tgo_pixel_ray_direction_nn(C.byref(rays), 640, 360, px, py)

            def mulv(v, w):  # tgm_m4_mulv4 order, float32, no contraction
                return [((f32(v[0]) * mm[i, 0] + f32(v[1]) * mm[i, 1]) + f32(v[2]) * mm[i, 2]) + f32(w) * mm[i, 3] for i in range(3)]
            o_ref = mulv([rays.camera.x, rays.camera.y, rays.camera.z], 1.0)
            r = mulv([d_ws.x, d_ws.y, d_ws.z], 0.0)
            mag = np.sqrt(f32(r[0] * r[0] + r[1] * r[1]) + r[2] * r[2], dtype=f32)
            ref = np.array(o_ref + [f32(r[i]) / mag for i in range(3)], dtype=f32)
            o, d = T.v3(), T.v3()
            L.tgb200_debug_cluster_ray(C.byref(od), C.byref(rays), 640, 360, px, py, cp, C.byref(o), C.byref(d))
            got = np.array([o.x, o.y, o.z, d.x, d.y, d.z], dtype=f32)
            assert np.array_equal(got, ref), (trial, got, ref)  # == treats +0/-0 alike: a signed zero never reaches a packed word
            n_checked += 1
    assert n_checked == 1200


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a CUDA device tg_raytracer_create must fail with an error (and with one, succeed): never a CPU path."""
    L = tg_b200.lib()
    cam = T.make_camera((0, 0, 10), 0, 0, 0, 70, 1.0, 0.1, 100)
    if L.tgb200_device_count() == 0:
        with pytest.raises(tg_b200.TgError, match="no CUDA device|no CPU fallback"):
            tg_b200.Raytracer(cam, 1, 1, 16, 16)
        rt = T.tg_raytracer()
        L.tg_raytracer_create(C.byref(cam), 1, 1, C.byref(rt))
        assert not rt.p_device and L.tgb200_last_error()
        L.tgb200_clear_error()
        L.tg_raytracer_render(C.byref(rt))  # every later call is a recorded error, not a computation
        assert b"no device state" in L.tgb200_last_error()
        L.tgb200_clear_error()
    else:
        tg_b200.Raytracer(cam, 1, 1, 16, 16).destroy()


def test_product_does_not_reference_the_oracle():
    """The product tree must not include, link or import anything under oracle/."""
    pkg = os.path.join(ROOT, "tg_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".c", ".cu", ".cuh", ".h", ".py", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for pattern in (r"#include\s+\"[^\"]*(oracle|tgo)", r"\bimport\s+oracle", r"\bfrom\s+oracle", r"\btgo_\w+\s*\(", r"libtgo"):
                    assert not re.search(pattern, text), (os.path.join(dirpath, f), pattern)
    import subprocess
    out = subprocess.run(["ldd", tg_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtgo" not in out


def test_tg_svo_traverse_host_query_matches_the_oracle_c_variant():
    """tg_svo_traverse (tg_sparse_voxel_octree.h:53) is a HOST function in the reference and here: random rays over an
    oracle-built SVO, library vs the oracle's restatement of tg_sparse_voxel_octree.c:558-740, exact equality."""
    from oracle import oracle as O
    L = tg_b200.lib()
    s = scenes.small_grid()
    svo = O.svo_create(O.SceneView.from_scene(s, with_lut=False))
    rng = np.random.default_rng(5)
    n_hit = 0
    for k in range(2000):
        o = rng.uniform(-90, 90, 3).astype(np.float32)
        if k % 3 == 0:
            o[1] = np.float32(rng.uniform(20, 200))
        d = rng.normal(size=3).astype(np.float32)
        if k % 7 == 0:
            d[int(rng.integers(3))] = 0.0  # axis-parallel components: the -/+F32_MAX branches
        d = d / np.float32(np.sqrt((d * d).sum(dtype=np.float32)))
        got = (T.f32(), T.u32(), T.u32())
        want = (T.f32(), T.u32(), T.u32())
        a = L.tg_svo_traverse(C.byref(svo), T.v3(*o), T.v3(*d), C.byref(got[0]), C.byref(got[1]), C.byref(got[2]))
        b = O.lib().tgo_svo_traverse_c(C.byref(svo), T.v3(*o), T.v3(*d), C.byref(want[0]), C.byref(want[1]), C.byref(want[2]))
        assert bool(a) == bool(b), k
        assert [g.value for g in got] == [w.value for w in want], (k, [g.value for g in got], [w.value for w in want])
        n_hit += bool(a)
    assert 100 < n_hit < 1900
    O.svo_destroy(svo)
    assert L.tgb200_last_error() is None


def test_synthetic_generator_host_twin_equals_scene_definition():
    """tgb200_synthetic_solid_bits (C, the host twin of k_synthetic_fill) == scenes.random_solid_bits (numpy): the seeded random
    fill of BASELINE's configs is one definition wherever it is evaluated."""
    import ctypes as C
    import tg_b200
    from tg_b200 import ctypes_defs as T, scenes
    L = tg_b200.lib()
    for seed, n, k in ((1, 64, 3), (7, 257, 1), (12288, 2048, 3), (0xFFFFFFFF, 5, 5)):
        out = np.empty((n, 16), dtype=np.uint32)
        L.tgb200_synthetic_solid_bits(seed, k, n, T.ptr(out, T.u32))
        assert np.array_equal(out, scenes.random_solid_bits(seed, n, k)), (seed, n, k)


def test_bmp_writer_container(tmp_path):
    """tgb200_write_bmp_bgra8: the container of the reference's tg_image_store_to_disc (tg_image_io.c:438-520) -- 'BM', V5 header
    (124 bytes), BI_BITFIELDS masks of B8G8R8A8, 32 bpp, negative height (top-down), pixels at byte 150 -- read back byte for byte."""
    import struct
    import tg_b200
    from tg_b200 import ctypes_defs as T
    L = tg_b200.lib()
    w, h = 5, 3
    px = (np.arange(w * h, dtype=np.uint32) * np.uint32(0x01030507)) | np.uint32(0xFF000000)
    path = str(tmp_path / "frame.bmp")
    assert L.tgb200_write_bmp_bgra8(path.encode(), w, h, T.ptr(px, T.u32))
    raw = open(path, "rb").read()
    assert raw[:2] == b"BM" and len(raw) == 150 + w * h * 4
    size, _, _, offset = struct.unpack_from("<IHHI", raw, 2)
    assert size == len(raw) and offset == 150
    hdr, bw, bh, planes, bpp, compression, image_size = struct.unpack_from("<IiiHHII", raw, 14)
    assert (hdr, bw, bh, planes, bpp, compression, image_size) == (124, w, -h, 1, 32, 3, w * h * 4)
    assert struct.unpack_from("<IIII", raw, 14 + 40) == (0x00FF0000, 0x0000FF00, 0x000000FF, 0xFF000000)
    assert np.array_equal(np.frombuffer(raw, dtype="<u4", offset=150), px)
    assert not L.tgb200_write_bmp_bgra8(str(tmp_path / "no_dir" / "x.bmp").encode(), w, h, T.ptr(px, T.u32))
    L.tgb200_clear_error()


def test_row_mapping_header_equals_the_sharding_plan(tmp_path):
    """tg_b200/csrc/tgb_rows.h (16-row bands dealt out to the ranks; what K1, k_shade and the host read-back paths use) compiled
    with gcc and compared with tg_b200/sharding.py for ragged and full-size frames: tile size, frame row of every tile row,
    virtual row of every frame row, identity on one rank."""
    import subprocess
    from tg_b200 import sharding
    src = tmp_path / "rows.c"
    src.write_text(r'''
#include <stdio.h>
#include "tgb_rows.h"
int main(void)
{
    const unsigned cases[][2] = { {2160, 8}, {2160, 4}, {2160, 1}, {177, 2}, {10, 4}, {33, 8}, {16, 3} };
    for (unsigned c = 0; c < sizeof(cases) / sizeof(cases[0]); c++)
    {
        const unsigned h = cases[c][0], n = cases[c][1], t = tgb_tile_rows_for(h, n);
        printf("case %u %u %u\n", h, n, t);
        for (unsigned v = 0; v < n * t; v++) printf("%u ", tgb_row_to_physical(v, n, t));
        printf("\n");
        for (unsigned p = 0; p < h; p++) printf("%u ", tgb_row_to_virtual(p, n, t));
        printf("\n");
    }
    return 0;
}
''')
    exe = tmp_path / "rows"
    csrc = os.path.join(ROOT, "tg_b200", "csrc")
    subprocess.check_call(["gcc", "-std=gnu11", "-O1", "-I", csrc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-lm"])
    lines = subprocess.check_output([str(exe)], text=True).strip().split("\n")
    for i in range(0, len(lines), 3):
        _, h, n, t = lines[i].split()
        h, n, t = int(h), int(n), int(t)
        phys = np.array(lines[i + 1].split(), dtype=np.int64)
        virt = np.array(lines[i + 2].split(), dtype=np.int64)
        assert t == sharding.tile_row_count(h, n)
        for r in range(n):
            want = sharding.tile_physical_rows(h, n, r)
            got = phys[r * t:(r + 1) * t]
            if n == 1:
                assert np.array_equal(got[:h], want)      # one rank: frame order, padding rows at the end
            else:
                assert np.array_equal(np.where(got < h, got, -1), want), (h, n, r)
        assert np.array_equal(phys[virt], np.arange(h)), (h, n)   # the two maps are inverse to each other on the frame's rows
        if n == 1:
            assert np.array_equal(virt, np.arange(h))


def test_non_unit_rotation_axis_is_rejected_and_a_failed_create_leaves_the_scene_untouched():
    """tgm_m4_angle_axis does not normalise its axis and the object-level culling assumes a rigid transform, so the axis is
    checked where it enters; and a create that fails must not consume a slot (host-only scene bookkeeping, no GPU needed)."""
    import ctypes as C
    import tg_b200
    from tg_b200 import ctypes_defs as T
    L = tg_b200.lib()
    L.tgb200_clear_error()
    scene = T.tg_scene()
    L.tgb200_scene_init(C.byref(scene), 4, 64)
    try:
        a = L.tgb200_scene_alloc_object(C.byref(scene), T.v3(0, 0, 0), T.v3u(8, 8, 8), 0.5, T.v3(0.0, 1.0, 0.0))
        assert a != 0xFFFFFFFF and L.tgb200_last_error() is None
        before = (scene.n_objects, scene.n_cluster_pointers, scene.n_available_object_indices, scene.n_available_cluster_indices)
        for axis in ((0.0, 2.0, 0.0), (0.0, 0.0, 0.0), (0.6, 0.6, 0.6)):
            b = L.tgb200_scene_alloc_object(C.byref(scene), T.v3(0, 0, 0), T.v3u(8, 8, 8), 0.5, T.v3(*axis))
            assert b == 0xFFFFFFFF and b"unit vector" in L.tgb200_last_error()
            L.tgb200_clear_error()
            assert before == (scene.n_objects, scene.n_cluster_pointers, scene.n_available_object_indices, scene.n_available_cluster_indices)
        # a normalised non-axis-aligned axis is fine
        n = 3 ** -0.5
        c = L.tgb200_scene_alloc_object(C.byref(scene), T.v3(0, 0, 0), T.v3u(8, 8, 8), 0.5, T.v3(n, n, n))
        assert c != 0xFFFFFFFF and L.tgb200_last_error() is None
    finally:
        L.tgb200_scene_free(C.byref(scene))
        L.tgb200_clear_error()

"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/libtgo.so (the scalar C restatement of the reference's voxel-rendering path)
and of oracle/_ref/libtg_ref.so (the reference's own portable C files -- math, physics, Amanatides-Woo, the CPU SVO builder and
traversal -- compiled from /root/reference by oracle/Makefile; used to pin the restatement).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this;
nothing under tg_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from tg_b200 import ctypes_defs as T

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

VIS_BRUTE_FORCE = 0
VIS_SCREEN_RECT = 1


class tgo_scene_view(C.Structure):
    _fields_ = [("n_objects_capacity", T.u32), ("p_objects", C.POINTER(T.tg_object_data)),
                ("n_cluster_pointers", T.u32), ("p_cluster_pointers", C.POINTER(T.u32)),
                ("p_cluster_idx_to_object_idx", C.POINTER(T.u32)), ("p_voxel_cluster_data", C.POINTER(T.u32)),
                ("p_color_lut_idx_data", C.POINTER(T.u8)), ("p_color_lut", C.POINTER(T.u32)),
                ("global_pointer_base", T.u32)]


def build(force=False):
    so = os.path.join(_HERE, "libtgo.so")
    srcs = [os.path.join(_HERE, f) for f in ("tgo_visibility.c", "tgo_svo.c", "tgo_shade.c", "tgo_procedural.c", "tgo.h", "tgo_math.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "libtgo.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/tg/src/graphics/tg_sparse_voxel_octree.c") and (force or not os.path.exists(os.path.join(_HERE, "_ref", "libtg_ref.so"))):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "libtgo.so"))
        L.tgo_camera_rays.argtypes = [C.POINTER(T.tg_camera), C.POINTER(T.tg_camera_rays)]
        L.tgo_object_data.argtypes = [C.POINTER(T.tg_voxel_object), T.u32, C.POINTER(T.tg_object_data)]
        L.tgo_pack_color.argtypes = [T.f32, T.f32, T.f32]
        L.tgo_pack_color.restype = T.u32
        L.tgo_ws2ms.argtypes = [C.POINTER(T.tg_object_data), T.u32]
        L.tgo_ws2ms.restype = T.m4
        L.tgo_pixel_ray_direction_nn.argtypes = [C.POINTER(T.tg_camera_rays), T.u32, T.u32, T.u32, T.u32]
        L.tgo_pixel_ray_direction_nn.restype = T.v3
        L.tgo_visibility_fragment.argtypes = [C.POINTER(tgo_scene_view), C.POINTER(T.tg_camera_rays), T.u32, T.u32, T.u32, T.u32, T.u32]
        L.tgo_visibility_fragment.restype = T.u64
        L.tgo_visibility.argtypes = [C.POINTER(tgo_scene_view), C.POINTER(T.tg_camera_rays), T.u32, T.u32, T.u32, T.u32, T.u32, T.u32, C.POINTER(T.u64)]
        L.tgo_visibility.restype = T.u64
        L.tgo_visibility_window.argtypes = [C.POINTER(tgo_scene_view), C.POINTER(T.tg_camera_rays), T.u32, T.u32, T.u32, T.u32, T.u32, T.u32, C.POINTER(T.u64)]
        L.tgo_visibility_window.restype = T.u64
        L.tgo_max_threads.restype = T.i32
        L.tgo_set_threads.argtypes = [T.i32]
        L.tgo_cluster_dda.argtypes = [C.POINTER(T.u32), T.v3, T.v3, T.f32]
        L.tgo_cluster_dda.restype = T.i32
        L.tgo_svo_create.argtypes = [T.v3, T.v3, C.POINTER(tgo_scene_view), C.POINTER(T.tg_voxel_object), C.POINTER(T.tg_svo)]
        L.tgo_svo_destroy.argtypes = [C.POINTER(T.tg_svo)]
        L.tgo_svo_traverse_glsl.argtypes = [C.POINTER(T.tg_svo), T.f32, T.v3, T.v3, C.POINTER(T.v3), C.POINTER(T.v3), C.POINTER(T.u32), C.POINTER(T.u32)]
        L.tgo_svo_traverse_glsl.restype = T.f32
        L.tgo_visibility_svo.argtypes = [C.POINTER(T.tg_svo), C.POINTER(T.tg_camera_rays), T.u32, T.u32, T.u32, T.u32, T.u32, C.POINTER(T.u64)]
        L.tgo_svo_traverse_c.argtypes = [C.POINTER(T.tg_svo), T.v3, T.v3, C.POINTER(T.f32), C.POINTER(T.u32), C.POINTER(T.u32)]
        L.tgo_svo_traverse_c.restype = T.b32
        L.tgo_amanatides_woo.argtypes = [T.v3, T.v3, T.v3, C.POINTER(T.u32), C.POINTER(T.v3i)]
        L.tgo_amanatides_woo.restype = T.b32
        L.tgo_intersect_aabb_obb_ignore_contact.argtypes = [T.v3, T.v3, C.POINTER(T.v3)]
        L.tgo_intersect_aabb_obb_ignore_contact.restype = T.b32
        L.tgo_shade.argtypes = [C.POINTER(tgo_scene_view), C.POINTER(T.tg_camera_rays), T.u32, T.u32, C.POINTER(T.u64), C.POINTER(T.tg_svo),
                                T.u32, T.u32, T.u32, T.u32, T.u32, T.u32, C.POINTER(T.f32)]
        L.tgo_shade_gi_rays.argtypes = [C.POINTER(tgo_scene_view), C.POINTER(T.tg_camera_rays), T.u32, T.u32, C.POINTER(T.u64), C.POINTER(T.tg_svo),
                                        T.u32, T.u32, T.u32, T.u32, C.POINTER(T.f32)]
        L.tgo_present_bgra8.argtypes = [C.POINTER(T.f32), T.u64, C.POINTER(T.u32)]
        L.tgo_simplex_noise.argtypes = [T.f32, T.f32, T.f32]
        L.tgo_simplex_noise.restype = T.f32
        L.tgo_procedural_voxel_is_solid.argtypes = [T.u32, T.u32, T.u32, T.u32]
        L.tgo_procedural_voxel_is_solid.restype = T.b32
        L.tgo_procedural_solid_bits.argtypes = [T.u32, T.v3u, C.POINTER(T.u32)]
        for name, res, args in _MATH_SIGNATURES:
            fn = getattr(L, "tgo_pin_" + name)
            fn.restype, fn.argtypes = res, args
        L.tgo_pin_xorshift32_next.argtypes = [C.POINTER(T.u32)]
        L.tgo_pin_xorshift32_next.restype = T.u32
        L.tgo_pin_xorshift32_next_f32.argtypes = [C.POINTER(T.u32)]
        L.tgo_pin_xorshift32_next_f32.restype = T.f32
        L.tgo_pin_xorshift32_next_f32_range.argtypes = [C.POINTER(T.u32), T.f32, T.f32]
        L.tgo_pin_xorshift32_next_f32_range.restype = T.f32
        L.tgo_pin_intersect_ray_aabb_c.argtypes = [T.v3, T.v3, T.v3, T.v3, C.POINTER(T.f32), C.POINTER(T.f32)]
        L.tgo_pin_intersect_ray_aabb_c.restype = T.b32
        _LIB = L
    return _LIB


# the tgo_math.h routines exported for the pins and their twins in the reference's math/tg_math.c (tgm_<name>)
_MATH_SIGNATURES = [
    ("m4_mul", T.m4, [T.m4, T.m4]), ("m4_inverse", T.m4, [T.m4]), ("m4_angle_axis", T.m4, [T.f32, T.v3]), ("m4_euler", T.m4, [T.f32, T.f32, T.f32]),
    ("m4_perspective", T.m4, [T.f32, T.f32, T.f32, T.f32]), ("m4_translate", T.m4, [T.v3]), ("m4_mulv4", T.v4, [T.m4, T.v4]),
    ("v3_normalized", T.v3, [T.v3]), ("v3_lerp", T.v3, [T.v3, T.v3, T.f32]),
]


def ref():
    """The reference's own portable C (oracle/_ref/libtg_ref.so: math/tg_math.c, physics/tg_physics.c, util/tg_amanatides_woo.c,
    graphics/tg_sparse_voxel_octree.c), or None when it was never built (no reference tree at build time)."""
    global _REF
    if _REF is None:
        build()
        path = os.path.join(_HERE, "_ref", "libtg_ref.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.tg_amanatides_woo.argtypes = [T.v3, T.v3, T.v3, C.POINTER(T.u32), C.POINTER(T.v3i)]
        R.tg_amanatides_woo.restype = T.b32
        for name, res, args in _MATH_SIGNATURES:
            fn = getattr(R, "tgm_" + name)
            fn.restype, fn.argtypes = res, args
        R.tgm_simplex_noise.argtypes = [T.f32, T.f32, T.f32]
        R.tgm_simplex_noise.restype = T.f32
        R.tgm_rand_xorshift32_next_u32.argtypes = [C.POINTER(T.u32)]
        R.tgm_rand_xorshift32_next_u32.restype = T.u32
        R.tgm_rand_xorshift32_next_f32.argtypes = [C.POINTER(T.u32)]
        R.tgm_rand_xorshift32_next_f32.restype = T.f32
        R.tgm_rand_xorshift32_next_f32_inclusive_range.argtypes = [C.POINTER(T.u32), T.f32, T.f32]
        R.tgm_rand_xorshift32_next_f32_inclusive_range.restype = T.f32
        R.tg_intersect_ray_aabb.argtypes = [T.v3, T.v3, T.v3, T.v3, C.POINTER(T.f32), C.POINTER(T.f32)]
        R.tg_intersect_ray_aabb.restype = T.b32
        R.tg_intersect_aabb_obb_ignore_contact.argtypes = [T.v3, T.v3, C.POINTER(T.v3)]
        R.tg_intersect_aabb_obb_ignore_contact.restype = T.b32
        R.tg_svo_create.argtypes = [T.v3, T.v3, C.POINTER(T.tg_scene), C.POINTER(T.tg_svo)]
        R.tg_svo_destroy.argtypes = [C.POINTER(T.tg_svo)]
        R.tg_svo_traverse.argtypes = [C.POINTER(T.tg_svo), T.v3, T.v3, C.POINTER(T.f32), C.POINTER(T.u32), C.POINTER(T.u32)]
        R.tg_svo_traverse.restype = T.b32
        _REF = R
    return _REF


def ref_aw():
    """Kept name: the library that holds the reference's tg_amanatides_woo."""
    return ref()


def procedural_solid_bits(object_idx, dims):
    """tgvk_raytracer.c:871-943 restated (oracle/tgo_procedural.c): [n_clusters, 16] uint32 in pointer order."""
    n = dims[0] * dims[1] * dims[2]
    out = np.zeros((n, 16), dtype=np.uint32)
    lib().tgo_procedural_solid_bits(object_idx, T.v3u(*dims), T.ptr(out, T.u32))
    return out


def ref_scene(view):
    """A tg_scene (tgvk_raytracer.h:90-108) over the arrays of a SceneView, as the reference's tg_svo_create reads it."""
    s = T.tg_scene()
    s.object_capacity = len(view.voxel_objects)
    s.n_objects = len(view.voxel_objects)
    s.p_objects = view.voxel_objects.ctypes.data_as(C.POINTER(T.tg_voxel_object))
    s.cluster_pointer_capacity = len(view.cluster_pointers)
    s.n_cluster_pointers = len(view.cluster_pointers)
    s.p_cluster_pointers = T.ptr(view.cluster_pointers, T.u32)
    s.p_voxel_cluster_data = T.ptr(view.masks, T.u32)
    s.p_cluster_idx_to_object_idx = T.ptr(view.c2o, T.u32)
    return s


def camera_rays(cam):
    out = T.tg_camera_rays()
    lib().tgo_camera_rays(C.byref(cam), C.byref(out))
    return out


def camera_from_spec(spec):
    return T.make_camera(spec.position, spec.pitch, spec.yaw, spec.roll, spec.fov_y_deg, spec.aspect, spec.near, spec.far)


class SceneView:
    """Owns the numpy arrays behind a tgo_scene_view."""

    def __init__(self, objects, cluster_pointers, c2o, masks, lut_idx=None, color_lut=None, object_lut_idx=None, global_pointer_base=0, **_):
        L = lib()
        self.voxel_objects = np.ascontiguousarray(objects)
        n_obj = len(self.voxel_objects)
        self.object_data = np.zeros(n_obj, dtype=T.OBJECT_DATA_DTYPE)
        vo = self.voxel_objects.ctypes.data_as(C.POINTER(T.tg_voxel_object))
        od = self.object_data.ctypes.data_as(C.POINTER(T.tg_object_data))
        for i in range(n_obj):
            li = int(object_lut_idx[i]) if object_lut_idx is not None else 0
            L.tgo_object_data(C.byref(vo[i]), li, C.byref(od[i]))
        self.cluster_pointers = np.ascontiguousarray(cluster_pointers, dtype=np.uint32)
        self.c2o = np.ascontiguousarray(c2o, dtype=np.uint32)
        self.masks = np.ascontiguousarray(masks, dtype=np.uint32)
        self.lut_idx = None if lut_idx is None else np.ascontiguousarray(lut_idx, dtype=np.uint8)
        self.color_lut = None if color_lut is None else np.ascontiguousarray(color_lut, dtype=np.uint32)
        v = tgo_scene_view()
        v.n_objects_capacity = n_obj
        v.p_objects = od
        v.n_cluster_pointers = len(self.cluster_pointers)
        v.p_cluster_pointers = T.ptr(self.cluster_pointers, T.u32)
        v.p_cluster_idx_to_object_idx = T.ptr(self.c2o, T.u32)
        v.p_voxel_cluster_data = T.ptr(self.masks, T.u32)
        v.p_color_lut_idx_data = T.ptr(self.lut_idx, T.u8) if self.lut_idx is not None else None
        v.p_color_lut = T.ptr(self.color_lut, T.u32) if self.color_lut is not None else None
        v.global_pointer_base = global_pointer_base
        self.view = v

    @classmethod
    def from_scene(cls, scene, global_pointer_base=0, with_lut=True):
        from tg_b200.scenes import flat_arrays
        return cls(**flat_arrays(scene, global_pointer_base, with_lut))


def visibility(view, rays, w, h, mode=VIS_SCREEN_RECT, y0=0, y1=None, ystep=1):
    out = np.empty(w * h, dtype=np.uint64)
    n = lib().tgo_visibility(C.byref(view.view), C.byref(rays), w, h, mode, y0, h if y1 is None else y1, ystep, T.ptr(out, T.u64))
    return out.reshape(h, w), int(n)


def visibility_window(view, rays, w, h, x0, x1, y0, y1):
    """Unpruned brute force (every pixel of the window x every cluster pointer) -> (words [h, w] with the clear value outside, fragments)."""
    out = np.empty(w * h, dtype=np.uint64)
    n = lib().tgo_visibility_window(C.byref(view.view), C.byref(rays), w, h, x0, x1, y0, y1, T.ptr(out, T.u64))
    return out.reshape(h, w), int(n)


def visibility_svo(svo, rays, w, h, y0=0, y1=None, ystep=1):
    """debug_visibility_svo.frag: the BLOCKS view's primary rays through the SVO -> u64 words [h, w]."""
    out = np.empty(w * h, dtype=np.uint64)
    lib().tgo_visibility_svo(C.byref(svo), C.byref(rays), w, h, y0, h if y1 is None else y1, ystep, T.ptr(out, T.u64))
    return out.reshape(h, w)


def visibility_fragment(view, rays, w, h, px, py, cluster_pointer):
    return int(lib().tgo_visibility_fragment(C.byref(view.view), C.byref(rays), w, h, px, py, cluster_pointer))


def svo_create(view, extent_min=(-512.0, -512.0, -512.0), extent_max=(512.0, 512.0, 512.0), capacities=None):
    svo = T.tg_svo()
    if capacities:
        svo.voxel_buffer_capacity_in_u32, svo.leaf_node_data_buffer_capacity, svo.node_buffer_capacity = capacities
    vo = view.voxel_objects.ctypes.data_as(C.POINTER(T.tg_voxel_object))
    lib().tgo_svo_create(T.v3(*extent_min), T.v3(*extent_max), C.byref(view.view), vo, C.byref(svo))
    return svo


def svo_arrays(svo):
    """(nodes u32[count], leaf_data u32[count, 65], voxels u32[count_in_u32]) copies."""
    nodes = np.ctypeslib.as_array(svo.p_node_buffer, shape=(svo.node_buffer_count,)).copy()
    nl = svo.leaf_node_data_buffer_count
    leaf = np.ctypeslib.as_array(C.cast(svo.p_leaf_node_data_buffer, C.POINTER(T.u32)), shape=(max(nl, 1), 65))[:nl].copy()
    nv = svo.voxel_buffer_count_in_u32
    vox = np.ctypeslib.as_array(svo.p_voxels_buffer, shape=(max(nv, 1),))[:nv].copy()
    return nodes, leaf, vox


def svo_destroy(svo):
    lib().tgo_svo_destroy(C.byref(svo))


def shade(view, rays, w, h, vis, svo=None, gi=False, frame_seed=1, debug=0, y0=0, y1=None, out=None, ystep=1):
    if out is None:
        out = np.zeros((h, w, 4), dtype=np.float32)
    vis = np.ascontiguousarray(vis, dtype=np.uint64)
    lib().tgo_shade(C.byref(view.view), C.byref(rays), w, h, T.ptr(vis, T.u64), C.byref(svo) if svo is not None else None,
                    1 if gi else 0, frame_seed, debug, y0, h if y1 is None else y1, ystep, T.ptr(out, T.f32))
    return out


def gi_rays(view, rays, w, h, vis, svo, frame_seed=1, y0=0, y1=None, ystep=1):
    """the secondary rays shade(..., gi=True) traces: (origins [n, 3], directions [n, 3]) of the pixels of those rows that shoot one"""
    out = np.zeros((h, w, 6), dtype=np.float32)
    vis = np.ascontiguousarray(vis, dtype=np.uint64)
    lib().tgo_shade_gi_rays(C.byref(view.view), C.byref(rays), w, h, T.ptr(vis, T.u64), C.byref(svo), frame_seed, y0, h if y1 is None else y1, ystep, T.ptr(out, T.f32))
    out = out.reshape(-1, 6)
    out = out[(out[:, 3:] != 0).any(axis=1)]
    return np.ascontiguousarray(out[:, :3]), np.ascontiguousarray(out[:, 3:])


def present(radiance):
    """present.frag + B8G8R8A8_UNORM conversion: float32 [..., 4] -> uint32 [...] (a << 24 | r << 16 | g << 8 | b)."""
    rad = np.ascontiguousarray(radiance, dtype=np.float32)
    out = np.empty(rad.shape[:-1], dtype=np.uint32)
    lib().tgo_present_bgra8(T.ptr(rad, T.f32), out.size, T.ptr(out, T.u32))
    return out

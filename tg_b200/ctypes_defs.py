"""ctypes mirrors of include/tg_types.h and include/tg_raytracer.h (plain C structs, no torch types).

Layouts follow the reference (paths under /root/reference/tg/src): math/tg_math.h:72-275 (v3, m4
column-major), graphics/tg_graphics_core.h:111-150 (tg_camera, tg_voxel_object),
graphics/tg_sparse_voxel_octree.h:6-47 (tg_svo*), graphics/vulkan/tgvk_raytracer.h:90-108 (tg_scene),
graphics/vulkan/tgvk_raytracer.c:35-42,65-74 (object record, camera block).
"""
import ctypes as C

import numpy as np

b32, f32, u8, u16, u32, u64, i32 = C.c_int32, C.c_float, C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64, C.c_int32

TG_VIS_CLEAR = 0xFFFFFFFFFFFFFFFF
TG_CLUSTER_MASK_WORDS = 16
TG_SVO_BLOCK_WORDS = 1024
TG_SVO_LEAF_MAX_CLUSTERS = 64
TG_U32_MAX = 0xFFFFFFFF


class v3(C.Structure):
    _fields_ = [("x", f32), ("y", f32), ("z", f32)]

    def __init__(self, x=0.0, y=0.0, z=0.0):
        super().__init__(float(x), float(y), float(z))

    def tuple(self):
        return (self.x, self.y, self.z)


class v3u(C.Structure):
    _fields_ = [("x", u32), ("y", u32), ("z", u32)]

    def tuple(self):
        return (self.x, self.y, self.z)


class v3i(C.Structure):
    _fields_ = [("x", i32), ("y", i32), ("z", i32)]


class v4(C.Structure):
    _fields_ = [("x", f32), ("y", f32), ("z", f32), ("w", f32)]


class m4(C.Structure):
    # column-major: m<row><col>
    _fields_ = [(n, f32) for n in ("m00", "m10", "m20", "m30", "m01", "m11", "m21", "m31",
                                   "m02", "m12", "m22", "m32", "m03", "m13", "m23", "m33")]


class _ortho(C.Structure):
    _fields_ = [(n, f32) for n in ("l", "r", "b", "t", "n", "f")]


class _persp(C.Structure):
    _fields_ = [(n, f32) for n in ("fov_y_in_radians", "aspect", "n", "f")]


class _camera_union(C.Union):
    _fields_ = [("ortho", _ortho), ("persp", _persp)]


TG_CAMERA_TYPE_ORTHOGRAPHIC = 0
TG_CAMERA_TYPE_PERSPECTIVE = 1


class tg_camera(C.Structure):
    _anonymous_ = ("u",)
    _fields_ = [("type", C.c_int), ("position", v3), ("pitch", f32), ("yaw", f32), ("roll", f32), ("u", _camera_union)]


class tg_voxel_object(C.Structure):
    _fields_ = [("n_cluster_pointers_per_dim", v3u), ("first_cluster_pointer", u32), ("translation", v3),
                ("angle_in_radians", f32), ("axis", v3)]


class tg_object_data(C.Structure):
    _fields_ = [("n_cluster_pointers_per_dim", v3u), ("first_cluster_pointer", u32), ("translation", v3),
                ("lut_idx", u32), ("rotation", m4)]


class tg_camera_rays(C.Structure):
    _fields_ = [("camera", v4), ("ray_bl", v4), ("ray_br", v4), ("ray_tr", v4), ("ray_tl", v4),
                ("near_plane", f32), ("far_plane", f32), ("pad", f32 * 2)]


class tg_svo_leaf_node_data(C.Structure):
    _fields_ = [("n", u32), ("p_cluster_idcs", u32 * TG_SVO_LEAF_MAX_CLUSTERS)]


class tg_svo(C.Structure):
    _fields_ = [("min", v3), ("max", v3),
                ("voxel_buffer_capacity_in_u32", u32), ("voxel_buffer_count_in_u32", u32),
                ("leaf_node_data_buffer_capacity", u32), ("leaf_node_data_buffer_count", u32),
                ("node_buffer_capacity", u32), ("node_buffer_count", u32),
                ("p_voxels_buffer", C.POINTER(u32)),
                ("p_leaf_node_data_buffer", C.POINTER(tg_svo_leaf_node_data)),
                ("p_node_buffer", C.POINTER(u32))]


class tg_scene(C.Structure):
    _fields_ = [("object_capacity", u32), ("n_objects", u32), ("p_objects", C.POINTER(tg_voxel_object)),
                ("n_available_object_indices", u32), ("p_available_object_indices", C.POINTER(u32)),
                ("cluster_pointer_capacity", u32), ("n_cluster_pointers", u32), ("p_cluster_pointers", C.POINTER(u32)),
                ("n_available_cluster_indices", u32), ("p_available_cluster_indices", C.POINTER(u32)),
                ("p_voxel_cluster_data", C.POINTER(u32)), ("p_cluster_idx_to_object_idx", C.POINTER(u32)),
                ("svo", tg_svo)]


class tg_raytracer(C.Structure):
    _fields_ = [("p_camera", C.POINTER(tg_camera)), ("scene", tg_scene), ("p_device", C.c_void_p),
                ("width", u32), ("height", u32), ("debug_visualization", u32), ("gi_enabled", u32),
                ("frame_seed", u32), ("svo_dirty", u32), ("p_object_lut_idx", C.POINTER(u32)), ("n_color_luts", u32),
                ("n_moved_objects", u32), ("p_moved_objects", C.POINTER(u32))]


class tgb200_timings(C.Structure):
    _fields_ = [("clear_ms", f32), ("cull_ms", f32), ("visibility_ms", f32), ("svo_ms", f32), ("shading_ms", f32),
                ("merge_ms", f32), ("n_visible_objects", u32), ("n_kernel_launches", u32), ("n_gi_rays", u32), ("n_gi_rays_exact", u32),
                ("n_gi_node_visits", u64), ("n_gi_dda_steps", u64), ("n_gi_advances", u64),
                ("merge_resolve_ms", f32), ("merge_gather_ms", f32), ("merge_kernel_ms", f32), ("pad2", u32)]


OBJECT_DATA_DTYPE = np.dtype([("dims", "<u4", 3), ("first_cluster_pointer", "<u4"), ("translation", "<f4", 3),
                              ("lut_idx", "<u4"), ("rotation", "<f4", 16)])
assert OBJECT_DATA_DTYPE.itemsize == 96 == C.sizeof(tg_object_data)
VOXEL_OBJECT_DTYPE = np.dtype([("dims", "<u4", 3), ("first_cluster_pointer", "<u4"), ("translation", "<f4", 3),
                               ("angle_in_radians", "<f4"), ("axis", "<f4", 3)])
assert VOXEL_OBJECT_DTYPE.itemsize == 44 == C.sizeof(tg_voxel_object)
assert C.sizeof(tg_svo_leaf_node_data) == 260
assert C.sizeof(tg_camera) == 52
assert C.sizeof(tg_camera_rays) == 96


def ptr(arr, ctype):
    """Typed pointer to a C-contiguous numpy array (kept alive by the caller)."""
    assert arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(C.POINTER(ctype))


def make_camera(position, pitch, yaw, roll, fov_y_deg, aspect, near, far):
    cam = tg_camera()
    cam.type = TG_CAMERA_TYPE_PERSPECTIVE
    cam.position = v3(*position)
    cam.pitch, cam.yaw, cam.roll = pitch, yaw, roll
    # TG_DEG2RAD in float32, like the reference (tg_application.c:56)
    cam.persp.fov_y_in_radians = float(np.float32(fov_y_deg) * (np.float32(3.14159265358979323846) / np.float32(180.0)))
    cam.persp.aspect = aspect
    cam.persp.n = near
    cam.persp.f = far
    return cam

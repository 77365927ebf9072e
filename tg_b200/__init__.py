"""tg_b200 -- B200-native voxel visibility / SVO / GI path behind TG's raytracer C API.

The product is `libtgb200.so` (C host code + hand-written sm_100a CUDA, `tg_b200/csrc/`, C ABI declared in
`include/tg_raytracer.h`). This package is only the Python binding a test / bench driver needs: it loads the
library with ctypes and mirrors the reference's raytracer interface (`Raytracer`). There is no CPU fallback:
if the library is missing the import of `tg_b200.lib()` raises, and without a CUDA device
`Raytracer(...)` raises with the library's own error string.
"""
import ctypes as C
import os
import subprocess

from . import ctypes_defs as T

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtgb200.so")
_LIB = None

# every symbol include/tg_raytracer.h declares: name -> (restype, argtypes)
_P = C.POINTER
_RT = _P(T.tg_raytracer)
SYMBOLS = {
    "tg_raytracer_create": (None, [_P(T.tg_camera), T.u32, T.u32, _RT]),
    "tg_raytracer_destroy": (None, [_RT]),
    "tg_raytracer_set_debug_visualization": (None, [_RT, C.c_int]),
    "tg_raytracer_create_object": (None, [_RT, T.v3, T.v3u]),
    "tg_raytracer_destroy_object": (None, [_RT, T.u32]),
    "tg_object_is_initialized": (T.b32, [_P(T.tg_scene), T.u32]),
    "tg_raytracer_color_lut_set": (None, [_RT, T.u8, T.f32, T.f32, T.f32]),
    "tg_raytracer_render": (None, [_RT]),
    "tg_raytracer_clear": (None, [_RT]),
    "tg_raytracer_get_hovered_voxel": (T.b32, [_RT, T.u32, T.u32, _P(T.f32), _P(T.u32), _P(T.u32)]),
    "tg_svo_create": (None, [T.v3, T.v3, _P(T.tg_scene), _P(T.tg_svo)]),
    "tg_svo_destroy": (None, [_P(T.tg_svo)]),
    "tg_svo_traverse": (T.b32, [_P(T.tg_svo), T.v3, T.v3, _P(T.f32), _P(T.u32), _P(T.u32)]),
    "tgb200_last_error": (C.c_char_p, []),
    "tgb200_clear_error": (None, []),
    "tgb200_device_count": (T.i32, []),
    "tgb200_set_device": (None, [T.i32]),
    "tgb200_set_default_resolution": (None, [T.u32, T.u32]),
    "tg_raytracer_set_resolution": (None, [_RT, T.u32, T.u32]),
    "tg_raytracer_create_object_from_data": (T.u32, [_RT, T.v3, T.v3u, T.f32, T.v3, T.u32, _P(T.u32), _P(T.u8)]),
    "tg_raytracer_create_object_synthetic": (T.u32, [_RT, T.v3, T.v3u, T.f32, T.v3, T.u32, T.u32, T.u32]),
    "tgb200_synthetic_solid_bits": (None, [T.u32, T.u32, T.u32, _P(T.u32)]),
    "tg_raytracer_set_object_transform": (None, [_RT, T.u32, T.v3, T.f32, T.v3]),
    "tg_raytracer_color_lut_set_ex": (None, [_RT, T.u32, T.u8, T.f32, T.f32, T.f32]),
    "tg_raytracer_set_gi": (None, [_RT, T.b32, T.u32]),
    "tgb200_set_gi_traversal": (None, [_RT, T.u32]),
    "tgb200_set_frame_sink": (None, [_RT, C.c_void_p, T.u32]),
    "tgb200_set_frame_sink_ex": (None, [_RT, C.c_void_p, T.u32, C.c_int]),
    "tg_raytracer_read_present": (None, [_RT, _P(T.u32)]),
    "tgb200_save_frame_bmp": (T.b32, [_RT, C.c_char_p]),
    "tgb200_write_bmp_bgra8": (T.b32, [C.c_char_p, T.u32, T.u32, _P(T.u32)]),
    "tgb200_scene_save": (T.b32, [_RT, C.c_char_p]),
    "tgb200_scene_load": (T.b32, [_RT, C.c_char_p]),
    "tgb200_frame_ticket": (T.u64, [_RT]),
    "tgb200_wait_frame": (None, [_RT, T.u64]),
    "tgb200_render_visibility": (None, [_RT]),
    "tgb200_svo_update": (None, [_RT, T.b32]),
    "tgb200_svo_leaves_resampled": (T.u32, [_RT]),
    "tgb200_render_shading": (None, [_RT]),
    "tgb200_render_shading_rows": (None, [_RT, T.u32, T.u32]),
    "tgb200_synchronize": (None, [_RT]),
    "tg_raytracer_read_visibility": (None, [_RT, _P(T.u64)]),
    "tg_raytracer_read_radiance": (None, [_RT, _P(T.f32)]),
    "tg_raytracer_read_radiance_rows": (None, [_RT, T.u32, T.u32, _P(T.f32)]),
    "tg_raytracer_write_visibility": (None, [_RT, _P(T.u64)]),
    "tgb200_svo_download": (None, [_RT, _P(T.tg_svo)]),
    "tgb200_svo_upload": (None, [_RT, _P(T.tg_svo)]),
    "tgb200_get_timings": (None, [_RT, _P(T.tgb200_timings)]),
    "tgb200_reset_launch_counter": (None, [_RT]),
    "tgb200_device_visibility": (C.c_void_p, [_RT]),
    "tgb200_device_radiance": (C.c_void_p, [_RT]),
    "tgb200_stream": (C.c_void_p, [_RT]),
    "tgb200_set_shard": (None, [_RT, T.u32, T.u32, T.u32]),
    "tgb200_comm_unique_id": (None, [_P(T.u8)]),
    "tgb200_comm_init": (None, [_RT, _P(T.u8), T.u32, T.u32]),
    "tgb200_comm_destroy": (None, [_RT]),
    "tgb200_merge_visibility": (None, [_RT]),
    "tgb200_set_merge_kind": (None, [_RT, T.u32]),
    "tgb200_tile_rows": (None, [_RT, _P(T.u32), _P(T.u32)]),
    "tgb200_tile_physical_row": (T.u32, [_RT, T.u32]),
    "tgb200_gather_radiance": (None, [_RT]),
    "tgb200_mark_svo_dirty": (None, [_RT]),
    "tgb200_scene_init": (None, [_P(T.tg_scene), T.u32, T.u32]),
    "tgb200_scene_free": (None, [_P(T.tg_scene)]),
    "tgb200_scene_alloc_object": (T.u32, [_P(T.tg_scene), T.v3, T.v3u, T.f32, T.v3]),
    "tgb200_scene_free_object": (None, [_P(T.tg_scene), T.u32, _P(T.u32), _P(T.u32)]),
    "tgb200_camera_rays": (None, [_P(T.tg_camera), _P(T.tg_camera_rays)]),
    "tgb200_object_data": (None, [_P(T.tg_scene), T.u32, T.u32, _P(T.tg_object_data)]),
    "tgb200_pack_color": (T.u32, [T.f32, T.f32, T.f32]),
    "tgb200_debug_cluster_ray": (None, [_P(T.tg_object_data), _P(T.tg_camera_rays), T.u32, T.u32, T.u32, T.u32, T.u32, _P(T.v3), _P(T.v3)]),
    "tgb200_procedural_solid_bits": (None, [T.u32, T.v3u, _P(T.u32)]),
}


class TgError(RuntimeError):
    pass


def build(verbose=False):
    """Compile libtgb200.so in-tree (nvcc -gencode arch=compute_100a,code=sm_100a; no GPU needed)."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise TgError("building libtgb200.so failed")


def lib():
    """The loaded C-ABI library. Raises (loudly) when it is not built: there is no fallback path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise TgError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA library is the only implementation; there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError == a declared symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        _LIB = L
    return _LIB


def check():
    """Raise TgError if the library recorded an error since the last check."""
    err = lib().tgb200_last_error()
    if err:
        msg = err.decode()
        lib().tgb200_clear_error()
        raise TgError(msg)


from .raytracer import Raytracer  # noqa: E402,F401

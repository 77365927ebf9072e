"""TEST INFRASTRUCTURE: host builds of the per-ray state machines the CUDA kernels run (tg_b200/csrc/*.cuh), driven ray by ray
without a GPU. Built on demand with g++ (no FMA contraction, like the product)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "tg_b200", "csrc")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libtgbsim.so")
        srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".cpp")]
        deps = srcs + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".h", ".cuh"))]
        if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-Wall", "-o", so] + srcs)
        L = C.CDLL(so)
        u32p, f32p = C.POINTER(C.c_uint32), C.POINTER(C.c_float)
        L.tgbsim_flatten.argtypes = [u32p, u32p, C.c_uint32, C.c_uint32, u32p]
        L.tgbsim_gi_trace.argtypes = [f32p, f32p, C.c_float, u32p, u32p, C.c_uint32, f32p, f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)]
        L.tgbsim_gi_trace.restype = C.c_uint32
        L.tgbsim_gi_fast.argtypes = [f32p, f32p, C.c_float, u32p, u32p, C.c_uint32, f32p, f32p, C.c_uint32, C.c_float, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)]
        L.tgbsim_fast_tiling.argtypes = [u32p, u32p, C.c_uint32, u32p, u32p, u32p]
        L.tgbsim_gi_fast_tiled.argtypes = [f32p, f32p, C.c_float, u32p, u32p, u32p, u32p, u32p, C.c_uint32, C.c_uint32, f32p, f32p, C.c_uint32, C.c_float, C.c_uint32,
                                           C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), u32p]
        L.tgbsim_svo_traverse.argtypes = [u32p, u32p, u32p, f32p, f32p, C.c_float, C.c_uint32, f32p, f32p, f32p, u32p, u32p, C.POINTER(C.c_uint64)]
        L.tgbsim_visibility.argtypes = [C.c_void_p, C.c_uint32, u32p, u32p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                        C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def flatten(nodes, leaf_data):
    grid = np.zeros(32 ** 3 + 1, dtype=np.uint32)
    lib().tgbsim_flatten(_p(nodes, C.c_uint32), _p(leaf_data, C.c_uint32), len(nodes), len(leaf_data) // 65, _p(grid, C.c_uint32))
    return grid


def gi_trace(bmin, bmax, far_plane, grid, voxels, origins, dirs, tree_reps=4, dda_steps=16, grid16=False, uniform_dda=False):
    origins = np.ascontiguousarray(origins, dtype=np.float32)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32)
    bmin = np.asarray(bmin, dtype=np.float32); bmax = np.asarray(bmax, dtype=np.float32)
    occluded = np.zeros(len(origins), dtype=np.uint8)
    work = np.zeros(3, dtype=np.uint64)
    capped = lib().tgbsim_gi_trace(_p(bmin, C.c_float), _p(bmax, C.c_float), far_plane, _p(grid, C.c_uint32), _p(voxels, C.c_uint32), len(origins),
                                   _p(origins, C.c_float), _p(dirs, C.c_float), tree_reps, dda_steps, (1 if grid16 else 0) | (2 if uniform_dda else 0), _p(occluded, C.c_uint8), _p(work, C.c_uint64))
    return occluded.astype(bool), int(capped), work


def gi_fast(bmin, bmax, far_plane, grid, voxels, origins, dirs, steps=8, delta=1.0e-3):
    """the certified fast walk (tgb_gi_fast.cuh) per ray -> (result u8: 0 unoccluded / 1 occluded / 2 handed to the exact kernel, work[3])."""
    origins = np.ascontiguousarray(origins, dtype=np.float32)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32)
    bmin = np.asarray(bmin, dtype=np.float32); bmax = np.asarray(bmax, dtype=np.float32)
    result = np.zeros(len(origins), dtype=np.uint8)
    work = np.zeros(3, dtype=np.uint64)
    lib().tgbsim_gi_fast(_p(bmin, C.c_float), _p(bmax, C.c_float), far_plane, _p(grid, C.c_uint32), _p(voxels, C.c_uint32), len(origins),
                         _p(origins, C.c_float), _p(dirs, C.c_float), steps, delta, _p(result, C.c_uint8), _p(work, C.c_uint64))
    return result, work


def fast_tiling(grid, voxels):
    """the coarser tiling of the certified fast walk (tgb_gi_fast.cuh): (cells u32[32^3], bricks u32[64 * n_leaves], columns u32[2 * 1024 * n_leaves])"""
    n_leaves = len(voxels) // 1024
    cells = np.zeros(32 ** 3, dtype=np.uint32)
    bricks = np.zeros(max(1, 64 * n_leaves), dtype=np.uint32)
    columns = np.zeros(max(1, 2 * 1024 * n_leaves), dtype=np.uint32)
    lib().tgbsim_fast_tiling(_p(grid, C.c_uint32), _p(voxels, C.c_uint32), n_leaves, _p(cells, C.c_uint32), _p(bricks, C.c_uint32), _p(columns, C.c_uint32))
    return cells, bricks, columns


def gi_fast_tiled(bmin, bmax, far_plane, grid, voxels, tiling, origins, dirs, steps=8, delta=1.0e-3, want_steps=False, cube=False):
    """gi_fast over the coarser tiling -> (result, work[3]: boxes of free cells entered, cells of leaf blocks entered, rays handed over[, steps per ray]);
    cube: with the cube check of near-edge steps and without step caps, as the second stage (k_gi_trace_list) walks the rays the bulk kernels hand over"""
    origins = np.ascontiguousarray(origins, dtype=np.float32)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32)
    bmin = np.asarray(bmin, dtype=np.float32); bmax = np.asarray(bmax, dtype=np.float32)
    result = np.zeros(len(origins), dtype=np.uint8)
    work = np.zeros(3, dtype=np.uint64)
    per_ray = np.zeros(len(origins), dtype=np.uint32)
    lib().tgbsim_gi_fast_tiled(_p(bmin, C.c_float), _p(bmax, C.c_float), far_plane, _p(grid, C.c_uint32), _p(voxels, C.c_uint32), _p(tiling[0], C.c_uint32),
                               _p(tiling[1], C.c_uint32), _p(tiling[2], C.c_uint32), len(voxels) // 1024, len(origins), _p(origins, C.c_float), _p(dirs, C.c_float), steps, delta, 1 if cube else 0, _p(result, C.c_uint8),
                               _p(work, C.c_uint64), _p(per_ray, C.c_uint32))
    return (result, work, per_ray) if want_steps else (result, work)


def svo_traverse(nodes, leaf_data, voxels, bmin, bmax, far_plane, origins, dirs):
    """tgb_svo_traverse_stack per ray -> (result f32, node idx, voxel idx, packed BLOCKS-view word)."""
    origins = np.ascontiguousarray(origins, dtype=np.float32)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32)
    bmin = np.asarray(bmin, dtype=np.float32); bmax = np.asarray(bmax, dtype=np.float32)
    n = len(origins)
    res, node, vox, word = np.zeros(n, np.float32), np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint64)
    lib().tgbsim_svo_traverse(_p(nodes, C.c_uint32), _p(leaf_data, C.c_uint32), _p(voxels, C.c_uint32), _p(bmin, C.c_float), _p(bmax, C.c_float), far_plane, n,
                              _p(origins, C.c_float), _p(dirs, C.c_float), _p(res, C.c_float), _p(node, C.c_uint32), _p(vox, C.c_uint32), _p(word, C.c_uint64))
    return res, node, vox, word


def visibility(view, rays, w, h, y0=0, y1=None, ystep=1, defer=False):
    """K1's per-ray walk (tgb_k1_walk.cuh) over an oracle SceneView on the host -> (u64 words [h, w], [candidates marched, set-ups])."""
    out = np.empty(w * h, dtype=np.uint64)
    work = np.zeros(3, dtype=np.uint64)
    v = view.view
    lib().tgbsim_visibility(C.cast(v.p_objects, C.c_void_p), v.n_objects_capacity, v.p_cluster_pointers, v.p_voxel_cluster_data, C.cast(C.pointer(rays), C.c_void_p), w, h,
                            v.global_pointer_base, y0, h if y1 is None else y1, ystep, 1 if defer else 0, _p(out, C.c_uint64), _p(work, C.c_uint64))
    return out.reshape(h, w), work

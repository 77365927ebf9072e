"""Host-side mirror of the reference's raytracer interface over the C ABI (include/tg_raytracer.h).

Method names and argument meaning follow /root/reference/tg/src/graphics/vulkan/tgvk_raytracer.h:240-251
(`create_object`, `destroy_object`, `color_lut_set`, `clear`, `render`, `get_hovered_voxel`) so that tests
read like a TG application (tg_application.c:64-98,282,376). Every call goes through libtgb200.so.
"""
import ctypes as C

import numpy as np

from . import ctypes_defs as T


class Raytracer:
    def __init__(self, camera, max_n_objects, max_n_clusters, width, height, device=0):
        from . import check, lib
        self._lib = lib()
        self.camera = camera  # borrowed by the C side: keep it alive, mutate it in place to move the camera
        self._rt = T.tg_raytracer()
        self._lib.tgb200_clear_error()
        self._lib.tgb200_set_device(device)
        self._lib.tgb200_set_default_resolution(width, height)
        self._lib.tg_raytracer_create(C.byref(self.camera), max_n_objects, max_n_clusters, C.byref(self._rt))
        check()
        self.width, self.height = width, height
        self._alive = True

    # ---- lifetime -------------------------------------------------------------------------------
    def destroy(self):
        if getattr(self, "_alive", False):
            self._lib.tg_raytracer_destroy(C.byref(self._rt))
            self._alive = False

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.destroy()

    def _check(self):
        from . import check
        check()

    @property
    def scene(self):
        return self._rt.scene

    # ---- reference entry points -----------------------------------------------------------------
    def create_object(self, center, extent):
        self._lib.tg_raytracer_create_object(C.byref(self._rt), T.v3(*center), T.v3u(*extent))
        self._check()

    def create_object_from_data(self, center, extent, angle, axis, solid_bits, lut_indices=None, lut_idx=0):
        bits = np.ascontiguousarray(solid_bits, dtype=np.uint32)
        n = (extent[0] // 8) * (extent[1] // 8) * (extent[2] // 8)
        assert bits.size == n * 16, (bits.shape, n)
        lut_ptr = None
        if lut_indices is not None:
            lut_indices = np.ascontiguousarray(lut_indices, dtype=np.uint8)
            assert lut_indices.size == n * 512
            lut_ptr = T.ptr(lut_indices, T.u8)
        idx = self._lib.tg_raytracer_create_object_from_data(C.byref(self._rt), T.v3(*center), T.v3u(*extent), angle, T.v3(*axis), lut_idx,
                                                             T.ptr(bits, T.u32), lut_ptr)
        self._check()
        return idx

    def create_object_synthetic(self, center, extent, angle, axis, object_seed, k, lut_idx=0):
        """Seeded random bits generated on the device (scenes.random_solid_bits(object_seed, n, k) without the host array)."""
        idx = self._lib.tg_raytracer_create_object_synthetic(C.byref(self._rt), T.v3(*center), T.v3u(*extent), angle, T.v3(*axis), lut_idx, object_seed, k)
        self._check()
        return idx

    def destroy_object(self, object_idx):
        self._lib.tg_raytracer_destroy_object(C.byref(self._rt), object_idx)
        self._check()

    def set_object_transform(self, object_idx, translation, angle, axis=(0.0, 1.0, 0.0)):
        self._lib.tg_raytracer_set_object_transform(C.byref(self._rt), object_idx, T.v3(*translation), angle, T.v3(*axis))
        self._check()

    def color_lut_set(self, index, r, g, b, lut_idx=0):
        self._lib.tg_raytracer_color_lut_set_ex(C.byref(self._rt), lut_idx, index, r, g, b)
        self._check()

    def set_debug_visualization(self, kind):
        self._lib.tg_raytracer_set_debug_visualization(C.byref(self._rt), kind)
        self._check()

    def set_gi(self, enabled, frame_seed=1):
        self._lib.tg_raytracer_set_gi(C.byref(self._rt), 1 if enabled else 0, frame_seed)
        self._check()

    def set_resolution(self, width, height):
        self._lib.tg_raytracer_set_resolution(C.byref(self._rt), width, height)
        self._check()
        self.width, self.height = width, height

    def clear(self):
        self._lib.tg_raytracer_clear(C.byref(self._rt))

    def render(self):
        self._lib.tg_raytracer_render(C.byref(self._rt))

    def get_hovered_voxel(self, x, y):
        depth, cluster, voxel = T.f32(), T.u32(), T.u32()
        hit = self._lib.tg_raytracer_get_hovered_voxel(C.byref(self._rt), x, y, C.byref(depth), C.byref(cluster), C.byref(voxel))
        self._check()
        return bool(hit), depth.value, cluster.value, voxel.value

    # ---- stages / results -----------------------------------------------------------------------
    def render_visibility(self):
        self._lib.tgb200_render_visibility(C.byref(self._rt))

    def svo_update(self, force_full=False):
        self._lib.tgb200_svo_update(C.byref(self._rt), 1 if force_full else 0)

    def svo_leaves_resampled(self):
        return self._lib.tgb200_svo_leaves_resampled(C.byref(self._rt))

    def render_shading(self):
        self._lib.tgb200_render_shading(C.byref(self._rt))

    def synchronize(self):
        self._lib.tgb200_synchronize(C.byref(self._rt))
        self._check()

    def read_visibility(self, out=None):
        if out is None:
            out = np.empty((self.height, self.width), dtype=np.uint64)
        self._lib.tg_raytracer_read_visibility(C.byref(self._rt), T.ptr(out, T.u64))
        self._check()
        return out

    def write_visibility(self, vis):
        vis = np.ascontiguousarray(vis, dtype=np.uint64)
        assert vis.size == self.width * self.height
        self._lib.tg_raytracer_write_visibility(C.byref(self._rt), T.ptr(vis, T.u64))
        self.synchronize()

    def read_radiance(self, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._lib.tg_raytracer_read_radiance(C.byref(self._rt), T.ptr(out, T.f32))
        self._check()
        return out

    def read_radiance_rows(self, y0, y1, out=None):
        if out is None:
            out = np.empty((y1 - y0, self.width, 4), dtype=np.float32)
        self._lib.tg_raytracer_read_radiance_rows(C.byref(self._rt), y0, y1, T.ptr(out, T.f32))
        self._check()
        return out

    def svo_download(self):
        """(svo struct, nodes u32[n], leaf_data u32[n, 65], voxels u32[n_words]); struct freed with svo_free()."""
        svo = T.tg_svo()
        self._lib.tgb200_svo_download(C.byref(self._rt), C.byref(svo))
        self._check()
        nodes = np.ctypeslib.as_array(svo.p_node_buffer, shape=(max(svo.node_buffer_count, 1),))[:svo.node_buffer_count].copy()
        nl = svo.leaf_node_data_buffer_count
        leaf = np.ctypeslib.as_array(C.cast(svo.p_leaf_node_data_buffer, C.POINTER(T.u32)), shape=(max(nl, 1), 65))[:nl].copy()
        nv = svo.voxel_buffer_count_in_u32
        vox = np.ctypeslib.as_array(svo.p_voxels_buffer, shape=(max(nv, 1),))[:nv].copy()
        return svo, nodes, leaf, vox

    def svo_free(self, svo):
        self._lib.tg_svo_destroy(C.byref(svo))

    def svo_upload(self, svo):
        self._lib.tgb200_svo_upload(C.byref(self._rt), C.byref(svo))
        self._check()

    def timings(self):
        t = T.tgb200_timings()
        self._lib.tgb200_get_timings(C.byref(self._rt), C.byref(t))
        return {n: getattr(t, n) for n, _ in T.tgb200_timings._fields_}

    def reset_launch_counter(self):
        self._lib.tgb200_reset_launch_counter(C.byref(self._rt))

    # ---- multi-GPU --------------------------------------------------------------------------------
    def set_shard(self, rank, n_ranks, global_pointer_base):
        self._lib.tgb200_set_shard(C.byref(self._rt), rank, n_ranks, global_pointer_base)
        self._check()

    def comm_init(self, unique_id, rank, n_ranks):
        buf = (T.u8 * 128).from_buffer_copy(bytes(unique_id))
        self._lib.tgb200_comm_init(C.byref(self._rt), buf, rank, n_ranks)
        self._check()

    def set_merge_kind(self, kind):
        """0 = merge over peer memory when it can be mapped (default), 1 = NCCL collectives only."""
        self._lib.tgb200_set_merge_kind(C.byref(self._rt), kind)

    def merge_visibility(self):
        self._lib.tgb200_merge_visibility(C.byref(self._rt))

    def tile_rows(self):
        y0, y1 = T.u32(), T.u32()
        self._lib.tgb200_tile_rows(C.byref(self._rt), C.byref(y0), C.byref(y1))
        return y0.value, y1.value

    def tile_physical_rows(self):
        """Frame row of every row of this rank's tile, in tile order (what the frame sink delivers); -1 = padding row."""
        y0, y1 = self.tile_rows()
        rows = np.array([self._lib.tgb200_tile_physical_row(C.byref(self._rt), i) for i in range(y1 - y0)], dtype=np.int64)
        rows[rows == 0xFFFFFFFF] = -1
        return rows

    def gather_radiance(self):
        self._lib.tgb200_gather_radiance(C.byref(self._rt))

    def set_frame_sink(self, host_array, n_bands=8):
        """Every later render() copies its shaded rows into `host_array` band by band while the frame is still being shaded
        (pinned memory for overlap); None switches it off. float32 rows x w x 4 = the HDR rows; uint32 rows x w = the
        presented rows (B8G8R8A8_UNORM, present.frag + the swapchain's conversion)."""
        if host_array is None:
            self._sink = None
            self._lib.tgb200_set_frame_sink(C.byref(self._rt), None, 1)
            return
        assert host_array.dtype in (np.float32, np.uint32) and host_array.flags["C_CONTIGUOUS"]
        self._sink = host_array  # keep it alive while copies may be in flight
        self._lib.tgb200_set_frame_sink_ex(C.byref(self._rt), host_array.ctypes.data, n_bands, 0 if host_array.dtype == np.float32 else 1)

    def read_present(self, out=None):
        """The presented frame: uint32 [h, w], a << 24 | r << 16 | g << 8 | b."""
        if out is None:
            out = np.empty((self.height, self.width), dtype=np.uint32)
        self._lib.tg_raytracer_read_present(C.byref(self._rt), T.ptr(out, T.u32))
        self._check()
        return out

    def scene_save(self, path):
        ok = self._lib.tgb200_scene_save(C.byref(self._rt), str(path).encode())
        self._check()
        return bool(ok)

    def scene_load(self, path):
        ok = self._lib.tgb200_scene_load(C.byref(self._rt), str(path).encode())
        self._check()
        return bool(ok)

    def save_frame_bmp(self, path):
        ok = self._lib.tgb200_save_frame_bmp(C.byref(self._rt), str(path).encode())
        self._check()
        return bool(ok)

    def frame_ticket(self):
        return int(self._lib.tgb200_frame_ticket(C.byref(self._rt)))

    def wait_frame(self, ticket):
        self._lib.tgb200_wait_frame(C.byref(self._rt), ticket)
        self._check()

    def set_gi_traversal(self, kind):
        """0 = automatic (stackless over the flattened tree), 1 = the stack machine of svo_functions.inc."""
        self._lib.tgb200_set_gi_traversal(C.byref(self._rt), kind)

    def mark_svo_dirty(self):
        self._lib.tgb200_mark_svo_dirty(C.byref(self._rt))

    def comm_destroy(self):
        self._lib.tgb200_comm_destroy(C.byref(self._rt))

    # ---- scene loading -----------------------------------------------------------------------------
    def load_scene(self, scene):
        """Creates every object of a tg_b200.scenes.SceneSpec and its LUT (tg_application.c:64-98 analogue)."""
        for o in scene.objects:
            if o.bits is None and o.seed is not None:
                self.create_object_synthetic(o.center, o.extent, o.angle, o.axis, o.seed, o.k, o.lut_idx)
            else:
                self.create_object_from_data(o.center, o.extent, o.angle, o.axis, o.bits, o.lut_indices, o.lut_idx)
        for i, (r, g, b) in enumerate(scene.lut):
            self.color_lut_set(i, r, g, b)
        self.synchronize()


def comm_unique_id():
    from . import check, lib
    buf = (T.u8 * 128)()
    lib().tgb200_comm_unique_id(buf)
    check()
    return bytes(buf)


def from_scene(scene, device=0, max_n_objects=None, max_n_clusters=None):
    cam = T.make_camera(scene.camera.position, scene.camera.pitch, scene.camera.yaw, scene.camera.roll,
                        scene.camera.fov_y_deg, scene.camera.aspect, scene.camera.near, scene.camera.far)
    rt = Raytracer(cam, max_n_objects or max(len(scene.objects), 1), max_n_clusters or max(scene.n_clusters, 1), scene.width, scene.height, device)
    rt.load_scene(scene)
    return rt

/*
 * tg_raytracer.h -- the drop-in boundary of libtgb200.so.
 *
 * Every `tg_*` entry point below keeps the name, argument order and meaning of the reference's
 * raytracer API (/root/reference/tg/src/graphics/vulkan/tgvk_raytracer.h:240-251 and
 * graphics/tg_sparse_voxel_octree.h:51-53); what sits behind them is hand-written sm_100a CUDA
 * instead of Vulkan passes. `tgb200_*` / `*_ex` / `*_from_data` names are documented EXTENSIONS the
 * reference lacks but BASELINE.json's configs need (last row).
 *
 * Error convention (tg_common.h:33,42, tgvk_common.h:25-37): the reference returns void / b32 and
 * asserts in debug. We keep the signatures; a failed CUDA/NCCL call or violated precondition is
 * recorded and readable through tgb200_last_error() (NULL == no error). There is NO CPU fallback:
 * without a CUDA device tg_raytracer_create() fails loudly (error string set, p_device == NULL,
 * every later call on that raytracer is a recorded error).
 *
 * Threading: single-threaded like the reference (memory/tg_memory.c:270,302).
 */
#ifndef TG_RAYTRACER_H
#define TG_RAYTRACER_H

#include "tg_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define TG_EXPORT __attribute__((visibility("default")))

struct tgb_device; /* opaque: device buffers, streams, events */

/*
 * Replaces the reference's `tg_raytracer` (tgvk_raytracer.h:110-236): the Vulkan members are
 * gone, `p_camera` and `scene` keep their names because the application reads them directly
 * (tg_application.c:228 reads scene.p_cluster_idx_to_object_idx).
 */
typedef struct tg_raytracer
{
    const tg_camera*   p_camera;  /* BORROWED for the raytracer's lifetime, re-read every render() (tgvk_raytracer.c:674,1153) */
    tg_scene           scene;     /* owned; CPU mirror exactly as in the reference */
    struct tgb_device* p_device;  /* owned; NULL if creation failed */
    u32                width;     /* reference: swapchain extent (tgvk_raytracer.c:669-670) */
    u32                height;
    u32                debug_visualization;
    u32                gi_enabled;    /* extension: 1 = one secondary ray per hit pixel through the SVO */
    u32                frame_seed;    /* extension: seed of the secondary-ray RNG */
    u32                svo_dirty;     /* 0 = current; 1 = only transforms changed (incremental update); 2 = objects created / destroyed (full rebuild) */
    u32*               p_object_lut_idx; /* extension: per-object LUT index (README.md:12,18) */
    u32                n_color_luts;
    u32                n_moved_objects;  /* objects whose transform changed since the last SVO build */
    u32*               p_moved_objects;  /* [object_capacity] */
} tg_raytracer;

/* ---- reference entry points (tgvk_raytracer.h:240-251) ------------------------------------ */

/* tgvk_raytracer.c:662-789. Resolution = tgb200_set_default_resolution() (stands in for the swapchain). */
TG_EXPORT void tg_raytracer_create(const tg_camera* p_camera, u32 max_n_objects, u32 max_n_clusters, tg_raytracer* p_raytracer);
/* tgvk_raytracer.c:791-796 (reference: TG_NOT_IMPLEMENTED; implemented here, Q10). */
TG_EXPORT void tg_raytracer_destroy(tg_raytracer* p_raytracer);
/* tgvk_raytracer.c:798-803 */
/* TG_DEBUG_SHOW_BLOCKS: like the reference (tgvk_raytracer.c:1226-1272), the visibility buffer is then written by a primary-ray pass
 * through the SVO (debug_visibility_svo.frag:27-71: depth24 | leaf node index | voxel % 512; the SVO is built when stale) instead of
 * the cluster pass, and the shading pass hashes cluster_pointers[node index] (shading.frag:122-126,247-256; an index beyond the live
 * pointer range reads as cluster 0). On a sharded raytracer the pointer table is distributed: the BLOCKS view hashes the node index
 * itself, CLUSTER_INDEX the global pointer and COLOR_LUT_INDEX 0 -- those three debug views differ from the single-GPU image. */
TG_EXPORT void tg_raytracer_set_debug_visualization(tg_raytracer* p_raytracer, tg_debug_show type);
/* tgvk_raytracer.c:805-992: object with the reference's procedural simplex-noise terrain fill. */
TG_EXPORT void tg_raytracer_create_object(tg_raytracer* p_raytracer, v3 center, v3u extent);
/* tgvk_raytracer.c:994-1067: returns cluster indices, compacts the pointer table downward. */
TG_EXPORT void tg_raytracer_destroy_object(tg_raytracer* p_raytracer, u32 object_idx);
/* tgvk_raytracer.c:1069-1077 */
TG_EXPORT b32  tg_object_is_initialized(const tg_scene* p_scene, u32 object_idx);
/* tgvk_raytracer.c:1122-1142: LUT 0. packed = r<<24|g<<16|b<<8|255, channel = (u32)(c*255.0f). */
TG_EXPORT void tg_raytracer_color_lut_set(tg_raytracer* p_raytracer, u8 index, f32 r, f32 g, f32 b);
/* tgvk_raytracer.c:1144-1553: camera -> (SVO) -> visibility -> GI + shading. */
TG_EXPORT void tg_raytracer_render(tg_raytracer* p_raytracer);
/* tgvk_raytracer.c:1556-1603 + clear.comp:15-21: visibility buffer <- all ones. */
TG_EXPORT void tg_raytracer_clear(tg_raytracer* p_raytracer);
/* tgvk_raytracer.c:1605-1655: reads one u64; p_cluster_idx receives the raw 31-bit POINTER field (Q1). */
TG_EXPORT b32  tg_raytracer_get_hovered_voxel(tg_raytracer* p_raytracer, u32 screen_x, u32 screen_y, f32* p_depth, u32* p_cluster_idx, u32* p_voxel_idx);

/* ---- reference SVO entry points (tg_sparse_voxel_octree.h:51-53) --------------------------- */

/* tg_sparse_voxel_octree.c:466-542. Built on the GPU from the scene's device mirror, copied out
 * into malloc'ed host arrays owned by *p_svo. The scene must belong to a live tg_raytracer. */
TG_EXPORT void tg_svo_create(v3 extent_min, v3 extent_max, const tg_scene* p_scene, tg_svo* p_svo);
/* tg_sparse_voxel_octree.c:544-556 */
TG_EXPORT void tg_svo_destroy(tg_svo* p_svo);
/* tg_sparse_voxel_octree.c:558-740: host-side single-ray query (picking/debug), as in the reference. */
TG_EXPORT b32  tg_svo_traverse(const tg_svo* p_svo, v3 ray_origin, v3 ray_direction, f32* p_distance, u32* p_node_idx, u32* p_voxel_idx);

/* ---- extensions ---------------------------------------------------------------------------- */

/* NULL when no error has been recorded since the last tgb200_clear_error(). */
TG_EXPORT const char* tgb200_last_error(void);
TG_EXPORT void        tgb200_clear_error(void);
/* Number of CUDA devices visible, or 0 (never fails). */
TG_EXPORT i32         tgb200_device_count(void);
/* CUDA device ordinal used by the next tg_raytracer_create (default 0; multi-GPU: LOCAL_RANK). */
TG_EXPORT void        tgb200_set_device(i32 device);
/* Resolution picked up by the next tg_raytracer_create (default 1920x1080). */
TG_EXPORT void        tgb200_set_default_resolution(u32 width, u32 height);
/* Reallocates the visibility / radiance buffers (reference: swapchain resize). */
TG_EXPORT void        tg_raytracer_set_resolution(tg_raytracer* p_raytracer, u32 width, u32 height);

/*
 * Object whose voxels are given instead of generated (the reference can only fill procedurally,
 * tgvk_raytracer.c:871-978). `p_solid_bits`: 16 u32 per cluster, clusters in pointer order
 * (x fastest, then y, then z), bit 64z+8y+x. `p_lut_indices`: 512 u8 per cluster, same order, or
 * NULL for the reference's rule (8*rel_x + vx) % 256 (tgvk_raytracer.c:947-978).
 * `axis` must be a unit vector (|axis|^2 within 2e-5 of 1): tgm_m4_angle_axis does not normalise it and the object-level culling
 * assumes a rigid transform; anything else is a recorded error (the same holds for create_object_synthetic, set_object_transform
 * and scene_load). Returns the object index, TG_U32_MAX on error -- a failed create leaves the scene and the device as they were.
 */
TG_EXPORT u32  tg_raytracer_create_object_from_data(tg_raytracer* p_raytracer, v3 center, v3u extent, f32 angle_in_radians, v3 axis, u32 lut_idx,
                                                    const u32* p_solid_bits, const u8* p_lut_indices);
/*
 * Object filled ON THE DEVICE with seeded random solid bits (BASELINE.json's "random solid bits" configs; a 10^11-voxel world
 * cannot be generated on the host and uploaded): cluster `rel` of the object owns the stream
 * state0 = hash_u32(object_seed ^ hash_u32(rel)) | 1 (math/tg_math.c:809-820), every mask word is the AND of `k` successive
 * xorshift32 draws (math/tg_math.c:328-338), i.e. density 2^-k; materials follow the reference's rule (8*rel_x + vx) % 256.
 * The CPU mirror scene.p_voxel_cluster_data is filled like for every other object. tgb200_synthetic_solid_bits is the host
 * twin of the generator (usable without a GPU). Returns the object index, TG_U32_MAX on error.
 */
TG_EXPORT u32  tg_raytracer_create_object_synthetic(tg_raytracer* p_raytracer, v3 center, v3u extent, f32 angle_in_radians, v3 axis, u32 lut_idx,
                                                    u32 object_seed, u32 k);
TG_EXPORT void tgb200_synthetic_solid_bits(u32 object_seed, u32 k, u32 n_clusters, u32* p_out /* 16 u32 per cluster */);
/* Moves an object (the reference has no setter); marks the SVO for an incremental update. */
TG_EXPORT void tg_raytracer_set_object_transform(tg_raytracer* p_raytracer, u32 object_idx, v3 translation, f32 angle_in_radians, v3 axis);
/* color_lut_set for LUT `lut_idx` (lut[lut_idx*256 + index]); lut_idx 0 == the reference call. */
TG_EXPORT void tg_raytracer_color_lut_set_ex(tg_raytracer* p_raytracer, u32 lut_idx, u8 index, f32 r, f32 g, f32 b);
/* GI on/off + RNG seed of the secondary rays. */
TG_EXPORT void tg_raytracer_set_gi(tg_raytracer* p_raytracer, b32 enabled, u32 frame_seed);

/*
 * Frame sink: the reference hands its HDR target to the swapchain (tgvk_raytracer.c:1524-1553); a host application of this
 * library reads the frame back instead. With a sink set, every tg_raytracer_render() / tgb200_render_shading() shades in
 * `n_bands` row bands (1..16) and copies each finished band into `p_host` on a second stream while the next band is
 * shaded, so the PCIe transfer of the 16 B/pixel frame overlaps the rendering. `p_host` receives the rows this rank
 * shades (tgb200_tile_rows; the whole frame on one GPU), first shaded row first, 4 floats per pixel; pinned memory makes
 * the copy asynchronous. The sink may be changed between frames (double buffering): a frame's ticket is
 * tgb200_frame_ticket() right after its render call, tgb200_wait_frame(ticket) blocks until that frame is complete in
 * host memory, tgb200_synchronize() waits for everything. NULL switches the sink off.
 */
TG_EXPORT void tgb200_set_frame_sink(tg_raytracer* p_raytracer, f32* p_host, u32 n_bands);
/*
 * Present pass (present.frag:9-12; tgvk_raytracer.c:560-630,1524-1553): the reference ends a frame by copying the HDR target 1:1
 * into the swapchain image, VK_FORMAT_B8G8R8A8_UNORM (tgvk_core.c:4239-4246). TGB200_SINK_BGRA8 makes the frame sink deliver
 * that image instead of the HDR rows: one u32 per pixel, a << 24 | r << 16 | g << 8 | b (bytes B, G, R, A), each channel
 * NaN -> 0, clamped to [0, 1], round-to-nearest-even of c * 255 -- a quarter of the HDR frame's bytes over PCIe.
 */
typedef enum tgb200_sink_format { TGB200_SINK_RGBA32F = 0, TGB200_SINK_BGRA8 = 1 } tgb200_sink_format;
TG_EXPORT void tgb200_set_frame_sink_ex(tg_raytracer* p_raytracer, void* p_host, u32 n_bands, tgb200_sink_format format);
/* The presented frame (whole frame, synchronous): w*h u32, B8G8R8A8_UNORM as above. */
TG_EXPORT void tg_raytracer_read_present(tg_raytracer* p_raytracer, u32* p_out);
/*
 * The presented frame as a .bmp file in the container the reference's tg_image_store_to_disc writes (tg_image_io.c:438-520):
 * BITMAPV5HEADER, BI_BITFIELDS, 32 bits per pixel, top-down rows, pixel data at offset 150. Returns TG_FALSE on an I/O error.
 */
TG_EXPORT b32  tgb200_save_frame_bmp(tg_raytracer* p_raytracer, const char* p_filename);
/* The writer alone (pure host): `p_pixels` = w*h B8G8R8A8 words, first row on top. */
TG_EXPORT b32  tgb200_write_bmp_bgra8(const char* p_filename, u32 width, u32 height, const u32* p_pixels);
TG_EXPORT u64  tgb200_frame_ticket(tg_raytracer* p_raytracer);
TG_EXPORT void tgb200_wait_frame(tg_raytracer* p_raytracer, u64 ticket);

/* Secondary-ray kernel: 0 = automatic (stackless over the flattened tree whenever the SVO box corners are multiples of 32;
 * the stack machine of svo_functions.inc otherwise), 1 = always the stack machine. Both give the same radiance (tests). */
TG_EXPORT void tgb200_set_gi_traversal(tg_raytracer* p_raytracer, u32 kind);

/* Stages of render(), individually callable (bench / tests). All asynchronous on the raytracer's stream. */
TG_EXPORT void tgb200_render_visibility(tg_raytracer* p_raytracer);          /* camera + cull + K1 */
TG_EXPORT void tgb200_svo_update(tg_raytracer* p_raytracer, b32 force_full); /* K2: rebuild or incremental */
/* Leaves the most recent tgb200_svo_update re-sampled (== all leaves after a full build). */
TG_EXPORT u32  tgb200_svo_leaves_resampled(tg_raytracer* p_raytracer);
TG_EXPORT void tgb200_render_shading(tg_raytracer* p_raytracer);             /* K3: (GI +) LUT shading */
/* K3 over physical rows [first_row, one_past_last_row) only (one GPU): what one rank of a sharded frame shades, without the
 * exchange -- used to measure and test the shading stage at a screen tile's size. Other rows of the radiance buffer keep their content. */
TG_EXPORT void tgb200_render_shading_rows(tg_raytracer* p_raytracer, u32 first_row, u32 one_past_last_row);
TG_EXPORT void tgb200_synchronize(tg_raytracer* p_raytracer);

/* Copies of results into caller memory (synchronous). */
TG_EXPORT void tg_raytracer_read_visibility(tg_raytracer* p_raytracer, u64* p_out /* w*h */);
TG_EXPORT void tg_raytracer_read_radiance(tg_raytracer* p_raytracer, f32* p_out /* w*h*4, RGBA32F */);
/* Frame rows [first_row, one_past_last_row) only; p_out receives those rows. */
TG_EXPORT void tg_raytracer_read_radiance_rows(tg_raytracer* p_raytracer, u32 first_row, u32 one_past_last_row, f32* p_out);
/* Replaces the device visibility buffer (tests: feed an oracle-made buffer to the shading stage). */
TG_EXPORT void tg_raytracer_write_visibility(tg_raytracer* p_raytracer, const u64* p_in /* w*h */);
/* Device SVO -> freshly malloc'ed host arrays in *p_svo (free with tg_svo_destroy). */
TG_EXPORT void tgb200_svo_download(tg_raytracer* p_raytracer, tg_svo* p_svo);
/* Host SVO -> device (tests: shade with an oracle-made SVO). */
TG_EXPORT void tgb200_svo_upload(tg_raytracer* p_raytracer, const tg_svo* p_svo);

/* Per-stage device time of the most recent call of each stage, CUDA events, milliseconds. */
typedef struct tgb200_timings
{
    f32 clear_ms;
    f32 cull_ms;
    f32 visibility_ms;
    f32 svo_ms;
    f32 shading_ms;
    f32 merge_ms;
    u32 n_visible_objects;
    u32 n_kernel_launches; /* kernels of this library launched since create/reset */
    u32 n_gi_rays;         /* secondary rays the last frame traced through the SVO (those that enter its box) */
    u32 n_gi_rays_exact;   /* of those, the rays the certified fast walk (tgb_gi_fast.cu) handed to the exact kernel; all of them when the exact kernel runs alone */
    u64 n_gi_node_visits;  /* work of those rays: node visits, leaf DDA steps, advances (svo_functions.inc loop iterations) */
    u64 n_gi_dda_steps;
    u64 n_gi_advances;
    /* parts of merge_ms on the peer-memory path (0 otherwise): local material resolve | all-gather of the object records, which is
     * where a rank waits for the slowest rank's K1 | k_merge_tile (min + winner's material over NVLink) */
    f32 merge_resolve_ms;
    f32 merge_gather_ms;
    f32 merge_kernel_ms;
    u32 pad2;
} tgb200_timings;
TG_EXPORT void tgb200_get_timings(tg_raytracer* p_raytracer, tgb200_timings* p_out);
TG_EXPORT void tgb200_reset_launch_counter(tg_raytracer* p_raytracer);

/* Raw device pointers + stream (for zero-copy interop, e.g. wrapping in a torch tensor). Multi-GPU: the buffers are padded to whole
 * 16-row bands and keep rows in tile order, rank-major (tgb200_tile_rows); on one GPU that is frame order. */
TG_EXPORT void* tgb200_device_visibility(tg_raytracer* p_raytracer);
TG_EXPORT void* tgb200_device_radiance(tg_raytracer* p_raytracer);
TG_EXPORT void* tgb200_stream(tg_raytracer* p_raytracer);

/*
 * Multi-GPU: one process per GPU, clusters sharded by object. A shard's packed
 * words carry GLOBAL cluster pointers = local pointer + global_pointer_base.
 */
TG_EXPORT void tgb200_set_shard(tg_raytracer* p_raytracer, u32 rank, u32 n_ranks, u32 global_pointer_base);
/* 128-byte NCCL unique id (rank 0 creates, host code broadcasts, every rank joins). */
TG_EXPORT void tgb200_comm_unique_id(u8* p_out_128);
TG_EXPORT void tgb200_comm_init(tg_raytracer* p_raytracer, const u8* p_unique_id_128, u32 rank, u32 n_ranks);
TG_EXPORT void tgb200_comm_destroy(tg_raytracer* p_raytracer);
/* ncclAllReduce(ncclUint64, ncclMin) over the visibility buffer, in place: afterwards every rank holds the whole merged frame. */
TG_EXPORT void tgb200_merge_visibility(tg_raytracer* p_raytracer);
/*
 * How tg_raytracer_render / tgb200_render_shading exchange a sharded frame. 0 (default) = one kernel over peer memory: every rank
 * maps the other ranks' visibility / material buffers (CUDA IPC over NVLink, collective set-up on first use) and resolves its own
 * screen tile -- min over the ranks' words + the winner's material -- without the two whole-frame collectives; falls back to 1 on
 * every rank when the buffers cannot be mapped. 1 = NCCL only: all-reduce(min) + owner-resolved materials + reduce-scatter(max).
 * Both give the same frame, bit for bit. On the peer-memory path the device visibility buffer keeps this rank's LOCAL words;
 * tg_raytracer_read_visibility / get_hovered_voxel pull the merged frame from the peers (call them between frames, ranks in
 * lock-step), tgb200_merge_visibility still produces it in place.
 */
TG_EXPORT void tgb200_set_merge_kind(tg_raytracer* p_raytracer, u32 kind);
/*
 * The rows this rank shades. GI rays are split by screen tile; to give every rank an equal share of the hit pixels the frame is
 * cut into bands of 16 rows and band b belongs to rank b mod n_ranks (tg_b200/csrc/tgb_rows.h). A rank's rows are numbered
 * [first, one_past_last) in "tile order" -- its bands one after the other; tile order is what the frame sink delivers and what
 * tgb200_gather_radiance exchanges. tgb200_tile_physical_row maps the i-th row of this rank's tile (0 <= i < one_past_last - first)
 * to the frame row it shows, or TG_U32_MAX for a padding row (the last band of the frame may be partial, the last ranks may own
 * one band less). On one GPU the tile is the whole frame, [0, height), in frame order. The read_* / write_* / get_hovered_voxel
 * entry points always speak frame rows.
 */
TG_EXPORT void tgb200_tile_rows(tg_raytracer* p_raytracer, u32* p_first_row, u32* p_one_past_last_row);
TG_EXPORT u32  tgb200_tile_physical_row(tg_raytracer* p_raytracer, u32 tile_row);
/* ncclAllGather of the radiance tiles: afterwards every rank holds the full frame (optional; collective). */
TG_EXPORT void tgb200_gather_radiance(tg_raytracer* p_raytracer);
/* Marks the replicated SVO stale on a rank that does not own the object another rank moved (scene edits are mirrored
 * on every rank in lock-step; the next render() / tgb200_svo_update() rebuilds collectively). */
TG_EXPORT void tgb200_mark_svo_dirty(tg_raytracer* p_raytracer);

/*
 * Scene dump / load: every initialised object (dims, transform, LUT index, solid masks, material indices) and the colour LUTs in one
 * little-endian file (layout in tgb_host_extra.c). tgb200_scene_load creates the objects in a live raytracer through
 * tg_raytracer_create_object_from_data, in the file's order: a fresh raytracer gets object indices 0..n-1 (holes left by destroyed
 * objects are not reproduced). TG_FALSE + tgb200_last_error() on I/O errors, truncated files or exhausted capacity.
 */
TG_EXPORT b32  tgb200_scene_save(tg_raytracer* p_raytracer, const char* p_filename);
TG_EXPORT b32  tgb200_scene_load(tg_raytracer* p_raytracer, const char* p_filename);

/* ---- pure host logic, usable without a GPU (scene bookkeeping, camera) ---------------------- */

/* tgvk_raytracer.c:687-712: allocates the CPU arrays, fills the LIFO free-lists descending. */
TG_EXPORT void tgb200_scene_init(tg_scene* p_scene, u32 max_n_objects, u32 max_n_clusters);
TG_EXPORT void tgb200_scene_free(tg_scene* p_scene);
/* tgvk_raytracer.c:816-866: pops an object index and its clusters; returns the object index. */
TG_EXPORT u32  tgb200_scene_alloc_object(tg_scene* p_scene, v3 center, v3u extent, f32 angle_in_radians, v3 axis);
/* tgvk_raytracer.c:1000-1066; p_first_shifted_pointer/p_n_shifted describe the compacted range. */
TG_EXPORT void tgb200_scene_free_object(tg_scene* p_scene, u32 object_idx, u32* p_first_shifted_pointer, u32* p_n_shifted);
/* tgvk_core.c:382-444 + tgvk_raytracer.c:1171-1180 */
TG_EXPORT void tgb200_camera_rays(const tg_camera* p_camera, tg_camera_rays* p_out);
/* tgvk_raytracer.c:836-847: the 96-byte record of object `object_idx`. */
TG_EXPORT void tgb200_object_data(const tg_scene* p_scene, u32 object_idx, u32 lut_idx, tg_object_data* p_out);
/* tgvk_raytracer.c:1130-1134 */
TG_EXPORT u32  tgb200_pack_color(f32 r, f32 g, f32 b);
/* The ray of pixel (px,py) in the space of cluster `cluster_pointer`, evaluated on the HOST with the
 * per-object factorisation the kernels use (tg_b200/csrc/tgb_hoist.h); visibility.frag:35-69. Verification hook. */
TG_EXPORT void tgb200_debug_cluster_ray(const tg_object_data* p_object, const tg_camera_rays* p_cam, u32 width, u32 height, u32 px, u32 py,
                                        u32 cluster_pointer, v3* p_origin_ms, v3* p_direction_ms);
/* tgvk_raytracer.c:871-943: the reference's procedural terrain bits for one object (16 u32 per cluster). */
TG_EXPORT void tgb200_procedural_solid_bits(u32 object_idx, v3u n_cluster_pointers_per_dim, u32* p_out);

#ifdef __cplusplus
}
#endif

#endif

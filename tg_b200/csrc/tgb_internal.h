/*
 * tgb_internal.h -- the thin C-ABI seam between the C host code (tgb_host.c) and the CUDA
 * translation units (tgb_device.cu, tgb_visibility.cu, tgb_shade.cu, tgb_svo.cu). Plain pointers
 * and sizes only; the host side never sees a CUDA type.
 */
#ifndef TGB_INTERNAL_H
#define TGB_INTERNAL_H

#include "../../include/tg_raytracer.h"

#ifdef __cplusplus
extern "C" {
#endif

/* device buffer ids for tgbd_upload / tgbd_download */
enum
{
    TGB_BUF_CLUSTER_POINTERS = 0, /* D3: u32[cluster_capacity] */
    TGB_BUF_C2O,                  /* D4: u32[cluster_capacity] */
    TGB_BUF_OBJECTS,              /* D5: tg_object_data[object_capacity] */
    TGB_BUF_MASKS,                /* D1: 64 B per cluster idx */
    TGB_BUF_LUT_IDX,              /* D2: 512 B per cluster idx */
    TGB_BUF_COLOR_LUT,            /* D6: u32[n_luts * 256] */
    TGB_BUF_VISIBILITY,           /* D7: u64[w*h] */
    TGB_BUF_RADIANCE,             /* RGBA32F[w*h] */
    TGB_BUF_VISIBILITY_MERGED,    /* read-only view: the whole merged frame (== TGB_BUF_VISIBILITY on one GPU or after the all-reduce) */
    TGB_BUF_SVO_NODES,
    TGB_BUF_SVO_LEAF_DATA,
    TGB_BUF_SVO_VOXELS,
    TGB_BUF_COUNT
};

/* error recording (tgb_host.c) */
void tgb_set_error(const char* p_fmt, ...);

/* ---- tgb_device.cu ---- */
i32   tgbd_env_int(const char* p_name, i32 fallback);
i32   tgbd_device_count(void);
struct tgb_device* tgbd_create(i32 device, u32 object_capacity, u32 cluster_capacity, u32 n_color_luts, u32 width, u32 height);
void  tgbd_destroy(struct tgb_device* d);
b32   tgbd_resize(struct tgb_device* d, u32 width, u32 height);
b32   tgbd_upload(struct tgb_device* d, u32 buffer, u64 dst_offset_bytes, const void* p_src, u64 n_bytes);
b32   tgbd_download(struct tgb_device* d, u32 buffer, u64 src_offset_bytes, void* p_dst, u64 n_bytes);
/* device-to-device move inside one buffer (pointer-table compaction, tgvk_raytracer.c:1049-1056) */
b32   tgbd_move(struct tgb_device* d, u32 buffer, u64 dst_offset_bytes, u64 src_offset_bytes, u64 n_bytes);
void  tgbd_synchronize(struct tgb_device* d);
b32   tgbd_flush_objects(struct tgb_device* d); /* pending object-record uploads -> the device (one copy); every device entry point that reads them calls it */
void* tgbd_buffer(struct tgb_device* d, u32 buffer);
void* tgbd_stream(struct tgb_device* d);
void  tgbd_get_timings(struct tgb_device* d, tgb200_timings* p_out);
void  tgbd_reset_launch_counter(struct tgb_device* d);
void  tgbd_set_shard(struct tgb_device* d, u32 global_pointer_base);
void  tgbd_set_gi_traversal(struct tgb_device* d, u32 kind);
/* frame sink: shading in row bands, each band copied to p_host (rows this rank shades, first shaded row first) on a second stream */
b32   tgbd_set_frame_sink(struct tgb_device* d, void* p_host, u32 n_bands, u32 format);
/* present pass over the whole frame into caller memory (synchronous): B8G8R8A8_UNORM, one u32 per pixel */
b32   tgbd_read_present(struct tgb_device* d, u32* p_out);
u64   tgbd_frames_sunk(struct tgb_device* d);
b32   tgbd_wait_frame(struct tgb_device* d, u64 ticket);
b32   tgbd_set_comm(struct tgb_device* d, void* p_comm, u32 rank, u32 n_ranks);
u32   tgbd_tile_rows(struct tgb_device* d);
u64   tgbd_padded_pixels(struct tgb_device* d); /* pixels of every frame buffer: width * tile_rows * n_ranks, rows in virtual order (tgb_rows.h) */
void* tgbd_comm(struct tgb_device* d);
u32   tgbd_rank(struct tgb_device* d);
u32   tgbd_n_ranks(struct tgb_device* d);

/* reference material rule (8*rel_x + vx) % 256 for clusters [first_pointer, first_pointer+n) of an object (tgvk_raytracer.c:947-978) */
b32   tgbd_fill_default_lut_idx(struct tgb_device* d, u32 first_pointer, u32 n_cluster_pointers, u32 nx);

/* ---- tgb_procedural.cu: the reference's simplex-noise terrain fill (tgvk_raytracer.c:868-943) on the device ---- */
b32   tgbd_procedural_fill(struct tgb_device* d, u32 object_idx, u32 nx, u32 ny, u32 nz, u32 first_pointer);
/* seeded random bits (density 2^-k) of an object's clusters [first_pointer, first_pointer + n) into the resident mask array */
b32   tgbd_synthetic_fill(struct tgb_device* d, u32 object_seed, u32 k, u32 n_clusters, u32 first_pointer);
b32   tgbd_procedural_bits_to_host(i32 device, u32 object_idx, u32 nx, u32 ny, u32 nz, u32* p_out);
i32   tgbd_current_device(void);

/* ---- tgb_visibility.cu ---- */
b32   tgbd_clear(struct tgb_device* d);                                                              /* clear.comp */
b32   tgbd_render_visibility(struct tgb_device* d, const tg_camera_rays* p_cam, u32 object_capacity); /* cull + K1 */

/* ---- tgb_visibility_pool.cu: K1 with several pixels per lane (after cull + sort; declines frames with an object > 32767 clusters along an axis) ---- */
b32   tgbd_k1_pool_render(struct tgb_device* d, const tg_camera_rays* p_cam);

/* ---- tgb_debug_svo.cu: the BLOCKS view's primary rays through the SVO (debug_visibility_svo.frag), instead of cull + K1 ---- */
b32   tgbd_render_visibility_svo(struct tgb_device* d, const tg_camera_rays* p_cam);

/* ---- tgb_shade.cu ---- */
/* rows [y0, y1) only (multi-GPU: this rank's screen tile); pointers outside [base, base + n_local_pointers) shade to 0 */
b32   tgbd_render_shading(struct tgb_device* d, const tg_camera_rays* p_cam, u32 n_local_pointers, u32 gi_enabled, u32 frame_seed, u32 debug_visualization, u32 y0, u32 y1);

/* multi-GPU frame tail: owner-resolved materials -> reduce-scatter by screen tile -> GI + shading of this rank's tile */
b32   tgbd_render_shading_sharded(struct tgb_device* d, const tg_camera_rays* p_cam, u32 n_local_pointers, u32 gi_enabled, u32 frame_seed, u32 debug_visualization);
/* all-gather of the radiance tiles so that every rank holds the full frame */
b32   tgbd_gather_radiance(struct tgb_device* d);

/* ---- tgb_gi_pool.cu: the queued secondary rays of one band through the flattened SVO, several rays per lane ---- */
b32   tgbd_gi_pool_trace(struct tgb_device* d, f32 far_plane);
b32   tgbd_gi_pool_trace_list(struct tgb_device* d, f32 far_plane, const u32* p_list, u32 count_word); /* p_list: queue slots handed over by the fast walk, counted in d_gi_count[count_word] */
b32   tgbd_gi_fast_trace(struct tgb_device* d, f32 far_plane, b32 tiled); /* tgb_gi_fast.cu: certified fast walk (over octree cells, or over the coarser tiling), then the exact kernel on what it handed over */

/* ---- tgb_svo.cu ---- */
b32   tgbd_svo_build(struct tgb_device* d, v3 extent_min, v3 extent_max, u32 n_cluster_pointers, u32 object_capacity);
/* incremental: only the leaves a moved object reaches now or reached in the previous build are re-sampled, the tree is
 * laid out afresh and clean leaves are copied; the arrays equal a full rebuild. Falls back to a full build when needed. */
b32   tgbd_svo_update_objects(struct tgb_device* d, v3 extent_min, v3 extent_max, u32 n_cluster_pointers, u32 object_capacity, u32 n_moved, const u32* p_object_indices);
void  tgbd_svo_invalidate_incremental(struct tgb_device* d);
u32   tgbd_svo_leaves_resampled(struct tgb_device* d);
b32   tgbd_svo_counts(struct tgb_device* d, u32* p_n_nodes, u32* p_n_leaves, u32* p_n_voxel_words, v3* p_min, v3* p_max);
b32   tgbd_svo_set(struct tgb_device* d, v3 bmin, v3 bmax, u32 n_nodes, const void* p_nodes, u32 n_leaves, const void* p_leaf_data, u32 n_voxel_words, const void* p_voxels);

/* ---- tgb_nccl.c ---- */
b32   tgbn_unique_id(u8* p_out_128);
void* tgbn_init(const u8* p_unique_id_128, u32 rank, u32 n_ranks);
void  tgbn_destroy(void* p_comm);
b32   tgbn_allreduce_min_u64(void* p_comm, void* p_device_buffer, u64 count, void* p_stream);
b32   tgbn_allreduce_sum_u32(void* p_comm, void* p_device_buffer, u64 count, void* p_stream);
b32   tgbn_allgather_bytes(void* p_comm, const void* p_send, void* p_recv, u64 n_bytes, void* p_stream);
b32   tgbn_reducescatter_max_u64(void* p_comm, const void* p_send, void* p_recv, u64 count_per_rank, void* p_stream);
/* ---- tgb_peer.cu: merge over peer memory ---- */
void  tgbd_set_merge_kind(struct tgb_device* d, u32 kind);
b32   tgbd_p2p_prepare(struct tgb_device* d);   /* collective on first use; TG_TRUE when the fused path can run */
void  tgbd_p2p_teardown(struct tgb_device* d);  /* collective: unmaps the peers' buffers (before they are freed) */
void  tgbd_note_merged(struct tgb_device* d);   /* the all-reduce has run on the current buffer */
void* tgbd_visibility_for_read(struct tgb_device* d); /* whole merged frame: pulls the other tiles from the peers if necessary */
b32   tgbd_p2p_merge_tile(struct tgb_device* d);
b32   tgbd_gather_objects(struct tgb_device* d);     /* collective: every rank's object records, pointers globalised */
b32   tgbd_p2p_barrier(struct tgb_device* d);        /* publish "K1 done" for this frame, wait for every peer's (device side, asynchronous) */
b32   tgbd_p2p_flag_all_tiles(struct tgb_device* d); /* this rank's words did not come from K1's epilogue: every tile counts as hit */
void  tgbd_p2p_check(struct tgb_device* d);          /* after a stream synchronisation: did the device-side barrier time out? */
/* events around the merge live in tgb_device.cu */
void  tgbd_merge_begin(struct tgb_device* d);
void  tgbd_merge_end(struct tgb_device* d);

#ifdef __cplusplus
}
#endif

#endif

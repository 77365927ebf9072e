/*
 * tgb_device.cuh -- device-side state of one raytracer (replaces the Vulkan buffers of
 * tg_raytracer_data, /root/reference/tg/src/graphics/vulkan/tgvk_raytracer.h:110-137).
 * All arrays live in HBM for the raytracer's lifetime; nothing is re-uploaded per frame except the
 * 96-byte camera block (kernel parameter).
 */
#ifndef TGB_DEVICE_CUH
#define TGB_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdio.h>

#include "tgb_internal.h"
#include "tgb_math.h"
#include "tgb_hoist.h"
#include "tgb_rows.h"
#include "tgb_gi_walk.cuh" /* flattened-tree layout (TGB_TOP_*) + the per-ray traversal pieces */

#define TGB_MAX_BANDS  16
#define TGB_MAX_RANKS  16
#define TGB_FRAME_RING 4

struct tgb_svo_device
{
    v3   bmin, bmax;
    u32  node_capacity, leaf_capacity, voxel_word_capacity;
    u32* d_nodes;
    u32* d_leaf_data;   /* 65 u32 per leaf */
    u32* d_voxels;      /* 1024 u32 per leaf */
    u32* d_counts;      /* [0] nodes, [1] leaves, [2] overflow flag */
    u32* d_top_grid;    /* [32^3 + 1] the tree flattened per 32^3 cell (k_svo_flatten): terminal level | has data | leaf data pointer;
                           last word: non-zero = the grid describes the tree completely (leaves exactly at depth 5) */
    u32* d_fast_cells;  /* [3 * 32^3] the certified fast walk's coarser tiling of the free table cells (tgb_gi_fast.cuh) + two scratch passes; on first use */
    u32* d_fast_bricks; /* [leaf_capacity * 64] the same per 8^3 brick of every leaf block */
    u32* d_fast_columns; /* [2 * fast_columns_capacity_leaves * 1024] the blocks' voxels with y, then with z, as the bit index; grown on demand */
    u32  fast_columns_capacity_leaves;
    b32  fast_tiling_valid; /* both describe the current tree */
    u32  n_nodes, n_leaves, n_pairs;
    b32  valid;
    /* build scratch */
    u32* d_pairs_a;     /* cluster pointers grouped by leaf (segments in dense-leaf order), grown on demand */
    u32* d_pair_leaf_a; /* dense leaf of every pair (incremental update: which leaves did a moved object reach before) */
    u32* d_pairs_b;     /* the other half of the ping-pong: an incremental update reads the previous build's pairs */
    u32* d_pair_leaf_b;
    u8*  d_pair_flags;  /* per pair: the cluster set at least one bit of the leaf */
    u64  pair_capacity, pair_capacity_b, pair_flags_capacity;
    u32* d_moved_indices; /* [object_capacity] staging of the moved-object list */
    u32* d_voxels_alt;  /* spare leaf arrays: an incremental update writes here, copies clean leaves over, then swaps */
    u32* d_leaf_data_alt;
    u32* d_object_moved; /* [object_capacity] object moved since the last build */
    b32  incremental_ok; /* the current arrays and pair lists come from a K2 build of the current object set */
    u32  n_leaves_resampled; /* leaves the last update re-sampled (the rest were copied) */
    u32* d_scratch;     /* dense-tree arrays, see tgb_svo.cu */
    u64  scratch_capacity;
    u32* d_object_flags; /* [object_capacity] object can touch the SVO box */
    /* sharded build: this rank's contribution per leaf (voxels | leaf records) and the all-gathered copies of every rank */
    u32* d_part;
    u32* d_gather;
    u64  part_capacity_leaves, gather_capacity_leaves;
};

struct tgb_device
{
    i32          device;
    cudaStream_t stream;
    u32          object_capacity, cluster_capacity, n_color_luts, width, height;
    u32          global_pointer_base;

    u32*            d_cluster_pointers;
    u32*            d_c2o;
    tg_object_data* d_objects;
    u32*            d_masks;
    u8*             d_lut_idx;
    u32*            d_color_lut;
    u64*            d_vis;
    float4*         d_radiance;
    float4*         d_gi_q0;      /* secondary-ray queue, [w*h] each: origin.xyz + pixel | direction | ambient */
    float4*         d_gi_q1;
    float4*         d_gi_q2;
    u32*            d_gi_exact;   /* [w*h] queue slots the certified fast walk handed to the exact kernel (tgb_gi_fast.cu) */
    u32*            d_gi_count;   /* u32 [0] queued, [1] fetched; u64 [1] node visits, [2] DDA steps, [3] advances; u32 [10] rays of the frame, [12] handed over (band), [13] fetched of those, [14] handed over (frame) */
    u32*            h_gi_stats;   /* pinned copy of d_gi_count after the last frame */
    u32             n_sms;
    u32             gi_traversal; /* 0 = stackless when possible, 1 = stack machine */
    b32             gi_stats_valid; /* d_gi_count holds the counters of the last shaded frame */

    /* multi-GPU (one process per GPU): clusters sharded by object, SVO / objects replicated, GI split by screen tile */
    void*             p_comm;           /* NCCL communicator or NULL */
    u32               rank, n_ranks;
    u32               tile_rows;        /* tgb_rows.h: rank r shades VIRTUAL rows [r * tile_rows, (r + 1) * tile_rows) = the 16-row bands b with b mod n_ranks == r */
    u64*              d_mat;            /* [w * tile_rows * n_ranks] owner-resolved material words: global object idx << 32 | packed colour */
    u64*              d_mat_tile;       /* [w * tile_rows] this rank's tile after the reduce-scatter */
    /* merge over peer memory (tgb_peer.cu): every rank maps the other ranks' visibility / material buffers (CUDA IPC over NVLink) and
     * resolves ITS screen tile with one kernel: min over the ranks' words + the winner's material word. Both buffers are double
     * buffered so that one collective per frame (the all-gather of the object records) is the only synchronisation. */
    u32               merge_kind;       /* 0 = automatic: peer memory when it can be mapped, else NCCL; 1 = NCCL collectives */
    b32               p2p_ready, p2p_failed;
    b32               vis_merged;       /* d_vis holds the all-reduced words of the whole frame (tgb200_merge_visibility) */
    b32               tile_merged;      /* d_vis_tile holds the merged words of this rank's tile (fused path); d_vis the local ones */
    u64*              d_vis_tile;       /* [w * tile_rows] */
    u32               vis_flip;
    u64*              d_vis_pair[2];    /* [0] = the buffer tgbd_resize allocated, [1] on first use */
    u64*              d_mat_pair[2];
    u64*              peer_vis[2][TGB_MAX_RANKS]; /* [flip][rank]; own rank = local pointer */
    u64*              peer_mat[2][TGB_MAX_RANKS];
    u64*              d_vis_full;       /* whole merged frame pulled from the peers on demand (read-back, picking) */
    /* Behind the material words every d_mat buffer carries a small tail the peers read too (same IPC mapping): one u32 per 16x16 tile
     * "this rank has a hit in the tile" (written by K1; the material pass and the peers' k_merge_tile skip empty tiles: at N = 8 a tile is covered by one or
     * two ranks' objects, not eight) and the frame counter this rank publishes when its K1 is done (tgb_peer.cu: the device-side barrier). */
    u32               frame_seq;        /* frames started on the peer-memory path (tgbd_clear); equal on all ranks (they render in lock-step) */
    b32               tiles_flagged;    /* K1 wrote this frame's tile flags (false: uploaded words, BLOCKS view -> every tile counts as hit) */
    b32               objects_gathered; /* the object records were all-gathered since the last clear */
    u8*               d_ipc_stage;
    tg_object_data*   d_objects_global; /* [n_ranks * object_capacity] every rank's records, pointers globalised */
    tgb_object_frame* d_frames_global;

    tgb_object_frame* d_frames;        /* [object_capacity] compacted visible objects */
    tgb_object_frame* d_frames_sorted; /* [object_capacity] front-to-back */
    tgb_object_frame* d_frames_all;    /* [object_capacity] indexed by object idx (K3) */
    u32*              d_visible_count; /* [1] */
    u32*              h_visible_count; /* pinned */

    tgb_svo_device svo;

    /* frame sink (tgb200_set_frame_sink): the shading stage runs in row bands and every finished band is copied to host
     * memory on a second stream while the next band is shaded */
    /* object records are uploaded lazily: tgbd_upload(TGB_BUF_OBJECTS) lands in this host shadow, and the dirty byte range goes to the device in ONE
     * copy the next time anything on the device can read it (tgbd_flush_objects) -- 64 moved objects are one 6 KB copy instead of 64 96-byte ones */
    u8*          h_objects;         /* [object_capacity * sizeof(tg_object_data)] */
    u64          objects_dirty_lo, objects_dirty_hi; /* byte range, empty when lo >= hi */
    cudaStream_t svo_stream;        /* K2 runs here on one GPU, concurrently with K1 (tgb_svo.cu) */
    cudaEvent_t  ev_inputs;         /* recorded on the main stream by tgbd_clear: the frame's uploads are queued, the previous frame's shading too */
    b32          ev_inputs_valid, inputs_changed_since_clear;
    cudaStream_t copy_stream;
    f32*         p_sink;            /* caller memory (pinned for a truly asynchronous copy) or NULL */
    u32          sink_bands;        /* bands per frame, 1..TGB_MAX_BANDS */
    u32          sink_format;       /* TGB200_SINK_RGBA32F: the HDR rows; TGB200_SINK_BGRA8: the presented rows (k_present), 4 B / pixel */
    u32*         d_present_pair[2]; /* B8G8R8A8_UNORM frames, allocated on first use; [flip] like the radiance pair */
    cudaEvent_t  ev_band[TGB_MAX_BANDS];      /* band k shaded (main stream) */
    cudaEvent_t  ev_band_copied[2][TGB_MAX_BANDS]; /* band k of buffer p copied (copy stream): the next frame shaded into p waits for it */
    /* the radiance buffer is double-buffered while a sink is set (frame i+1 is shaded while frame i is still being copied) */
    float4*      d_radiance_pair[2]; /* [0] allocated by tgbd_resize, [1] on first use of a sink; d_radiance points at the current one */
    u32          radiance_flip;
    b32          band_copy_pending[2][TGB_MAX_BANDS];
    u32          band_row0[2][TGB_MAX_BANDS], band_row1[2][TGB_MAX_BANDS]; /* rows of the pending copies, per radiance buffer */
    cudaEvent_t  ev_frame_copied[TGB_FRAME_RING];
    u64          n_frames_sunk;     /* frames whose copies have been issued */

    cudaEvent_t ev[16];
    f32         clear_ms, cull_ms, visibility_ms, svo_ms, shading_ms, merge_ms, merge_resolve_ms, merge_gather_ms, merge_kernel_ms;
    b32         ev_merge_parts;
    b32         ev_clear, ev_vis, ev_svo, ev_shade, ev_merge;
    u32         n_visible_objects;
    u32         n_kernel_launches;
};

/*
 * Small per-frame bookkeeping goes through kernels, not cudaMemsetAsync / cudaMemcpyAsync: a memset or copy on the render
 * stream queues behind the frame sink's large device-to-host copies in the copy engine and stalls the stream until they
 * drain (measured: the shading bands of one frame and the visibility pass of the next waited for the whole 133 MB).
 * Counters stay on the device and are fetched by tgbd_get_timings only (a per-frame store to mapped host memory also
 * waits for the PCIe queue: +0.2 ms on the visibility pass).
 */
static __global__ void k_set_words(u32* __restrict__ p, u32 n, u32 value)
{
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) p[i] = value;
}

/* layout of a material buffer: [padded_px] u64 words | [n_tiles] u32 tile flags | [16] u32: frame counter, number of published objects, padding |
 * [object_capacity] tg_object_data published records (globalised pointers) | [object_capacity] u32 their object indices.
 * A rank publishes the objects that survived its cull (only those can have won a pixel), or all of them when no cull ran. */
static inline u32 tgbd_tiles_x(const struct tgb_device* d) { return (d->width + 15u) / 16u; }
static inline u32 tgbd_n_tiles(const struct tgb_device* d) { return tgbd_tiles_x(d) * (d->tile_rows * (d->n_ranks ? d->n_ranks : 1u) / TGB_BAND_ROWS); }
static inline u64 tgbd_padded_px(const struct tgb_device* d) { return (u64)d->width * d->tile_rows * (d->n_ranks ? d->n_ranks : 1u); }
static inline u64 tgbd_mat_flag_bytes(const struct tgb_device* d) { return (((u64)tgbd_n_tiles(d) + 16u) * sizeof(u32) + 15u) / 16u * 16u; }
static inline u64 tgbd_mat_bytes(const struct tgb_device* d) { return tgbd_padded_px(d) * sizeof(u64) + tgbd_mat_flag_bytes(d) + (u64)d->object_capacity * (sizeof(tg_object_data) + sizeof(u32)); }
static inline u32* tgbd_mat_tile_flags(const struct tgb_device* d, const u64* p_mat) { return (u32*)(p_mat + tgbd_padded_px(d)); }
static inline u32* tgbd_mat_signal(const struct tgb_device* d, const u64* p_mat) { return tgbd_mat_tile_flags(d, p_mat) + tgbd_n_tiles(d); } /* [0] frame counter, [1] published objects */
static inline tg_object_data* tgbd_mat_objects(const struct tgb_device* d, const u64* p_mat) { return (tg_object_data*)((u8*)tgbd_mat_tile_flags(d, p_mat) + tgbd_mat_flag_bytes(d)); }
static inline u32* tgbd_mat_object_indices(const struct tgb_device* d, const u64* p_mat) { return (u32*)(tgbd_mat_objects(d, p_mat) + d->object_capacity); }

#define TGB_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            tgb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
            return TG_FALSE;                                                                        \
        }                                                                                           \
    } while (0)

#define TGB_LAUNCH_CHECK(d)                                                                         \
    do {                                                                                            \
        (d)->n_kernel_launches++;                                                                   \
        cudaError_t e__ = cudaGetLastError();                                                       \
        if (e__ != cudaSuccess) {                                                                   \
            tgb_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return TG_FALSE;                                                                        \
        }                                                                                           \
    } while (0)

/* TGB_GI_KERNEL when the environment does not say: 4 = the certified fast walk over the coarser tiling of the free space, its first cell entered by
 * k_shade, the shader's own arithmetic (k_gi_trace_list) on the rays it hands over (tgb_gi_fast.cu: 0.93 ms for the stage against 1.38, identical frames,
 * profiles/r04*); 2 = the exact kernel on every ray (tgb_gi_pool.cu: the frame every other kernel must reproduce bit for bit); 1 / 3: see tgb_shade.cu */
#define TGB_GI_KERNEL_DEFAULT 4
extern "C" b32 tgbd_gi_fast_tiling_build(struct tgb_device* d, cudaStream_t st); /* tgb_gi_fast.cu */
extern "C" f32 tgbd_gi_fast_delta(void);
struct tgb_fast_tiling;
extern "C" void tgbd_gi_fast_tiling_get(struct tgb_device* d, struct tgb_fast_tiling* p_tiling);

#endif

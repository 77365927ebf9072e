/*
 * tgb_device.cu -- device memory, streams and events behind the thin C-ABI seam of tgb_internal.h.
 * Replaces the reference's Vulkan buffer / staging / queue layer for this path
 * (/root/reference/tg/src/graphics/vulkan/tgvk_raytracer.c:156-274 buffer set-up, tgvk_core.c:3345-3482
 * staging ring) with plain cudaMalloc'ed arrays on one stream.
 */
#include <stdlib.h>
#include <string.h>

#include "tgb_device.cuh"

/* tuning knobs read from the environment at every use (integers; unset or malformed = the default): a sweep can change them between frames */
extern "C" i32 tgbd_env_int(const char* p_name, i32 fallback)
{
    const char* p = getenv(p_name);
    if (!p || !*p) return fallback;
    char* p_end = NULL;
    const long v = strtol(p, &p_end, 10);
    return (p_end && *p_end == 0) ? (i32)v : fallback;
}

extern "C" i32 tgbd_current_device(void)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return dev;
}

extern "C" i32 tgbd_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static b32 tgbd__alloc(struct tgb_device* d)
{
    const u64 nc = d->cluster_capacity, no = d->object_capacity;
    TGB_CUDA(cudaMalloc(&d->d_cluster_pointers, nc * sizeof(u32)));
    TGB_CUDA(cudaMalloc(&d->d_c2o, nc * sizeof(u32)));
    TGB_CUDA(cudaMalloc(&d->d_objects, no * sizeof(tg_object_data)));
    d->h_objects = (u8*)calloc(no ? no : 1, sizeof(tg_object_data)); /* host shadow of the records (tgbd_flush_objects) */
    if (!d->h_objects) { tgb_set_error("out of host memory for %llu object records", (unsigned long long)no); return TG_FALSE; }
    d->objects_dirty_lo = ~0ull; d->objects_dirty_hi = 0;
    TGB_CUDA(cudaMalloc(&d->d_masks, nc * 64));
    TGB_CUDA(cudaMalloc(&d->d_lut_idx, nc * 512));
    TGB_CUDA(cudaMalloc(&d->d_color_lut, (u64)d->n_color_luts * 256 * sizeof(u32)));
    TGB_CUDA(cudaMalloc(&d->d_frames, no * sizeof(tgb_object_frame)));
    TGB_CUDA(cudaMalloc(&d->d_frames_sorted, no * sizeof(tgb_object_frame)));
    TGB_CUDA(cudaMalloc(&d->d_frames_all, no * sizeof(tgb_object_frame)));
    TGB_CUDA(cudaMalloc(&d->d_visible_count, 4 * sizeof(u32)));
    TGB_CUDA(cudaMalloc(&d->d_gi_count, 32 * sizeof(u32)));
    TGB_CUDA(cudaHostAlloc((void**)&d->h_gi_stats, 32 * sizeof(u32), cudaHostAllocMapped));
    memset(d->h_gi_stats, 0, 32 * sizeof(u32));
    {
        int n_sms = 0;
        TGB_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, d->device));
        d->n_sms = (u32)(n_sms > 0 ? n_sms : 148);
    }
    TGB_CUDA(cudaHostAlloc((void**)&d->h_visible_count, 4 * sizeof(u32), cudaHostAllocMapped));
    TGB_CUDA(cudaMemsetAsync(d->d_cluster_pointers, 0, nc * sizeof(u32), d->stream));
    TGB_CUDA(cudaMemsetAsync(d->d_c2o, 0, nc * sizeof(u32), d->stream));
    TGB_CUDA(cudaMemsetAsync(d->d_objects, 0, no * sizeof(tg_object_data), d->stream));
    TGB_CUDA(cudaMemsetAsync(d->d_masks, 0, nc * 64, d->stream));
    TGB_CUDA(cudaMemsetAsync(d->d_color_lut, 0, (u64)d->n_color_luts * 256 * sizeof(u32), d->stream));
    for (int i = 0; i < 16; i++) TGB_CUDA(cudaEventCreate(&d->ev[i]));
    TGB_CUDA(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
    TGB_CUDA(cudaStreamCreateWithFlags(&d->svo_stream, cudaStreamNonBlocking));
    TGB_CUDA(cudaEventCreateWithFlags(&d->ev_inputs, cudaEventDisableTiming));
    for (int i = 0; i < TGB_MAX_BANDS; i++)
    {
        TGB_CUDA(cudaEventCreateWithFlags(&d->ev_band[i], cudaEventDisableTiming));
        TGB_CUDA(cudaEventCreateWithFlags(&d->ev_band_copied[0][i], cudaEventDisableTiming));
        TGB_CUDA(cudaEventCreateWithFlags(&d->ev_band_copied[1][i], cudaEventDisableTiming));
    }
    for (int i = 0; i < TGB_FRAME_RING; i++) TGB_CUDA(cudaEventCreateWithFlags(&d->ev_frame_copied[i], cudaEventDisableTiming));
    d->sink_bands = 1;

    /*
     * SVO capacities. The reference reserves 2^14 nodes, 2^13 leaf records and 2^21 voxel words = 2048 blocks
     * (tg_sparse_voxel_octree.c:479-484) and asserts beyond; a 1024^3 box of 32^3 blocks can hold 37,449 nodes and
     * 32,768 leaves, which is 137 MB of a 180 GB HBM: reserve the worst case so a build can never overflow (Q12).
     */
    d->svo.node_capacity = 1u << 16;
    d->svo.leaf_capacity = 1u << 15;
    d->svo.voxel_word_capacity = 1u << 25;
    TGB_CUDA(cudaMalloc(&d->svo.d_nodes, (u64)d->svo.node_capacity * 4));
    TGB_CUDA(cudaMalloc(&d->svo.d_leaf_data, (u64)d->svo.leaf_capacity * 65 * 4));
    TGB_CUDA(cudaMalloc(&d->svo.d_voxels, (u64)d->svo.voxel_word_capacity * 4));
    TGB_CUDA(cudaMalloc(&d->svo.d_counts, 16 * sizeof(u32)));
    /* [32^3 + 1] u32 cells + completeness word, then the same cells in 16 bits */
    TGB_CUDA(cudaMalloc(&d->svo.d_top_grid, (TGB_TOP_GRID_CELLS + 1) * sizeof(u32) + TGB_TOP_GRID_CELLS * sizeof(unsigned short)));
    TGB_CUDA(cudaMemsetAsync(d->svo.d_top_grid, 0, (TGB_TOP_GRID_CELLS + 1) * sizeof(u32) + TGB_TOP_GRID_CELLS * sizeof(unsigned short), d->stream));
    TGB_CUDA(cudaMalloc(&d->svo.d_object_moved, no * sizeof(u32)));
    return TG_TRUE;
}

extern "C" b32 tgbd_resize(struct tgb_device* d, u32 width, u32 height)
{
    TGB_CUDA(cudaSetDevice(d->device));
    TGB_CUDA(cudaStreamSynchronize(d->stream));
    if (d->copy_stream) TGB_CUDA(cudaStreamSynchronize(d->copy_stream)); /* pending band copies read the buffers freed below */
    for (int i = 0; i < TGB_MAX_BANDS; i++) d->band_copy_pending[0][i] = d->band_copy_pending[1][i] = TG_FALSE;
    tgbd_p2p_teardown(d); /* collective when peer memory is mapped (a resize is mirrored on every rank); d_vis / d_mat are [0] of their pairs again */
    d->d_vis_pair[0] = NULL; d->d_mat_pair[0] = NULL;
    d->vis_merged = TG_FALSE; d->tile_merged = TG_FALSE;
    if (d->n_ranks == 0) d->n_ranks = 1;
    d->tile_rows = tgb_tile_rows_for(height, d->n_ranks);
    const u64 padded_px = (u64)width * d->tile_rows * d->n_ranks; /* >= width * height: whole bands, equal tiles; every frame buffer has this many pixels (tgb_rows.h) */
    if (d->d_mat) TGB_CUDA(cudaFree(d->d_mat));
    if (d->d_mat_tile) TGB_CUDA(cudaFree(d->d_mat_tile));
    d->d_mat = NULL; d->d_mat_tile = NULL;
    if (d->n_ranks > 1)
    {
        const u32 old_w = d->width; d->width = width; /* the tail layout (tile flags, frame counter) follows the NEW size */
        const u64 mat_bytes = tgbd_mat_bytes(d);
        d->width = old_w;
        TGB_CUDA(cudaMalloc(&d->d_mat, mat_bytes));
        TGB_CUDA(cudaMalloc(&d->d_mat_tile, (u64)width * d->tile_rows * sizeof(u64)));
        TGB_CUDA(cudaMemsetAsync(d->d_mat, 0, mat_bytes, d->stream));
        d->frame_seq = 0; d->tiles_flagged = TG_FALSE; d->objects_gathered = TG_FALSE;
    }
    if (d->d_vis) TGB_CUDA(cudaFree(d->d_vis));
    if (d->d_radiance_pair[0]) TGB_CUDA(cudaFree(d->d_radiance_pair[0]));
    if (d->d_radiance_pair[1]) TGB_CUDA(cudaFree(d->d_radiance_pair[1]));
    d->d_radiance_pair[0] = d->d_radiance_pair[1] = NULL;
    for (int k = 0; k < 2; k++) { if (d->d_present_pair[k]) TGB_CUDA(cudaFree(d->d_present_pair[k])); d->d_present_pair[k] = NULL; }
    d->radiance_flip = 0;
    if (d->d_gi_q0) TGB_CUDA(cudaFree(d->d_gi_q0));
    if (d->d_gi_q1) TGB_CUDA(cudaFree(d->d_gi_q1));
    if (d->d_gi_q2) TGB_CUDA(cudaFree(d->d_gi_q2));
    if (d->d_gi_exact) TGB_CUDA(cudaFree(d->d_gi_exact));
    d->d_gi_exact = NULL;
    d->d_vis = NULL;
    d->d_radiance = NULL;
    d->d_gi_q0 = d->d_gi_q1 = d->d_gi_q2 = NULL;
    d->width = width;
    d->height = height;
    TGB_CUDA(cudaMalloc(&d->d_vis, padded_px * sizeof(u64)));
    TGB_CUDA(cudaMalloc(&d->d_radiance_pair[0], padded_px * sizeof(float4)));
    d->d_radiance = d->d_radiance_pair[0];
    TGB_CUDA(cudaMalloc(&d->d_gi_q0, (u64)width * height * sizeof(float4)));
    TGB_CUDA(cudaMalloc(&d->d_gi_q1, (u64)width * height * sizeof(float4)));
    TGB_CUDA(cudaMalloc(&d->d_gi_q2, (u64)width * height * sizeof(float4)));
    TGB_CUDA(cudaMalloc(&d->d_gi_exact, (u64)2 * width * height * sizeof(u32))); /* two lists: handed over by the first pass, by the careful pass */
    TGB_CUDA(cudaMemsetAsync(d->d_vis, 0xFF, padded_px * sizeof(u64), d->stream));
    TGB_CUDA(cudaMemsetAsync(d->d_radiance, 0, padded_px * sizeof(float4), d->stream));
    return TG_TRUE;
}

extern "C" struct tgb_device* tgbd_create(i32 device, u32 object_capacity, u32 cluster_capacity, u32 n_color_luts, u32 width, u32 height)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
    {
        cudaGetLastError();
        tgb_set_error("tg_raytracer_create: no CUDA device (%s); libtgb200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return NULL;
    }
    if (device < 0 || device >= n)
    {
        tgb_set_error("tg_raytracer_create: device %d out of range (have %d)", device, n);
        return NULL;
    }
    if (cudaSetDevice(device) != cudaSuccess)
    {
        tgb_set_error("cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(cudaGetLastError()));
        return NULL;
    }
    struct tgb_device* d = (struct tgb_device*)calloc(1, sizeof(*d));
    d->device = device;
    d->n_ranks = 1;
    d->object_capacity = object_capacity;
    d->cluster_capacity = cluster_capacity;
    d->n_color_luts = n_color_luts ? n_color_luts : 1;
    if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess)
    {
        tgb_set_error("cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
        free(d);
        return NULL;
    }
    if (!tgbd__alloc(d) || !tgbd_resize(d, width, height))
    {
        tgbd_destroy(d);
        return NULL;
    }
    cudaStreamSynchronize(d->stream);
    return d;
}

extern "C" void tgbd_destroy(struct tgb_device* d)
{
    if (!d) return;
    cudaSetDevice(d->device);
    cudaStreamSynchronize(d->stream);
    if (d->copy_stream) cudaStreamSynchronize(d->copy_stream);
    d->p_comm = NULL;       /* not collective here: the communicator's owner (tgb200_comm_destroy) ran the collective teardown */
    tgbd_p2p_teardown(d);
    cudaFree(d->d_cluster_pointers); cudaFree(d->d_c2o); cudaFree(d->d_objects); cudaFree(d->d_masks);
    cudaFree(d->d_lut_idx); cudaFree(d->d_color_lut); cudaFree(d->d_vis); cudaFree(d->d_radiance_pair[0]); cudaFree(d->d_radiance_pair[1]); cudaFree(d->d_present_pair[0]); cudaFree(d->d_present_pair[1]); cudaFree(d->d_gi_q0); cudaFree(d->d_gi_q1); cudaFree(d->d_gi_q2); cudaFree(d->d_gi_exact); cudaFree(d->d_gi_count);
    cudaFree(d->d_frames); cudaFree(d->d_frames_sorted); cudaFree(d->d_frames_all); cudaFree(d->d_visible_count);
    if (d->h_visible_count) cudaFreeHost(d->h_visible_count);
    if (d->h_gi_stats) cudaFreeHost(d->h_gi_stats);
    free(d->h_objects);
    cudaFree(d->svo.d_nodes); cudaFree(d->svo.d_leaf_data); cudaFree(d->svo.d_voxels); cudaFree(d->svo.d_counts); cudaFree(d->svo.d_top_grid); cudaFree(d->svo.d_fast_cells); cudaFree(d->svo.d_fast_bricks); cudaFree(d->svo.d_fast_columns);
    cudaFree(d->svo.d_pairs_a); cudaFree(d->svo.d_pairs_b); cudaFree(d->svo.d_scratch); cudaFree(d->svo.d_pair_flags); cudaFree(d->svo.d_object_flags); cudaFree(d->svo.d_pair_leaf_a); cudaFree(d->svo.d_pair_leaf_b);
    cudaFree(d->svo.d_voxels_alt); cudaFree(d->svo.d_leaf_data_alt); cudaFree(d->svo.d_object_moved); cudaFree(d->svo.d_moved_indices); cudaFree(d->svo.d_part); cudaFree(d->svo.d_gather);
    cudaFree(d->d_mat); cudaFree(d->d_mat_tile); cudaFree(d->d_objects_global); cudaFree(d->d_frames_global);
    for (int i = 0; i < 16; i++) if (d->ev[i]) cudaEventDestroy(d->ev[i]);
    for (int i = 0; i < TGB_MAX_BANDS; i++) { if (d->ev_band[i]) cudaEventDestroy(d->ev_band[i]); if (d->ev_band_copied[0][i]) cudaEventDestroy(d->ev_band_copied[0][i]); if (d->ev_band_copied[1][i]) cudaEventDestroy(d->ev_band_copied[1][i]); }
    for (int i = 0; i < TGB_FRAME_RING; i++) if (d->ev_frame_copied[i]) cudaEventDestroy(d->ev_frame_copied[i]);
    if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
    if (d->svo_stream) { cudaStreamSynchronize(d->svo_stream); cudaStreamDestroy(d->svo_stream); }
    if (d->ev_inputs) cudaEventDestroy(d->ev_inputs);
    cudaStreamDestroy(d->stream);
    cudaGetLastError();
    free(d);
}

static b32 tgbd__buffer_range(struct tgb_device* d, u32 buffer, u8** pp, u64* p_size)
{
    const u64 nc = d->cluster_capacity, no = d->object_capacity, px = (u64)d->width * d->tile_rows * (d->n_ranks ? d->n_ranks : 1); /* padded frame, virtual row order */
    switch (buffer)
    {
    case TGB_BUF_CLUSTER_POINTERS: *pp = (u8*)d->d_cluster_pointers; *p_size = nc * 4; break;
    case TGB_BUF_C2O:              *pp = (u8*)d->d_c2o;              *p_size = nc * 4; break;
    case TGB_BUF_OBJECTS:          *pp = (u8*)d->d_objects;          *p_size = no * sizeof(tg_object_data); break;
    case TGB_BUF_MASKS:            *pp = (u8*)d->d_masks;            *p_size = nc * 64; break;
    case TGB_BUF_LUT_IDX:          *pp = (u8*)d->d_lut_idx;          *p_size = nc * 512; break;
    case TGB_BUF_COLOR_LUT:        *pp = (u8*)d->d_color_lut;        *p_size = (u64)d->n_color_luts * 1024; break;
    case TGB_BUF_VISIBILITY:       *pp = (u8*)d->d_vis;              *p_size = px * 8; break;
    case TGB_BUF_VISIBILITY_MERGED: *pp = (u8*)tgbd_visibility_for_read(d); *p_size = px * 8; break;
    case TGB_BUF_RADIANCE:         *pp = (u8*)d->d_radiance;         *p_size = (u64)d->width * d->tile_rows * d->n_ranks * 16; break;
    case TGB_BUF_SVO_NODES:        *pp = (u8*)d->svo.d_nodes;        *p_size = (u64)d->svo.node_capacity * 4; break;
    case TGB_BUF_SVO_LEAF_DATA:    *pp = (u8*)d->svo.d_leaf_data;    *p_size = (u64)d->svo.leaf_capacity * 260; break;
    case TGB_BUF_SVO_VOXELS:       *pp = (u8*)d->svo.d_voxels;       *p_size = (u64)d->svo.voxel_word_capacity * 4; break;
    default: tgb_set_error("unknown device buffer %u", buffer); return TG_FALSE;
    }
    return TG_TRUE;
}

extern "C" void* tgbd_buffer(struct tgb_device* d, u32 buffer)
{
    u8* p = NULL; u64 size = 0;
    tgbd_flush_objects(d);
    if (!tgbd__buffer_range(d, buffer, &p, &size)) return NULL;
    return p;
}

extern "C" void* tgbd_stream(struct tgb_device* d) { return (void*)d->stream; }

extern "C" b32 tgbd_flush_objects(struct tgb_device* d)
{
    if (d->objects_dirty_lo >= d->objects_dirty_hi) return TG_TRUE;
    const u64 lo = d->objects_dirty_lo, n = d->objects_dirty_hi - lo;
    d->objects_dirty_lo = ~0ull; d->objects_dirty_hi = 0;
    TGB_CUDA(cudaSetDevice(d->device));
    /* pageable source: staged before the call returns, so the shadow may be rewritten for the next frame while this frame is still in flight */
    TGB_CUDA(cudaMemcpyAsync((u8*)d->d_objects + lo, d->h_objects + lo, n, cudaMemcpyHostToDevice, d->stream));
    return TG_TRUE;
}

extern "C" b32 tgbd_upload(struct tgb_device* d, u32 buffer, u64 dst_offset_bytes, const void* p_src, u64 n_bytes)
{
    u8* p; u64 size;
    if (!tgbd__buffer_range(d, buffer, &p, &size)) return TG_FALSE;
    if (dst_offset_bytes + n_bytes > size) { tgb_set_error("upload past the end of device buffer %u (%llu + %llu > %llu)", buffer, (unsigned long long)dst_offset_bytes, (unsigned long long)n_bytes, (unsigned long long)size); return TG_FALSE; }
    if (buffer == TGB_BUF_OBJECTS && d->h_objects)
    {
        memcpy(d->h_objects + dst_offset_bytes, p_src, n_bytes);
        if (dst_offset_bytes < d->objects_dirty_lo) d->objects_dirty_lo = dst_offset_bytes;
        if (dst_offset_bytes + n_bytes > d->objects_dirty_hi) d->objects_dirty_hi = dst_offset_bytes + n_bytes;
        d->objects_gathered = TG_FALSE;
        d->inputs_changed_since_clear = TG_TRUE;
        return TG_TRUE;
    }
    TGB_CUDA(cudaSetDevice(d->device));
    /* pageable source: the copy is staged before the call returns, so the caller may reuse p_src */
    TGB_CUDA(cudaMemcpyAsync(p + dst_offset_bytes, p_src, n_bytes, cudaMemcpyHostToDevice, d->stream));
    if (buffer == TGB_BUF_VISIBILITY) d->tiles_flagged = TG_FALSE; /* uploaded words: the sharded shading stage resolves their materials itself */
    if (buffer == TGB_BUF_OBJECTS) d->objects_gathered = TG_FALSE;
    if (buffer != TGB_BUF_VISIBILITY && buffer != TGB_BUF_RADIANCE) d->inputs_changed_since_clear = TG_TRUE; /* scene data uploaded after tgbd_clear: K2 must not start before it */
    return TG_TRUE;
}

extern "C" b32 tgbd_download(struct tgb_device* d, u32 buffer, u64 src_offset_bytes, void* p_dst, u64 n_bytes)
{
    u8* p; u64 size;
    if (!tgbd__buffer_range(d, buffer, &p, &size)) return TG_FALSE;
    if (src_offset_bytes + n_bytes > size) { tgb_set_error("download past the end of device buffer %u", buffer); return TG_FALSE; }
    if (!tgbd_flush_objects(d)) return TG_FALSE;
    TGB_CUDA(cudaSetDevice(d->device));
    TGB_CUDA(cudaMemcpyAsync(p_dst, p + src_offset_bytes, n_bytes, cudaMemcpyDeviceToHost, d->stream));
    TGB_CUDA(cudaStreamSynchronize(d->stream));
    return TG_TRUE;
}

extern "C" b32 tgbd_move(struct tgb_device* d, u32 buffer, u64 dst_offset_bytes, u64 src_offset_bytes, u64 n_bytes)
{
    u8* p; u64 size;
    if (!tgbd__buffer_range(d, buffer, &p, &size)) return TG_FALSE;
    if (dst_offset_bytes + n_bytes > size || src_offset_bytes + n_bytes > size) { tgb_set_error("move past the end of device buffer %u", buffer); return TG_FALSE; }
    if (n_bytes == 0) return TG_TRUE;
    TGB_CUDA(cudaSetDevice(d->device));
    /* ranges may overlap (shift down): bounce through a temporary */
    void* p_tmp = NULL;
    TGB_CUDA(cudaMalloc(&p_tmp, n_bytes));
    TGB_CUDA(cudaMemcpyAsync(p_tmp, p + src_offset_bytes, n_bytes, cudaMemcpyDeviceToDevice, d->stream));
    TGB_CUDA(cudaMemcpyAsync(p + dst_offset_bytes, p_tmp, n_bytes, cudaMemcpyDeviceToDevice, d->stream));
    TGB_CUDA(cudaStreamSynchronize(d->stream));
    TGB_CUDA(cudaFree(p_tmp));
    return TG_TRUE;
}

extern "C" void tgbd_synchronize(struct tgb_device* d)
{
    cudaSetDevice(d->device);
    tgbd_flush_objects(d);
    cudaError_t e = cudaStreamSynchronize(d->stream);
    if (e == cudaSuccess && d->copy_stream) e = cudaStreamSynchronize(d->copy_stream);
    if (e != cudaSuccess) tgb_set_error("cudaStreamSynchronize -> %s", cudaGetErrorString(e));
    else tgbd_p2p_check(d);
}

/* ---- frame sink ---- */
extern "C" b32 tgbd_set_frame_sink(struct tgb_device* d, void* p_host, u32 n_bands, u32 format)
{
    d->p_sink = (f32*)p_host;
    d->sink_format = format;
    d->sink_bands = n_bands < 1 ? 1 : (n_bands > TGB_MAX_BANDS ? TGB_MAX_BANDS : n_bands);
    return TG_TRUE;
}

extern "C" u64 tgbd_frames_sunk(struct tgb_device* d) { return d->n_frames_sunk; }

extern "C" b32 tgbd_wait_frame(struct tgb_device* d, u64 ticket)
{
    TGB_CUDA(cudaSetDevice(d->device));
    if (ticket == 0 || ticket > d->n_frames_sunk) { tgb_set_error("wait_frame: ticket %llu was never issued (%llu frames so far)", (unsigned long long)ticket, (unsigned long long)d->n_frames_sunk); return TG_FALSE; }
    if (d->n_frames_sunk - ticket >= TGB_FRAME_RING) { TGB_CUDA(cudaStreamSynchronize(d->copy_stream)); return TG_TRUE; } /* older than the ring: everything up to now */
    TGB_CUDA(cudaEventSynchronize(d->ev_frame_copied[ticket % TGB_FRAME_RING]));
    return TG_TRUE;
}

extern "C" void tgbd_set_shard(struct tgb_device* d, u32 global_pointer_base) { d->global_pointer_base = global_pointer_base; }
extern "C" void tgbd_set_gi_traversal(struct tgb_device* d, u32 kind) { d->gi_traversal = kind; }

extern "C" b32 tgbd_set_comm(struct tgb_device* d, void* p_comm, u32 rank, u32 n_ranks)
{
    TGB_CUDA(cudaSetDevice(d->device));
    TGB_CUDA(cudaStreamSynchronize(d->stream));
    tgbd_p2p_teardown(d); /* with the OLD communicator still in place */
    d->p_comm = p_comm;
    d->rank = p_comm ? rank : 0;
    d->n_ranks = p_comm ? n_ranks : 1;
    if (d->d_objects_global) TGB_CUDA(cudaFree(d->d_objects_global));
    if (d->d_frames_global) TGB_CUDA(cudaFree(d->d_frames_global));
    d->d_objects_global = NULL; d->d_frames_global = NULL;
    if (d->n_ranks > 1)
    {
        TGB_CUDA(cudaMalloc(&d->d_objects_global, (u64)d->n_ranks * d->object_capacity * sizeof(tg_object_data)));
        TGB_CUDA(cudaMalloc(&d->d_frames_global, (u64)d->n_ranks * d->object_capacity * sizeof(tgb_object_frame)));
        /* slots no rank has published yet must read as uninitialised objects (all dims 0), not as garbage */
        TGB_CUDA(cudaMemsetAsync(d->d_objects_global, 0, (u64)d->n_ranks * d->object_capacity * sizeof(tg_object_data), d->stream));
        TGB_CUDA(cudaMemsetAsync(d->d_frames_global, 0, (u64)d->n_ranks * d->object_capacity * sizeof(tgb_object_frame), d->stream));
    }
    return tgbd_resize(d, d->width, d->height); /* tile-sized buffers depend on n_ranks */
}

extern "C" u32 tgbd_tile_rows(struct tgb_device* d) { return d->tile_rows; }
extern "C" u64 tgbd_padded_pixels(struct tgb_device* d) { return (u64)d->width * d->tile_rows * (d->n_ranks ? d->n_ranks : 1); }
extern "C" void* tgbd_comm(struct tgb_device* d) { return d->p_comm; }
extern "C" u32 tgbd_rank(struct tgb_device* d) { return d->rank; }
extern "C" u32 tgbd_n_ranks(struct tgb_device* d) { return d->n_ranks ? d->n_ranks : 1; }

extern "C" void tgbd_reset_launch_counter(struct tgb_device* d) { d->n_kernel_launches = 0; }

extern "C" void tgbd_get_timings(struct tgb_device* d, tgb200_timings* p_out)
{
    cudaSetDevice(d->device);
    tgbd_flush_objects(d);
    cudaStreamSynchronize(d->stream);
    tgbd_p2p_check(d);
    if (d->ev_clear) cudaEventElapsedTime(&d->clear_ms, d->ev[0], d->ev[1]);
    if (d->ev_vis)   { cudaEventElapsedTime(&d->cull_ms, d->ev[2], d->ev[3]); cudaEventElapsedTime(&d->visibility_ms, d->ev[3], d->ev[4]); }
    if (d->ev_svo)   cudaEventElapsedTime(&d->svo_ms, d->ev[5], d->ev[6]);
    if (d->ev_shade) cudaEventElapsedTime(&d->shading_ms, d->ev[7], d->ev[8]);
    if (d->ev_merge) cudaEventElapsedTime(&d->merge_ms, d->ev[9], d->ev[10]);
    if (d->ev_merge && d->ev_merge_parts)
    {
        /* peer-memory merge: publish "K1 done" (+ material resolve when K1's epilogue did not run) | wait for the slowest rank's K1 | k_merge_tile */
        cudaEventElapsedTime(&d->merge_resolve_ms, d->ev[9], d->ev[11]);
        cudaEventElapsedTime(&d->merge_gather_ms, d->ev[11], d->ev[12]);
        cudaEventElapsedTime(&d->merge_kernel_ms, d->ev[12], d->ev[10]);
    }
    else d->merge_resolve_ms = d->merge_gather_ms = d->merge_kernel_ms = 0.0f;
    cudaGetLastError();
    p_out->clear_ms = d->clear_ms;
    p_out->cull_ms = d->cull_ms;
    p_out->visibility_ms = d->visibility_ms;
    p_out->svo_ms = d->svo_ms;
    p_out->shading_ms = d->shading_ms;
    p_out->merge_ms = d->merge_ms;
    p_out->merge_resolve_ms = d->merge_resolve_ms;
    p_out->merge_gather_ms = d->merge_gather_ms;
    p_out->merge_kernel_ms = d->merge_kernel_ms;
    p_out->pad2 = 0;
    /* the counters live on the device during the frames (no per-frame read-back); fetch them now */
    if (d->ev_vis) { cudaMemcpy(d->h_visible_count, d->d_visible_count, 2 * sizeof(u32), cudaMemcpyDeviceToHost); d->n_visible_objects = d->h_visible_count[0]; }
    if (d->gi_stats_valid) cudaMemcpy(d->h_gi_stats, d->d_gi_count, 32 * sizeof(u32), cudaMemcpyDeviceToHost);
    p_out->n_visible_objects = d->n_visible_objects;
    if (d->gi_stats_valid && d->h_gi_stats[20]) tgb_set_error("k_gi_trace_list: a bulk copy of a leaf block never completed (mbarrier wait gave up); the frame's GI term is incomplete");
    p_out->n_gi_rays = d->h_gi_stats[10];
    p_out->n_gi_rays_exact = d->h_gi_stats[14];
    p_out->n_gi_node_visits = ((const u64*)d->h_gi_stats)[1];
    p_out->n_gi_dda_steps = ((const u64*)d->h_gi_stats)[2];
    p_out->n_gi_advances = ((const u64*)d->h_gi_stats)[3];
    if (getenv("TGB200_GI_HISTOGRAM"))
    {
        fprintf(stderr, "[tgb200] GI rays: max node visits %u; log2 histogram:", d->h_gi_stats[8]);
        for (int i = 0; i < 13; i++) fprintf(stderr, " %u", d->h_gi_stats[16 + i]);
        fprintf(stderr, "\n");
    }
    p_out->n_kernel_launches = d->n_kernel_launches;
}

extern "C" void tgbd_merge_begin(struct tgb_device* d) { cudaSetDevice(d->device); cudaEventRecord(d->ev[9], d->stream); d->ev_merge_parts = TG_FALSE; }
extern "C" void tgbd_merge_end(struct tgb_device* d) { cudaEventRecord(d->ev[10], d->stream); d->ev_merge = TG_TRUE; }

/* (8*rel_x + vx) % 256, tgvk_raytracer.c:947-978; one thread per 16 voxels (uint4 stores) */
__global__ void k_fill_default_lut_idx(const u32* __restrict__ p_cluster_pointers, u8* __restrict__ p_lut_idx, u32 first_pointer, u32 n_cluster_pointers, u32 nx)
{
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 n_vec = (u64)n_cluster_pointers * 32; /* 512 B / 16 B */
    if (t >= n_vec) return;
    const u32 rel = (u32)(t / 32);
    const u32 part = (u32)(t % 32);      /* voxels [16*part, 16*part+16): two x-rows */
    const u32 rel_x = rel % nx;
    const u32 cluster_idx = p_cluster_pointers[first_pointer + rel];
    u32 w[4];
#pragma unroll
    for (u32 k = 0; k < 4; k++)
    {
        u32 word = 0;
#pragma unroll
        for (u32 b = 0; b < 4; b++)
        {
            const u32 vx = (k * 4 + b) % 8;
            word |= ((8u * rel_x + vx) % 256u) << (8 * b);
        }
        w[k] = word;
    }
    uint4* p_dst = (uint4*)(p_lut_idx + (u64)cluster_idx * 512) + part;
    *p_dst = make_uint4(w[0], w[1], w[2], w[3]);
}

extern "C" b32 tgbd_fill_default_lut_idx(struct tgb_device* d, u32 first_pointer, u32 n_cluster_pointers, u32 nx)
{
    if (n_cluster_pointers == 0) return TG_TRUE;
    TGB_CUDA(cudaSetDevice(d->device));
    const u64 n_vec = (u64)n_cluster_pointers * 32;
    k_fill_default_lut_idx<<<(u32)((n_vec + 255) / 256), 256, 0, d->stream>>>(d->d_cluster_pointers, d->d_lut_idx, first_pointer, n_cluster_pointers, nx);
    TGB_LAUNCH_CHECK(d);
    return TG_TRUE;
}

/*
 * tgb_peer.cu -- the multi-GPU exchange of the frame as ONE kernel over peer memory (NVLink / NVSwitch).
 *
 * With NCCL alone a sharded frame needs two 66 MB collectives at 4K: ncclAllReduce(u64, min) over the visibility buffers, so
 * that every rank learns which pixels it won, and ncclReduceScatter(u64, max) over the material words the winners then
 * resolve (tgb_nccl.c; the 512 B / cluster material data exists on the owner only). But a rank only SHADES its own screen
 * tile, so all it needs is, per pixel of that tile, the minimum of the ranks' words and the material word of the rank that
 * holds the minimum. Here every rank resolves the material of its LOCAL winners right after K1 (k_resolve_material on the
 * unmerged buffer: every hit is its own), maps the other ranks' visibility and material buffers into its address space
 * (CUDA IPC handles, exchanged once through the communicator) and runs k_merge_tile: N coalesced 8-byte loads per pixel --
 * N - 1 of them over NVLink -- a running min, one more load from the winner. 58 MB cross the switch per rank at N = 8
 * instead of two all-to-all collectives, and nothing is written that another rank reads.
 *
 * What crosses the switch is cut further by K1 itself: k_visibility<.., SHARDED> records per 16x16 tile whether the rank has any hit
 * there; the material pass only visits those tiles (k_resolve_material_tiles) and k_merge_tile reads a peer's words only for tiles
 * that peer flagged -- with the objects dealt out over the ranks a tile is covered by one or two ranks, not N.
 *
 * Synchronisation: both buffers are double buffered (flipped by tgbd_clear) and carry, behind the material words, the tile flags
 * and a frame counter. A rank PUBLISHES the frame number when its K1 is done (k_signal: system-scope fence, then the store) and
 * WAITS until every peer has published it (k_wait_peers polls the peers' counters over NVLink) before k_merge_tile runs -- no
 * collective sits in the frame any more: the 96-byte object records the shading stage needs from every rank (objects may have moved)
 * are published in the same tail before the counter and collected from the peers after the wait (k_collect_objects). Why two buffers suffice: a rank that starts frame i + 2 (and clears the buffer
 * of frame i) has finished its merge of frame i + 1, which waited for every peer's K1 of frame i + 1, which those peers started
 * after their merge of frame i -- the last reader of that buffer. The min is associative and commutative, so the merged tile is
 * bit-identical to the all-reduce and to a single-GPU frame (tests/test_multi_gpu.py runs both paths; bench.py re-checks it at
 * full size inside its warm-up).
 * Peer loads use ld.volatile semantics (__ldcv): a word another GPU wrote this frame must not come from a stale line.
 */
#include "tgb_device.cuh"

struct tgb_peer_table
{
    const u64* p_vis[TGB_MAX_RANKS];
    const u64* p_mat[TGB_MAX_RANKS];
};

struct tgb_peer_flags
{
    const u32* p_tile_flags[TGB_MAX_RANKS]; /* NULL table entry 0 = no flags: read every rank's words */
};

/* pixels [first, first + n) of the frame (virtual row order), every rank's words: merged word -> p_vis_out[first + i] (read-back / picking) */
__global__ void __launch_bounds__(256) k_merge_linear(const tgb_peer_table t, u32 n_ranks, u64 first, u64 n, u64* __restrict__ p_vis_out)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 best = TG_VIS_CLEAR;
    for (u32 r = 0; r < n_ranks; r++)
    {
        const u64 v = __ldcv(&t.p_vis[r][first + i]);
        best = v < best ? v : best;
    }
    p_vis_out[first + i] = best;
}

/*
 * This rank's screen tile, one CTA per 16x16 pixel tile (blockIdx.y = band of the rank's tile, virtual rows first_row + 16 y ...):
 * the CTA first fetches every rank's flag for the tile (N independent 4-byte loads), and its pixels then read words only from the
 * ranks that have a hit there -- none at all for a tile of sky. Merged word -> p_vis_out[pixel] (whole-frame pixel index), the
 * winner's material word -> p_mat_out[pixel - first_row * w].
 */
__global__ void __launch_bounds__(256) k_merge_tile(const tgb_peer_table t, const tgb_peer_flags fl, u32 n_ranks, u32 w, u32 tiles_x, u32 first_row,
                                                    u64* __restrict__ p_vis_out, u64* __restrict__ p_mat_out)
{
    __shared__ u32 s_ranks;
    const u32 vy0 = first_row + blockIdx.y * TGB_BAND_ROWS;
    const u32 tile = (vy0 / TGB_BAND_ROWS) * tiles_x + blockIdx.x;
    if (threadIdx.x == 0) s_ranks = 0u;
    __syncthreads();
    if (threadIdx.x < n_ranks && __ldcv(&fl.p_tile_flags[threadIdx.x][tile]) != 0u) atomicOr(&s_ranks, 1u << threadIdx.x);
    __syncthreads();
    const u32 x = blockIdx.x * 16u + (threadIdx.x & 15u), vy = vy0 + (threadIdx.x >> 4);
    if (x >= w) return;
    const u64 pixel = (u64)vy * w + x;
    u64 best = TG_VIS_CLEAR, mat = 0;
    u32 who = 0;
    for (u32 m = s_ranks; m != 0; m &= m - 1u)
    {
        const u32 r = (u32)__ffs((int)m) - 1u;
        const u64 v = __ldcv(&t.p_vis[r][pixel]);
        if (v < best) { best = v; who = r; } /* shards own disjoint pointer ranges: two ranks never hold the same word */
    }
    if (best != TG_VIS_CLEAR) mat = __ldcv(&t.p_mat[who][pixel]);
    p_vis_out[pixel] = best;
    p_mat_out[pixel - (u64)first_row * w] = mat;
}

__global__ void k_fill_words(u32* __restrict__ p, u32 n, u32 value)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = value;
}

/* local object records with pointers globalised, into this rank's slice of the global table (the LUT-independent fields are copied) */
__global__ void k_globalize_objects(const tg_object_data* __restrict__ p_objects, u32 object_capacity, u32 global_pointer_base, tg_object_data* __restrict__ p_out)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= object_capacity) return;
    tg_object_data o = p_objects[i];
    if (o.n_cluster_pointers_per_dim.x != 0 && o.n_cluster_pointers_per_dim.y != 0 && o.n_cluster_pointers_per_dim.z != 0) o.first_cluster_pointer += global_pointer_base;
    p_out[i] = o;
}

/*
 * What a rank publishes for the peers' shading stage: the 96-byte records (pointers globalised) of the objects that can have won a
 * pixel -- the survivors of its cull (p_count[0] of them, named by the front-to-back sorted frames), or every object slot when no cull
 * ran for the words in the buffer (p_frames == NULL). One thread per published object.
 */
__global__ void k_publish_objects(const tg_object_data* __restrict__ p_objects, u32 object_capacity, u32 global_pointer_base, const tgb_object_frame* __restrict__ p_frames,
                                  const u32* __restrict__ p_count, tg_object_data* __restrict__ p_out, u32* __restrict__ p_out_idx, u32* __restrict__ p_out_count)
{
    const u32 n = p_frames ? min(p_count[0], object_capacity) : object_capacity;
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *p_out_count = n;
    if (i >= n) return;
    const u32 object_idx = p_frames ? p_frames[i].object_idx : i;
    tg_object_data o = p_objects[object_idx];
    if (o.n_cluster_pointers_per_dim.x != 0 && o.n_cluster_pointers_per_dim.y != 0 && o.n_cluster_pointers_per_dim.z != 0) o.first_cluster_pointer += global_pointer_base;
    p_out[i] = o;
    p_out_idx[i] = object_idx;
}

/* every rank's published records (tail of its material buffer of this frame) -> the dense global table [n_ranks * capacity]: blockIdx.y = rank */
struct tgb_peer_objects { const tg_object_data* p_records[TGB_MAX_RANKS]; const u32* p_indices[TGB_MAX_RANKS]; const u32* p_counts[TGB_MAX_RANKS]; };
__global__ void k_collect_objects(const tgb_peer_objects src, u32 object_capacity, tg_object_data* __restrict__ p_out)
{
    /* 96-byte records as 24 words each; volatile loads: the peers wrote them this frame */
    const u32 r = blockIdx.y, words = (u32)(sizeof(tg_object_data) / 4u);
    const u32 n = min(__ldcv(src.p_counts[r]), object_capacity);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n * words; i += gridDim.x * blockDim.x)
    {
        const u32 entry = i / words, k = i - entry * words;
        const u32 object_idx = __ldcv(&src.p_indices[r][entry]);
        if (object_idx < object_capacity) reinterpret_cast<u32*>(p_out + (u64)r * object_capacity + object_idx)[k] = __ldcv(reinterpret_cast<const u32*>(src.p_records[r] + entry) + k);
    }
}

/* "this rank's K1 (and material words) of frame `seq` are complete": everything written before is visible system-wide first */
__global__ void k_signal(u32* __restrict__ p_signal, u32 seq)
{
    __threadfence_system();
    *reinterpret_cast<volatile u32*>(p_signal) = seq;
}

/*
 * Lane r waits until rank r has published frame `seq` (wrap-safe comparison). Bounded: a peer that never arrives (a crashed
 * process) must not hang this GPU for ever -- after ~4 s the kernel gives up and raises the error word the host checks at the
 * next synchronisation point.
 */
__global__ void k_wait_peers(const tgb_peer_flags signals, u32 n_ranks, u32 seq, u32* __restrict__ p_error)
{
    const u32 r = threadIdx.x;
    if (r >= n_ranks) return;
    const volatile u32* p = reinterpret_cast<const volatile u32*>(signals.p_tile_flags[r]);
    const long long t0 = clock64();
    while ((i32)(*p - seq) < 0)
    {
        __nanosleep(64);
        if (clock64() - t0 > 8000000000ll) { atomicExch(p_error, 1u + r); break; }
    }
    __threadfence_system();
}

extern "C" void tgbd_set_merge_kind(struct tgb_device* d, u32 kind) { d->merge_kind = kind; }
extern "C" void tgbd_note_merged(struct tgb_device* d) { d->vis_merged = TG_TRUE; }

static void tgbd__p2p_close(struct tgb_device* d)
{
    for (u32 f = 0; f < 2; f++)
    {
        for (u32 r = 0; r < TGB_MAX_RANKS; r++)
        {
            if (r != d->rank && d->peer_vis[f][r]) cudaIpcCloseMemHandle(d->peer_vis[f][r]);
            if (r != d->rank && d->peer_mat[f][r]) cudaIpcCloseMemHandle(d->peer_mat[f][r]);
            d->peer_vis[f][r] = NULL; d->peer_mat[f][r] = NULL;
        }
    }
    cudaGetLastError();
}

/*
 * Collective. Before a rank frees buffers its peers have mapped (resize, communicator teardown): every rank first finishes its
 * own reads (stream sync), then a tiny all-reduce makes sure all of them did, then the mappings go.
 */
extern "C" void tgbd_p2p_teardown(struct tgb_device* d)
{
    if (!d->p2p_ready && !d->d_vis_pair[1] && !d->d_vis_full && !d->d_vis_tile) { d->p2p_failed = TG_FALSE; return; }
    cudaSetDevice(d->device);
    if (d->p2p_ready && d->p_comm && d->d_ipc_stage)
    {
        tgbn_allreduce_sum_u32(d->p_comm, d->d_ipc_stage, 1, d->stream);
        cudaStreamSynchronize(d->stream);
    }
    tgbd__p2p_close(d);
    /* back to the single buffers tgbd_resize owns */
    if (d->d_vis_pair[1]) cudaFree(d->d_vis_pair[1]);
    if (d->d_mat_pair[1]) cudaFree(d->d_mat_pair[1]);
    if (d->d_vis_full) cudaFree(d->d_vis_full);
    if (d->d_vis_tile) cudaFree(d->d_vis_tile);
    if (d->d_ipc_stage) cudaFree(d->d_ipc_stage);
    d->d_vis_pair[1] = NULL; d->d_mat_pair[1] = NULL; d->d_vis_full = NULL; d->d_vis_tile = NULL; d->d_ipc_stage = NULL;
    if (d->d_vis_pair[0]) d->d_vis = d->d_vis_pair[0];
    if (d->d_mat_pair[0]) d->d_mat = d->d_mat_pair[0];
    d->vis_flip = 0;
    d->p2p_ready = TG_FALSE; d->p2p_failed = TG_FALSE;
    cudaGetLastError();
}

/*
 * Collective on first use (every rank calls it at the same point of its frame): allocates the second buffer pair, exchanges the
 * four IPC handles through the communicator, maps the peers' buffers. All ranks agree on the outcome (a sum of failure flags):
 * if any mapping failed everywhere falls back to the NCCL collectives, which are the same function of the same inputs.
 */
extern "C" b32 tgbd_p2p_prepare(struct tgb_device* d)
{
    if (d->n_ranks < 2 || !d->p_comm || d->merge_kind == 1) return TG_FALSE;
    if (d->p2p_ready) return TG_TRUE;
    if (d->p2p_failed) return TG_FALSE;
    if (d->n_ranks > TGB_MAX_RANKS) { d->p2p_failed = TG_TRUE; return TG_FALSE; }
    TGB_CUDA(cudaSetDevice(d->device));
    const u64 padded_px = (u64)d->width * d->tile_rows * d->n_ranks, px = padded_px; /* every frame buffer is padded to whole tiles */
    /*
     * From here on every rank MUST reach both collectives below whatever happens locally (a rank that returned early would leave its
     * peers blocked in ncclAllGather): local CUDA / allocation failures only count as failed mappings, and the outcome all ranks
     * agree on is the sum of those counts.
     */
    u32 n_failed = 0;
#define TGB_P2P_TRY(call) do { if ((call) != cudaSuccess) { n_failed++; cudaGetLastError(); } } while (0)
    d->d_vis_pair[0] = d->d_vis; d->d_mat_pair[0] = d->d_mat; d->vis_flip = 0;
    TGB_P2P_TRY(cudaMalloc(&d->d_vis_pair[1], px * sizeof(u64)));
    TGB_P2P_TRY(cudaMalloc(&d->d_mat_pair[1], tgbd_mat_bytes(d)));
    TGB_P2P_TRY(cudaMalloc(&d->d_vis_tile, (u64)d->width * d->tile_rows * sizeof(u64)));
    if (!d->d_ipc_stage && cudaMalloc(&d->d_ipc_stage, (u64)d->n_ranks * 4u * sizeof(cudaIpcMemHandle_t) + 64u) != cudaSuccess)
    {
        /* without the staging buffer the collectives themselves cannot run on any rank that got it; this one failure is fatal for the
         * exchange and is reported (the other ranks would wait): there is nothing smaller to allocate instead */
        cudaGetLastError();
        d->d_ipc_stage = NULL;
        tgb_set_error("p2p_prepare: out of device memory for the %u-byte handle staging buffer", (u32)(d->n_ranks * 4u * sizeof(cudaIpcMemHandle_t)));
        if (d->d_vis_pair[1]) cudaFree(d->d_vis_pair[1]);
        if (d->d_mat_pair[1]) cudaFree(d->d_mat_pair[1]);
        if (d->d_vis_tile) cudaFree(d->d_vis_tile);
        d->d_vis_pair[1] = NULL; d->d_mat_pair[1] = NULL; d->d_vis_tile = NULL;
        d->p2p_failed = TG_TRUE;
        return TG_FALSE;
    }
    if (d->d_vis_pair[1]) TGB_P2P_TRY(cudaMemsetAsync(d->d_vis_pair[1], 0xFF, px * sizeof(u64), d->stream));
    if (d->d_mat_pair[1]) TGB_P2P_TRY(cudaMemsetAsync(d->d_mat_pair[1], 0, tgbd_mat_bytes(d), d->stream));

    cudaIpcMemHandle_t mine[4];
    /* test hook: TGB200_FAIL_P2P_ON_RANK=r makes rank r report a failed mapping, which must send EVERY rank to the NCCL path */
    if (getenv("TGB200_FAIL_P2P_ON_RANK") && (u32)atoi(getenv("TGB200_FAIL_P2P_ON_RANK")) == d->rank) n_failed++;
    void* p_mine[4] = { d->d_vis_pair[0], d->d_vis_pair[1], d->d_mat_pair[0], d->d_mat_pair[1] };
    for (int k = 0; k < 4; k++)
    {
        if (!p_mine[k] || cudaIpcGetMemHandle(&mine[k], p_mine[k]) != cudaSuccess) { n_failed++; memset(&mine[k], 0, sizeof(mine[k])); cudaGetLastError(); }
    }
    cudaIpcMemHandle_t* p_all = (cudaIpcMemHandle_t*)malloc((size_t)d->n_ranks * sizeof(mine));
    if (!p_all) n_failed++;
    b32 collectives_ok = TG_TRUE;
    TGB_P2P_TRY(cudaMemcpyAsync(d->d_ipc_stage + (u64)d->rank * sizeof(mine), mine, sizeof(mine), cudaMemcpyHostToDevice, d->stream));
    if (!tgbn_allgather_bytes(d->p_comm, d->d_ipc_stage + (u64)d->rank * sizeof(mine), d->d_ipc_stage, sizeof(mine), d->stream)) collectives_ok = TG_FALSE;
    if (p_all)
    {
        TGB_P2P_TRY(cudaMemcpyAsync(p_all, d->d_ipc_stage, (u64)d->n_ranks * sizeof(mine), cudaMemcpyDeviceToHost, d->stream));
        TGB_P2P_TRY(cudaStreamSynchronize(d->stream));
    }
    for (u32 r = 0; r < d->n_ranks; r++)
    {
        for (u32 k = 0; k < 4; k++)
        {
            void* p = NULL;
            if (r == d->rank) p = p_mine[k];
            else if (n_failed == 0 && collectives_ok && cudaIpcOpenMemHandle(&p, p_all[r * 4u + k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { p = NULL; n_failed++; cudaGetLastError(); }
            if (k < 2) d->peer_vis[k][r] = (u64*)p; else d->peer_mat[k - 2][r] = (u64*)p;
        }
    }
    free(p_all);
    /* agree: the stage buffer's first word becomes the number of failed mappings over all ranks */
    u32 total_failed = 1;
    TGB_P2P_TRY(cudaMemcpyAsync(d->d_ipc_stage, &n_failed, sizeof(u32), cudaMemcpyHostToDevice, d->stream));
    if (!tgbn_allreduce_sum_u32(d->p_comm, d->d_ipc_stage, 1, d->stream)) collectives_ok = TG_FALSE;
    if (cudaMemcpyAsync(&total_failed, d->d_ipc_stage, sizeof(u32), cudaMemcpyDeviceToHost, d->stream) != cudaSuccess || cudaStreamSynchronize(d->stream) != cudaSuccess)
    {
        cudaGetLastError();
        total_failed = 1;
    }
#undef TGB_P2P_TRY
    if (total_failed || n_failed || !collectives_ok)
    {
        /* every early-out path ends here: mappings closed, the partial allocations freed, the failure remembered (no retry, no leak) */
        tgbd__p2p_close(d);
        if (d->d_vis_pair[1]) cudaFree(d->d_vis_pair[1]);
        if (d->d_mat_pair[1]) cudaFree(d->d_mat_pair[1]);
        if (d->d_vis_tile) cudaFree(d->d_vis_tile);
        if (d->d_ipc_stage) cudaFree(d->d_ipc_stage);
        d->d_vis_pair[1] = NULL; d->d_mat_pair[1] = NULL; d->d_vis_tile = NULL; d->d_ipc_stage = NULL;
        d->p2p_failed = TG_TRUE;
        cudaGetLastError();
        if (getenv("TGB200_VERBOSE")) fprintf(stderr, "[tgb200] rank %u: peer memory unavailable (%u failed mappings), using the NCCL collectives\n", d->rank, total_failed);
        return TG_FALSE;
    }
    cudaMemsetAsync(d->d_ipc_stage + (u64)d->n_ranks * 4u * sizeof(cudaIpcMemHandle_t), 0, 64, d->stream); /* k_wait_peers' error word */
    d->frame_seq = 1; /* the frame in flight (its clear ran before the mapping existed) publishes 1; every counter starts at 0 */
    d->p2p_ready = TG_TRUE;
    return TG_TRUE;
}

static tgb_peer_table tgbd__peer_table(struct tgb_device* d)
{
    tgb_peer_table t;
    for (u32 r = 0; r < TGB_MAX_RANKS; r++)
    {
        t.p_vis[r] = r < d->n_ranks ? d->peer_vis[d->vis_flip][r] : NULL;
        t.p_mat[r] = r < d->n_ranks ? d->peer_mat[d->vis_flip][r] : NULL;
    }
    return t;
}

/* the peers' tile flags (tail of their material buffers of the current flip), or their frame counters */
static tgb_peer_flags tgbd__peer_tails(struct tgb_device* d, bool signals)
{
    tgb_peer_flags f;
    for (u32 r = 0; r < TGB_MAX_RANKS; r++)
    {
        const u64* p_mat = r < d->n_ranks ? d->peer_mat[d->vis_flip][r] : NULL;
        f.p_tile_flags[r] = p_mat ? (signals ? tgbd_mat_signal(d, p_mat) : tgbd_mat_tile_flags(d, p_mat)) : NULL;
    }
    return f;
}

/* the error word of k_wait_peers lives behind the IPC handle staging area */
static u32* tgbd__wait_error(struct tgb_device* d) { return (u32*)(d->d_ipc_stage + (u64)d->n_ranks * 4u * sizeof(cudaIpcMemHandle_t)); }

/* the 96-byte object records of every rank (pointers globalised by the owner) through NCCL: collective, every rank calls it at the same
 * point of its frame (the NCCL exchange path; the peer-memory path publishes them in tgbd_p2p_barrier instead) */
extern "C" b32 tgbd_gather_objects(struct tgb_device* d)
{
    if (!d->p_comm || d->n_ranks < 2) return TG_TRUE;
    if (!tgbd_flush_objects(d)) return TG_FALSE;
    const u32 cap = d->object_capacity;
    k_globalize_objects<<<(cap + 127) / 128, 128, 0, d->stream>>>(d->d_objects, cap, d->global_pointer_base, d->d_objects_global + (u64)d->rank * cap);
    TGB_LAUNCH_CHECK(d);
    if (!tgbn_allgather_bytes(d->p_comm, d->d_objects_global + (u64)d->rank * cap, d->d_objects_global, (u64)cap * sizeof(tg_object_data), d->stream)) return TG_FALSE;
    d->objects_gathered = TG_TRUE;
    return TG_TRUE;
}

/*
 * Publish this rank's frame -- the records (pointers globalised) of the objects that survived its cull into the tail of the current
 * material buffer, then the frame counter behind a system-scope fence --, wait for every peer's counter, and collect the peers' object records. Everything the
 * shading stage needs from the other ranks then sits in this GPU's memory or is read by k_merge_tile; no NCCL call in the frame.
 * Asynchronous on the stream.
 */
extern "C" b32 tgbd_p2p_barrier(struct tgb_device* d)
{
    if (!d->p2p_ready) { tgb_set_error("p2p_barrier: peer memory is not mapped"); return TG_FALSE; }
    if (!tgbd_flush_objects(d)) return TG_FALSE;
    const u32 cap = d->object_capacity;
    k_publish_objects<<<(cap + 127) / 128, 128, 0, d->stream>>>(d->d_objects, cap, d->global_pointer_base, d->tiles_flagged ? d->d_frames_sorted : NULL, d->d_visible_count,
                                                              tgbd_mat_objects(d, d->d_mat), tgbd_mat_object_indices(d, d->d_mat), tgbd_mat_signal(d, d->d_mat) + 1);
    TGB_LAUNCH_CHECK(d);
    k_signal<<<1, 1, 0, d->stream>>>(tgbd_mat_signal(d, d->d_mat), d->frame_seq);
    TGB_LAUNCH_CHECK(d);
    TGB_CUDA(cudaEventRecord(d->ev[11], d->stream));
    k_wait_peers<<<1, 32, 0, d->stream>>>(tgbd__peer_tails(d, true), d->n_ranks, d->frame_seq, tgbd__wait_error(d));
    TGB_LAUNCH_CHECK(d);
    TGB_CUDA(cudaEventRecord(d->ev[12], d->stream));
    tgb_peer_objects src;
    for (u32 r = 0; r < TGB_MAX_RANKS; r++)
    {
        const u64* p_mat = r < d->n_ranks ? d->peer_mat[d->vis_flip][r] : NULL;
        src.p_records[r] = p_mat ? tgbd_mat_objects(d, p_mat) : NULL;
        src.p_indices[r] = p_mat ? tgbd_mat_object_indices(d, p_mat) : NULL;
        src.p_counts[r] = p_mat ? tgbd_mat_signal(d, p_mat) + 1 : NULL;
    }
    k_collect_objects<<<dim3(8, d->n_ranks), 256, 0, d->stream>>>(src, cap, d->d_objects_global);
    TGB_LAUNCH_CHECK(d);
    d->objects_gathered = TG_TRUE;
    return TG_TRUE;
}

/* every tile flag of this rank's current buffer = 1: the words were not written by K1's epilogue (uploaded buffer, BLOCKS view) */
extern "C" b32 tgbd_p2p_flag_all_tiles(struct tgb_device* d)
{
    const u32 n = tgbd_n_tiles(d);
    k_fill_words<<<(n + 255) / 256, 256, 0, d->stream>>>(tgbd_mat_tile_flags(d, d->d_mat), n, 1u);
    TGB_LAUNCH_CHECK(d);
    return TG_TRUE;
}

/* this rank's tile: merged words into d_vis_tile, the winners' material words into d_mat_tile (asynchronous on the stream) */
extern "C" b32 tgbd_p2p_merge_tile(struct tgb_device* d)
{
    if (!d->p2p_ready) { tgb_set_error("p2p_merge_tile: peer memory is not mapped"); return TG_FALSE; }
    const u64 tile_px = (u64)d->width * d->tile_rows;
    const u64 first = (u64)d->rank * tile_px; /* this rank's virtual rows: one contiguous block of every buffer */
    /* d_vis keeps this rank's LOCAL words (the peers read them, and a re-render without a clear must find them): the merged tile
     * goes to its own buffer, addressed with whole-frame pixel indices like d_vis */
    const dim3 tiles(tgbd_tiles_x(d), d->tile_rows / TGB_BAND_ROWS);
    k_merge_tile<<<tiles, 256, 0, d->stream>>>(tgbd__peer_table(d), tgbd__peer_tails(d, false), d->n_ranks, d->width, tgbd_tiles_x(d), d->rank * d->tile_rows,
                                               d->d_vis_tile - first, d->d_mat_tile);
    TGB_LAUNCH_CHECK(d);
    d->tile_merged = TG_TRUE;
    return TG_TRUE;
}

/*
 * The whole merged frame for a read-back or a pick. After the all-reduce that is d_vis itself; on the fused path d_vis holds this
 * rank's local words and only the tile was merged, so the frame is pulled from the peers into a separate buffer (every rank's words,
 * no tile flags: the slow, always-valid form). Valid while the peers have finished K1 of this frame and not cleared this buffer
 * again (they render in lock-step; call it between frames).
 */
extern "C" void* tgbd_visibility_for_read(struct tgb_device* d)
{
    if (d->n_ranks < 2 || d->vis_merged || !d->p2p_ready) return d->d_vis;
    if (cudaSetDevice(d->device) != cudaSuccess) return d->d_vis;
    const u64 px = (u64)d->width * d->tile_rows * d->n_ranks;
    if (!d->d_vis_full && cudaMalloc(&d->d_vis_full, px * sizeof(u64)) != cudaSuccess) { tgb_set_error("visibility_for_read: out of device memory"); return d->d_vis; }
    k_merge_linear<<<(u32)((px + 255) / 256), 256, 0, d->stream>>>(tgbd__peer_table(d), d->n_ranks, 0, px, d->d_vis_full);
    d->n_kernel_launches++;
    return d->d_vis_full;
}

/* host side of k_wait_peers' time-out: called where the stream has just been synchronised */
extern "C" void tgbd_p2p_check(struct tgb_device* d)
{
    if (!d->p2p_ready || !d->d_ipc_stage) return;
    u32 err = 0;
    if (cudaMemcpy(&err, tgbd__wait_error(d), sizeof(u32), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return; }
    if (err)
    {
        tgb_set_error("peer-memory merge: rank %u never published its frame (waited ~4 s for its K1): that process is gone or its frames are out of step", err - 1u);
        cudaMemset(tgbd__wait_error(d), 0, sizeof(u32));
    }
}

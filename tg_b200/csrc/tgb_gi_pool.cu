/*
 * tgb_gi_pool.cu -- K3b, the queued secondary rays through the 1-bit SVO with SEVERAL RAYS PER LANE.
 *
 *   SVO traversal      assets/shaders/raytracer/svo_functions.inc:1-329 (per-ray pieces: tgb_gi_walk.cuh)
 *   secondary rays     tgvk_raytracer.c:1405-1431, TODO.h:33-43 (queued by k_shade, tgb_shade.cu)
 *
 * Secondary rays are incoherent: at any moment some rays of a warp cross empty space (TREE: advance to the far border of
 * the terminal box + look the next one up), some walk a 32^3 leaf block voxel by voxel (DDA), a few wait for a decision
 * (HIT / MISS) or a new ray (IDLE). k_gi_trace_flat (tgb_shade.cu) gave every lane one ray and let the warp run the phase
 * most lanes waited for: 12.8 of 32 lanes did useful work per issued instruction (profiles/r01q_k3b_summary.txt), the
 * kernel being issue-bound that is what its duration follows.
 *
 * Here every lane owns K rays whose state (16 words, tgb_gi_walk.cuh) lives in shared memory -- one column per word,
 * [word][k][thread], so a lane's accesses never conflict -- and registers only hold the working set of the phase being run.
 * Each warp iteration counts the lanes that have AT LEAST ONE ray waiting for each phase, runs the phase with the larger
 * count, and every such lane picks one of its waiting rays: with p the fraction of rays in a phase, 1 - (1 - p)^K of the
 * lanes take part instead of p. Nothing is exchanged between lanes (no filing, no sorting: the shared-memory pools of
 * round 1 lost to exactly that); a phase loads ~10 words and stores ~5 per lane around a few hundred instructions.
 * The arithmetic per ray is untouched, so hit / miss decisions are those of k_gi_trace_flat and of the shader.
 */
#include "tgb_device.cuh"
#include "tgb_gi_walk.cuh"
#include "tgb_gi_fast.cuh"

#define TGB_POOL_THREADS 128
#define TGB_POOL_WORDS   16

/* shared-memory columns */
enum { W_DX = 0, W_DY, W_DZ, W_PX, W_PY, W_PZ, W_TDX, W_TDY, W_TDZ, W_TMX, W_TMY, W_TMZ, W_CELL, W_VOX, W_DATA, W_SLOT };

template <int K, int GRID, bool LIST>
__global__ void __launch_bounds__(TGB_POOL_THREADS) k_gi_trace_pool(const tgb_gi_frame fr, const float4* __restrict__ p_q0, const float4* __restrict__ p_q1,
                                                                    const float4* __restrict__ p_q2, u32* __restrict__ p_q_count, const u32* __restrict__ p_list, u32 count_word,
                                                                    float4* __restrict__ p_out, u32 service_slots, u32 dda_bias, u32 tree_reps, u32 dda_steps, u32 min_rays_per_slot, u32 n_sms, u32 tail_boost)
{
    if (fr.p_grid[TGB_TOP_GRID_CELLS] == 0) return; /* not tabulated: k_gi_trace runs */

    extern __shared__ u32 s_pool[];
    const u32 tid = threadIdx.x, lane = tid & 31u;
#define S(w, k) s_pool[((u32)(w) * K + (u32)(k)) * TGB_POOL_THREADS + tid]
#define SF(w, k) __uint_as_float(S(w, k))

    /* the whole queue (p_list == NULL: p_q_count[0] rays, fetch counter [1]) or the slots k_gi_trace_fast handed over (tgb_gi_fast.cu:
     * p_q_count[count_word] entries of p_list, fetch counter [count_word + 1]) */
    const u32 n_rays = p_q_count[LIST ? count_word : 0u];
    if (!LIST && blockIdx.x == 0 && tid == 0) { atomicAdd(&p_q_count[10], n_rays); atomicAdd(&p_q_count[14], n_rays); }
    if (LIST && blockIdx.x == 0 && tid == 0) atomicAdd(&p_q_count[14], n_rays); /* rays of the frame, summed over its bands; all of them traced exactly */
    /* CTAs beyond what the queue can feed (a band, a screen tile of a sharded frame) leave at once. min_rays_per_slot > 1 would keep
     * even fewer CTAs so that every ray slot sees several rays; measured on a 272-row tile (873 k rays): 0.41 ms with 1, 0.49 with 4, 0.69
     * with 8 -- with few rays the longest dependent chains set the duration and parallelism is all that helps (profiles/r02k) */
    if (!LIST && blockIdx.x >= n_sms && (u64)blockIdx.x * (TGB_POOL_THREADS * K) * min_rays_per_slot >= n_rays) return;
    /* The handed-over rays are few and long (every one walks out of the box): what bounds this pass is how many of them ONE warp holds.
     * A warp therefore takes its fair share of the list per round (all CTAs of the grid are resident) and comes back for more only when
     * those are decided; the whole queue (p_list == NULL) has no quota. */
    /* `rounds` (tuning, whole-queue mode): a warp takes 1 / rounds of its fair share at a time, so that the last grabs of a small batch (a screen tile) are small */
    const u32 n_warps = gridDim.x * (TGB_POOL_THREADS / 32u);
    const u32 quota = LIST ? (n_rays + n_warps - 1u) / n_warps + 1u : 0xFFFFFFFFu; /* (the same for the whole queue was measured worse on tile batches, profiles/r03f) */
    u32 round_left = quota;

    u32 kinds = 0; /* 4 bits per ray slot, all IDLE */
    bool exhausted = false;
    u32 n_visits = 0, n_steps = 0, n_advances = 0;

    for (;;)
    {
        u32 has_tree = 0, has_dda = 0, n_svc = 0;
#pragma unroll
        for (int k = 0; k < K; k++)
        {
            const u32 kd = (kinds >> (4 * k)) & 15u;
            has_tree |= (kd == TGB_RAY_TREE) ? 1u : 0u;
            has_dda |= (kd == TGB_RAY_DDA) ? 1u : 0u;
            n_svc += ((kd >= TGB_RAY_HIT) | ((kd == TGB_RAY_IDLE) & !exhausted & (!LIST || round_left != 0u))) ? 1u : 0u;
        }
        const u32 counts = __reduce_add_sync(0xFFFFFFFFu, has_tree | (has_dda << 8) | (n_svc << 16));
        const u32 n_tree = counts & 0xFFu, n_dda = (counts >> 8) & 0xFFu, n_service = counts >> 16;
        /* the last few rays of a warp are the long ones and nobody waits behind them: longer phases, less scheduling per step */
        const u32 boost = (n_tree + n_dda <= 6u) ? tail_boost : 1u;
        if (n_tree + n_dda == 0 && n_service == 0)
        {
            if (LIST && !exhausted && round_left == 0u) { round_left = quota; continue; } /* the round's rays are decided: next round */
            break; /* queue drained and every ray finished */
        }

        if (n_service >= service_slots || n_tree + n_dda == 0)
        {
            /* ---- service: unoccluded rays return their ambient term, voxel hits are decided, idle slots fetch rays ---- */
#pragma unroll
            for (int k = 0; k < K; k++)
            {
                u32 kd = (kinds >> (4 * k)) & 15u;
                if (kd == TGB_RAY_MISS)
                {
                    const u32 vox = S(W_VOX, k);
                    u32 flags = vox >> 16;
                    if (flags & TGB_RF_BORDER)
                    {
                        kd = tgb_gi_border_test(&fr, tgb_v3(SF(W_DX, k), SF(W_DY, k), SF(W_DZ, k)), tgb_v3(SF(W_PX, k), SF(W_PY, k), SF(W_PZ, k)), &flags);
                        S(W_VOX, k) = (vox & 0xFFFFu) | (flags << 16);
                    }
                }
                if (kd == TGB_RAY_MISS)
                {
                    /* ambient * 1 + lo; float addition commutes and the reductions do not stall the lane */
                    const u32 slot = S(W_SLOT, k);
                    const float4 q0 = __ldcg(&p_q0[slot]), q2 = __ldcs(&p_q2[slot]); /* q2 is read once (streaming); q0 was read at the refill: L2 only, keep L1 for the tree */
                    f32* p_pixel = reinterpret_cast<f32*>(&p_out[__float_as_uint(q0.w)]);
                    atomicAdd(p_pixel + 0, q2.x);
                    atomicAdd(p_pixel + 1, q2.y);
                    atomicAdd(p_pixel + 2, q2.z);
                    kd = TGB_RAY_IDLE;
                }
                else if (kd == TGB_RAY_HIT)
                {
                    const float4 q0 = __ldcg(&p_q0[S(W_SLOT, k)]);
                    const u32 vox = S(W_VOX, k);
                    v3 child_min; f32 child_size;
                    tgb_cell_box(&fr, S(W_CELL, k), &child_min, &child_size);
                    kd = tgb_gi_hit_test(&fr, tgb_v3(q0.x, q0.y, q0.z), tgb_v3(SF(W_DX, k), SF(W_DY, k), SF(W_DZ, k)), child_min,
                                         (i32)(vox & 31u), (i32)((vox >> 5) & 31u), (i32)((vox >> 10) & 31u));
                }
                if (!exhausted && (!LIST || round_left != 0u))
                {
                    u32 idle = __ballot_sync(0xFFFFFFFFu, kd == TGB_RAY_IDLE);
                    if (LIST && (u32)__popc(idle) > round_left) idle &= (1u << __fns(idle, 0u, (int)round_left + 1)) - 1u; /* the first round_left idle lanes */
                    if (idle)
                    {
                        const u32 n = (u32)__popc(idle);
                        u32 base = 0;
                        const u32 leader = (u32)(__ffs(idle) - 1);
                        if (lane == leader) base = atomicAdd(&p_q_count[LIST ? count_word + 1u : 1u], n);
                        base = __shfl_sync(0xFFFFFFFFu, base, (int)leader);
                        const u32 mine = base + (u32)__popc(idle & ((1u << lane) - 1u));
                        if (kd == TGB_RAY_IDLE && ((idle >> lane) & 1u) && mine < n_rays)
                        {
                            const u32 queue_slot = LIST ? __ldcs(&p_list[mine]) : mine;
                            const float4 q0 = __ldcg(&p_q0[queue_slot]), q1 = __ldcs(&p_q1[queue_slot]); /* q0 is read again when the ray is decided (L2), q1 never */
                            const v3 d = tgb_v3(q1.x, q1.y, q1.z);
                            v3 position, t_delta; u32 flags;
                            tgb_gi_ray_start(&fr, tgb_v3(q0.x, q0.y, q0.z), d, q1.w, &position, &t_delta, &flags);
                            S(W_DX, k) = __float_as_uint(d.x); S(W_DY, k) = __float_as_uint(d.y); S(W_DZ, k) = __float_as_uint(d.z);
                            S(W_PX, k) = __float_as_uint(position.x); S(W_PY, k) = __float_as_uint(position.y); S(W_PZ, k) = __float_as_uint(position.z);
                            S(W_TDX, k) = __float_as_uint(t_delta.x); S(W_TDY, k) = __float_as_uint(t_delta.y); S(W_TDZ, k) = __float_as_uint(t_delta.z);
                            S(W_CELL, k) = 0u;          /* iterations = 0 */
                            S(W_VOX, k) = flags << 16;  /* no advance pending: the first tree phase looks the entry cell up */
                            S(W_SLOT, k) = queue_slot;
                            kd = TGB_RAY_TREE;
                        }
                        exhausted = base + n >= n_rays;
                        if (LIST) round_left -= n;
                    }
                }
                kinds = (kinds & ~(15u << (4 * k))) | (kd << (4 * k));
            }
            continue;
        }

        if (n_dda + dda_bias > n_tree && n_dda > 0)
        {
            if (has_dda)
            {
                u32 k = 0;
#pragma unroll
                for (int j = K - 1; j >= 0; j--) if (((kinds >> (4 * j)) & 15u) == TGB_RAY_DDA) k = (u32)j;
                const u32 vox = S(W_VOX, k);
                u32 flags = vox >> 16;
                const v3 t_delta = tgb_v3(SF(W_TDX, k), SF(W_TDY, k), SF(W_TDZ, k));
                const u32* __restrict__ p_block = fr.p_voxels + (u64)S(W_DATA, k) * TG_SVO_BLOCK_WORDS;
                i32 x, y, z; v3 t_max;
                if (flags & TGB_RF_SETUP)
                {
                    /* :111-176: the lanes that entered a leaf since their last DDA phase set up together */
                    flags &= ~TGB_RF_SETUP;
                    v3 child_min; f32 child_size;
                    tgb_cell_box(&fr, S(W_CELL, k), &child_min, &child_size);
                    tgb_gi_dda_setup(tgb_v3(SF(W_DX, k), SF(W_DY, k), SF(W_DZ, k)), tgb_v3(SF(W_PX, k), SF(W_PY, k), SF(W_PZ, k)), child_min, child_size, &x, &y, &z, &t_max);
                }
                else
                {
                    x = (i32)(vox & 31u); y = (i32)((vox >> 5) & 31u); z = (i32)((vox >> 10) & 31u);
                    t_max = tgb_v3(SF(W_TMX, k), SF(W_TMY, k), SF(W_TMZ, k));
                }
                const u32 kd = tgb_gi_dda_phase(p_block, t_delta, flags >> TGB_RF_STEP_SHIFT, &t_max, &x, &y, &z, dda_steps * boost, &n_steps);
                S(W_TMX, k) = __float_as_uint(t_max.x); S(W_TMY, k) = __float_as_uint(t_max.y); S(W_TMZ, k) = __float_as_uint(t_max.z);
                S(W_VOX, k) = ((u32)x & 31u) | (((u32)y & 31u) << 5) | (((u32)z & 31u) << 10) | (flags << 16);
                kinds = (kinds & ~(15u << (4 * k))) | (kd << (4 * k));
            }
        }
        else if (has_tree)
        {
            u32 k = 0;
#pragma unroll
            for (int j = K - 1; j >= 0; j--) if (((kinds >> (4 * j)) & 15u) == TGB_RAY_TREE) k = (u32)j;
            const u32 vox = S(W_VOX, k);
            u32 flags = vox >> 16, cell = S(W_CELL, k), data = 0;
            v3 position = tgb_v3(SF(W_PX, k), SF(W_PY, k), SF(W_PZ, k));
            const u32 kd = tgb_gi_tree_phase_t<GRID>(&fr, tgb_v3(SF(W_DX, k), SF(W_DY, k), SF(W_DZ, k)), tgb_v3(SF(W_TDX, k), SF(W_TDY, k), SF(W_TDZ, k)),
                                             &position, &cell, &flags, &data, tree_reps * boost, &n_visits, &n_advances);
            S(W_PX, k) = __float_as_uint(position.x); S(W_PY, k) = __float_as_uint(position.y); S(W_PZ, k) = __float_as_uint(position.z);
            S(W_CELL, k) = cell;
            S(W_VOX, k) = (vox & 0xFFFFu) | (flags << 16);
            if (kd == TGB_RAY_DDA) S(W_DATA, k) = data;
            kinds = (kinds & ~(15u << (4 * k))) | (kd << (4 * k));
        }
    }
#undef S
#undef SF
    /* [2] look-ups, [3] DDA steps, [4] advances of this frame */
    n_visits = __reduce_add_sync(0xFFFFFFFFu, n_visits);
    n_steps = __reduce_add_sync(0xFFFFFFFFu, n_steps);
    n_advances = __reduce_add_sync(0xFFFFFFFFu, n_advances);
    if (lane == 0)
    {
        atomicAdd(reinterpret_cast<unsigned long long*>(p_q_count) + 1, (unsigned long long)n_visits);
        atomicAdd(reinterpret_cast<unsigned long long*>(p_q_count) + 2, (unsigned long long)n_steps);
        atomicAdd(reinterpret_cast<unsigned long long*>(p_q_count) + 3, (unsigned long long)n_advances);
    }
}

/*
 * Launch for one band of rays (called by tgbd__shade_launch, tgb_shade.cu, after k_shade queued them). Tuning knobs are read
 * from the environment once (benchmark sweeps only): rays per lane, CTAs per SM, phase budgets, service threshold.
 */
template <int K, int GRID, bool LIST>
static b32 tgbd__gi_pool_launch(struct tgb_device* d, const tgb_gi_frame& fr, const u32* p_list, u32 count_word, u32 rounds, u32 ctas_per_sm, u32 service_slots, u32 dda_bias, u32 tree_reps, u32 dda_steps, u32 min_rays_per_slot)
{
    const size_t smem = (size_t)TGB_POOL_WORDS * K * TGB_POOL_THREADS * sizeof(u32);
    static bool attr_set = false;
    if (!attr_set)
    {
        TGB_CUDA(cudaFuncSetAttribute(k_gi_trace_pool<K, GRID, LIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    /* resident CTAs per SM: shared memory (227 KB, 1 KB reserved per CTA) and the 2048-thread limit */
    u32 fit = (u32)((227u * 1024u) / (smem + 1024u));
    if (fit > 2048u / TGB_POOL_THREADS) fit = 2048u / TGB_POOL_THREADS;
    if (ctas_per_sm == 0 || ctas_per_sm > fit) ctas_per_sm = fit < 8u ? fit : 8u;
    k_gi_trace_pool<K, GRID, LIST><<<d->n_sms * ctas_per_sm, TGB_POOL_THREADS, smem, d->stream>>>(fr, d->d_gi_q0, d->d_gi_q1, d->d_gi_q2, d->d_gi_count, p_list, count_word, d->d_radiance,
                                                                                     service_slots, dda_bias, tree_reps, dda_steps, min_rays_per_slot, d->n_sms, rounds);
    TGB_LAUNCH_CHECK(d);
    return TG_TRUE;
}

/* ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) of one leaf block into shared memory, completion on an mbarrier ---------------------------- */
__device__ __forceinline__ void tgb_mbar_init(u64* p_bar, u32 n_arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((u32)__cvta_generic_to_shared(p_bar)), "r"(n_arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
/* one thread: expect `bytes`, then start the copy (16-byte aligned on both sides, a multiple of 16 bytes) */
__device__ __forceinline__ void tgb_bulk_copy_g2s(void* p_shared, const void* p_global, u32 bytes, u64* p_bar)
{
    const u32 bar = (u32)__cvta_generic_to_shared(p_bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* the block's previous contents were read through the generic proxy */
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((u32)__cvta_generic_to_shared(p_shared)), "l"(p_global), "r"(bytes), "r"(bar) : "memory");
}
/* every consumer: wait for the phase with this parity; bounded (a copy that never lands reports false instead of hanging the GPU) */
__device__ __forceinline__ bool tgb_mbar_wait(u64* p_bar, u32 parity)
{
    const u32 bar = (u32)__cvta_generic_to_shared(p_bar);
    for (u32 spin = 0; spin < (1u << 22); spin++)
    {
        u32 done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}

/*
 * The rays the certified fast walk hands over (tgb_gi_fast.cu) are few -- under a thousandth of the queue -- and long: every one
 * leaves the box, a hundred voxel steps and a dozen look-ups on average, four times that at worst, one after the other. With so little
 * work the pool's phase machinery (votes, state in shared memory, a ray waiting for the phase its neighbours need) only stretches those
 * dependent chains, and so does every cache miss on the way: what this pass takes is the LONGEST ray's chain (0.16 - 0.20 ms whether
 * 1 k or 35 k rays were handed over, profiles/r04a - r04f). So here ONE WARP runs ONE ray: all lanes execute the same per-ray
 * functions (tgb_gi_walk.cuh) on the same ray -- no divergence, the issue cost of one lane -- and when the ray enters a leaf block the 32
 * lanes have the block's 4 KB in shared memory before the leaf DDA takes its first step, which then runs at shared-memory latency
 * instead of one dependent, mostly missing load per voxel row. The block is one contiguous, aligned 4 KB: ONE TMA bulk copy
 * (cp.async.bulk, completion counted on an mbarrier) issued by one lane moves it (TMA = true, default); the predecessor -- 32 coalesced loads and
 * stores per lane -- stays selectable (TGB_GI_LIST_TMA=0).
 */
#define TGB_LIST_THREADS 32
template <bool TMA>
__global__ void __launch_bounds__(TGB_LIST_THREADS, 32) k_gi_trace_list(const tgb_gi_frame fr, const float4* __restrict__ p_q0, const float4* __restrict__ p_q1,
                                                                        const float4* __restrict__ p_q2, u32* __restrict__ p_q_count, const u32* __restrict__ p_list, u32 count_word,
                                                                        float4* __restrict__ p_out, u32 tree_reps, u32 dda_steps)
{
    if (fr.p_grid[TGB_TOP_GRID_CELLS] == 0) return; /* not tabulated: k_gi_trace runs */
    __shared__ __align__(128) u32 s_block[TG_SVO_BLOCK_WORDS];
    __shared__ __align__(8) u64 s_bar;
    const u32 lane = threadIdx.x;
    const u32 n_rays = p_q_count[count_word];
    u32 n_visits = 0, n_steps = 0, n_advances = 0, n_traced = 0;
    u32 parity = 0;
    if (TMA)
    {
        if (lane == 0) tgb_mbar_init(&s_bar, 1u);
        __syncwarp();
    }
    for (;;)
    {
        u32 mine = 0;
        if (lane == 0) mine = atomicAdd(&p_q_count[count_word + 1u], 1u);
        mine = __shfl_sync(0xFFFFFFFFu, mine, 0);
        if (mine >= n_rays) break;
        const u32 slot = __ldcs(&p_list[mine]) & 0x7FFFFFFFu;
        const float4 q0 = __ldcg(&p_q0[slot]), q1 = __ldcs(&p_q1[slot]);
        const v3 origin = tgb_v3(q0.x, q0.y, q0.z), d = tgb_v3(q1.x, q1.y, q1.z);
        v3 position, t_delta, t_max = tgb_v3(0.0f, 0.0f, 0.0f);
        u32 flags, cell = 0, data = 0, kind = TGB_RAY_TREE;
        i32 x = 0, y = 0, z = 0;
        tgb_gi_ray_start(&fr, origin, d, q1.w, &position, &t_delta, &flags);
        bool occluded = false;
        n_traced++;
        for (;;) /* warp-uniform: every lane holds the same state */
        {
            if (kind == TGB_RAY_TREE) kind = tgb_gi_tree_phase(&fr, d, t_delta, &position, &cell, &flags, &data, tree_reps, &n_visits, &n_advances);
            else if (kind == TGB_RAY_DDA)
            {
                if (flags & TGB_RF_SETUP)
                {
                    flags &= ~TGB_RF_SETUP;
                    v3 child_min; f32 child_size;
                    tgb_cell_box(&fr, cell, &child_min, &child_size);
                    tgb_gi_dda_setup(d, position, child_min, child_size, &x, &y, &z, &t_max);
                    const u32* __restrict__ p_block = fr.p_voxels + (u64)data * TG_SVO_BLOCK_WORDS;
                    __syncwarp(); /* every lane is done with the previous block */
                    if (TMA)
                    {
                        if (lane == 0) tgb_bulk_copy_g2s(s_block, p_block, TG_SVO_BLOCK_WORDS * (u32)sizeof(u32), &s_bar);
                        if (!tgb_mbar_wait(&s_bar, parity)) { if (lane == 0) atomicExch(&p_q_count[20], 1u); return; } /* [20]: the copy never landed (reported by the host) */
                        parity ^= 1u;
                    }
                    else
                    {
#pragma unroll
                        for (u32 i = 0; i < 32u; i++) s_block[32u * i + lane] = __ldg(&p_block[32u * i + lane]);
                        __syncwarp();
                    }
                }
                kind = tgb_gi_dda_phase_uniform(s_block, t_delta, flags >> TGB_RF_STEP_SHIFT, &t_max, &x, &y, &z, dda_steps, &n_steps);
                if (kind == TGB_RAY_DDA || kind == TGB_RAY_HIT) { x &= 31; y &= 31; z &= 31; } /* what the pool keeps between phases */
            }
            else if (kind == TGB_RAY_HIT)
            {
                v3 child_min; f32 child_size;
                tgb_cell_box(&fr, cell, &child_min, &child_size);
                kind = tgb_gi_hit_test(&fr, origin, d, child_min, x, y, z);
                if (kind == TGB_RAY_IDLE) { occluded = true; break; }
            }
            else /* MISS */
            {
                if (flags & TGB_RF_BORDER) kind = tgb_gi_border_test(&fr, d, position, &flags);
                if (kind == TGB_RAY_MISS) break;
            }
        }
        if (!occluded && lane == 0)
        {
            const float4 q2 = __ldcs(&p_q2[slot]);
            f32* p_pixel = reinterpret_cast<f32*>(&p_out[__float_as_uint(q0.w)]);
            atomicAdd(p_pixel + 0, q2.x);
            atomicAdd(p_pixel + 1, q2.y);
            atomicAdd(p_pixel + 2, q2.z);
        }
    }
    if (lane == 0 && n_traced)
    {
        atomicAdd(&p_q_count[14], n_traced); /* rays that needed the shader's own arithmetic */
        atomicAdd(reinterpret_cast<unsigned long long*>(p_q_count) + 1, (unsigned long long)n_visits);
        atomicAdd(reinterpret_cast<unsigned long long*>(p_q_count) + 2, (unsigned long long)n_steps);
        atomicAdd(reinterpret_cast<unsigned long long*>(p_q_count) + 3, (unsigned long long)n_advances);
    }
}

extern "C" b32 tgbd_gi_pool_trace_list(struct tgb_device* d, f32 far_plane, const u32* p_list, u32 count_word)
{
    const int rays_per_lane = tgbd_env_int("TGB_GI_RAYS_PER_LANE", p_list ? 1 : 3); /* measured: 1.40 ms for the stage with 3, 1.43 with 2, 1.52 with 4 (L1 shrinks with the pool), 1.51 for k_gi_trace_flat */
    const u32 ctas_per_sm = (u32)tgbd_env_int("TGB_GI_POOL_CTAS_PER_SM", 0);
    const u32 dda_steps = (u32)max(1, tgbd_env_int("TGB_GI_POOL_DDA_STEPS", 16));
    const u32 tree_reps = (u32)max(1, tgbd_env_int("TGB_GI_POOL_TREE_REPS", 4));
    const u32 dda_bias = (u32)tgbd_env_int("TGB_GI_POOL_DDA_BIAS", 0);
    const u32 min_rays_per_slot = (u32)max(1, tgbd_env_int("TGB_GI_POOL_MIN_RAYS_PER_SLOT", 1));
    const int service_env = tgbd_env_int("TGB_GI_POOL_SERVICE_SLOTS", 0);
    const u32 rounds = (u32)max(1, tgbd_env_int("TGB_GI_POOL_TAIL_BOOST", 1)); /* phase budgets x this once at most 6 lanes of a warp still hold a working ray */
    tgb_gi_frame fr;
    tgb_gi_frame_init(&fr, d->svo.bmin, d->svo.bmax, far_plane, d->svo.d_top_grid, d->svo.d_voxels);
    if (p_list && tgbd_env_int("TGB_GI_LIST_KERNEL", 1))
    {
        /* the handed-over rays: k_gi_trace_list (TGB_GI_LIST_KERNEL=0: the pool kernel in list mode, the measured predecessor) */
        const u32 list_ctas = (u32)max(1, min(32, tgbd_env_int("TGB_GI_LIST_CTAS_PER_SM", 32)));
        const u32 tree_reps = (u32)max(1, tgbd_env_int("TGB_GI_LIST_TREE_REPS", 64)), dda_steps = (u32)max(1, tgbd_env_int("TGB_GI_LIST_DDA_STEPS", 1024));
        if (tgbd_env_int("TGB_GI_LIST_TMA", 1))
            k_gi_trace_list<true><<<d->n_sms * list_ctas, TGB_LIST_THREADS, 0, d->stream>>>(fr, d->d_gi_q0, d->d_gi_q1, d->d_gi_q2, d->d_gi_count, p_list, count_word, d->d_radiance, tree_reps, dda_steps);
        else
            k_gi_trace_list<false><<<d->n_sms * list_ctas, TGB_LIST_THREADS, 0, d->stream>>>(fr, d->d_gi_q0, d->d_gi_q1, d->d_gi_q2, d->d_gi_count, p_list, count_word, d->d_radiance, tree_reps, dda_steps);
        TGB_LAUNCH_CHECK(d);
        return TG_TRUE;
    }
    /* TGB_GI_POOL_GRID16=1 reads the 16-bit form of the table (half the footprint in what L1 the pool leaves): measured 1.381 vs 1.383 ms for the
     * stage -- the look-ups are concentrated on few cells and hit L1 either way; what misses is the voxel rows. Off; kept as the measured record. */
    const int grid16 = tgbd_env_int("TGB_GI_POOL_GRID16", 0);
    fr.p_grid16 = (const unsigned short*)(d->svo.d_top_grid + TGB_TOP_GRID_CELLS + 1);
#define TGB_POOL_ARGS(KK) d, fr, p_list, count_word, rounds, ctas_per_sm, service_env > 0 ? (u32)service_env : (KK == 3 ? 64u : 16u * KK), dda_bias, tree_reps, dda_steps, min_rays_per_slot
#define TGB_POOL_CASE(KK) case KK: return p_list ? tgbd__gi_pool_launch<KK, 0, true>(TGB_POOL_ARGS(KK)) \
                                                 : (grid16 ? tgbd__gi_pool_launch<KK, 1, false>(TGB_POOL_ARGS(KK)) : tgbd__gi_pool_launch<KK, 0, false>(TGB_POOL_ARGS(KK)))
    switch (rays_per_lane)
    {
    TGB_POOL_CASE(1);
    TGB_POOL_CASE(2);
    TGB_POOL_CASE(3);
    TGB_POOL_CASE(4);
    default: return p_list ? tgbd__gi_pool_launch<1, 0, true>(TGB_POOL_ARGS(1)) : tgbd__gi_pool_launch<3, 0, false>(TGB_POOL_ARGS(3));
    }
#undef TGB_POOL_ARGS
#undef TGB_POOL_CASE
}

/* the whole queue of the band */
extern "C" b32 tgbd_gi_pool_trace(struct tgb_device* d, f32 far_plane)
{
    return tgbd_gi_pool_trace_list(d, far_plane, NULL, 0u);
}

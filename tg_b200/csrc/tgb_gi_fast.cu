/*
 * tgb_gi_fast.cu -- K3b, first pass: the queued secondary rays decided by the certified fast walk (tgb_gi_fast.cuh).
 *
 *   SVO traversal      assets/shaders/raytracer/svo_functions.inc:1-329
 *   secondary rays     tgvk_raytracer.c:1405-1431, TODO.h:33-43 (queued by k_shade, tgb_shade.cu)
 *
 * One ray per lane, everything in registers: the walk's state is the ray (origin, direction, reciprocals), a ray parameter and a
 * cell; no division, no tie rules, no accumulated position -- a fifth of the instructions of the exact walk per visited cell.
 * Decisions it can certify (tgb_gi_fast.cuh: every ray displaced sideways by less than DELTA decides the same) are final:
 * occluded rays keep the radiance k_shade wrote, unoccluded rays add their ambient term. The others -- a few per cent -- are
 * appended to a list of queue slots, and k_gi_trace_pool (tgb_gi_pool.cu: the shader's own arithmetic) traces exactly those right
 * after this kernel. The frame is therefore the one the exact kernel alone produces (tests: radiance bit-identical between
 * TGB_GI_KERNEL=2 and 3 on every pixel).
 *
 * Scheduling as in k_gi_trace_flat: lanes are TREE (boxes of the flattened tree) or DDA (voxels of a leaf block), each warp
 * iteration runs the phase most lanes wait for; finished lanes wait for a service phase (ambient term / hand-over / next ray).
 */
#include "tgb_device.cuh"
#include "tgb_gi_fast.cuh"

#define TGB_FAST_THREADS 128

__global__ void __launch_bounds__(TGB_FAST_THREADS) k_gi_trace_fast(const tgb_gi_frame fr, const float4* __restrict__ p_q0, const float4* __restrict__ p_q1,
                                                                    const float4* __restrict__ p_q2, u32* __restrict__ p_q_count, u32* __restrict__ p_exact_list,
                                                                    float4* __restrict__ p_out, u32 service_lanes, u32 tree_reps, u32 dda_steps, u32 dda_bias)
{
    if (fr.p_grid[TGB_TOP_GRID_CELLS] == 0) return; /* not tabulated: k_gi_trace runs */

    const u32 lane = threadIdx.x & 31u;
    const u32 n_rays = p_q_count[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&p_q_count[10], n_rays); /* rays of the frame, summed over its bands */

    tgb_fast_ray r;
    r.o = r.d = r.inv = r.p = r.t_max = tgb_v3(0.0f, 0.0f, 0.0f);
    r.w = r.t_cur = r.m_cur = r.w_leaf = 0.0f;
    r.cell = r.vox = r.data = r.entry_axis = r.uncertain = r.n_boxes = 0;
    u32 kind = TGB_FAST_IDLE, slot = 0, pixel = 0;
    bool exhausted = false;
    u32 n_visits = 0, n_steps = 0, n_exact = 0;

    for (;;)
    {
        const u32 counts = __reduce_add_sync(0xFFFFFFFFu, 1u << (6u * (kind > TGB_FAST_OCCLUDED ? TGB_FAST_OCCLUDED : kind)));
        const u32 n_idle = counts & 63u, n_tree = (counts >> 6) & 63u, n_dda = (counts >> 12) & 63u, n_done = (counts >> 18) & 63u;
        const u32 n_service = n_done + (exhausted ? 0u : n_idle);
        const u32 n_working = n_tree + n_dda;
        if (n_working == 0 && n_service == 0) break; /* queue drained and every ray finished */

        if (n_service >= service_lanes || n_working == 0)
        {
            /* ---- service: certain unoccluded rays return their ambient term, idle lanes fetch rays, uncertain rays are handed over ---- */
            if (kind == TGB_FAST_UNOCCLUDED)
            {
                if (r.uncertain) kind = TGB_FAST_EXACT;
                else
                {
                    const float4 q2 = __ldcs(&p_q2[slot]);
                    f32* p_pixel = reinterpret_cast<f32*>(&p_out[pixel]);
                    atomicAdd(p_pixel + 0, q2.x);
                    atomicAdd(p_pixel + 1, q2.y);
                    atomicAdd(p_pixel + 2, q2.z);
                    kind = TGB_FAST_IDLE;
                }
            }
            else if (kind == TGB_FAST_OCCLUDED) kind = TGB_FAST_IDLE;
            if (!exhausted)
            {
                const u32 idle = __ballot_sync(0xFFFFFFFFu, kind == TGB_FAST_IDLE);
                if (idle)
                {
                    const u32 n = (u32)__popc(idle);
                    u32 base = 0;
                    const u32 leader = (u32)(__ffs(idle) - 1);
                    if (lane == leader) base = atomicAdd(&p_q_count[1], n);
                    base = __shfl_sync(0xFFFFFFFFu, base, (int)leader);
                    const u32 mine = base + (u32)__popc(idle & ((1u << lane) - 1u));
                    if (kind == TGB_FAST_IDLE && mine < n_rays)
                    {
                        slot = mine;
                        const float4 q0 = __ldcs(&p_q0[mine]), q1 = __ldcs(&p_q1[mine]);
                        pixel = __float_as_uint(q0.w);
                        kind = tgb_fast_start(&fr, tgb_v3(q0.x, q0.y, q0.z), tgb_v3(q1.x, q1.y, q1.z), q1.w, TGB_FAST_DELTA, &r);
                    }
                    exhausted = base + n >= n_rays;
                }
            }
            {
                /* hand-over: the slot goes to the list k_gi_trace_pool reads (warp-aggregated append) */
                const u32 handed = __ballot_sync(0xFFFFFFFFu, kind == TGB_FAST_EXACT);
                if (handed)
                {
                    u32 base = 0;
                    const u32 leader = (u32)(__ffs(handed) - 1);
                    if (lane == leader) base = atomicAdd(&p_q_count[12], (u32)__popc(handed));
                    base = __shfl_sync(0xFFFFFFFFu, base, (int)leader);
                    if (kind == TGB_FAST_EXACT)
                    {
                        p_exact_list[base + (u32)__popc(handed & ((1u << lane) - 1u))] = slot;
                        n_exact++;
                        kind = TGB_FAST_IDLE;
                    }
                }
            }
            continue;
        }

        if (n_dda + dda_bias > n_tree && n_dda > 0)
        {
            if (kind == TGB_FAST_DDA) kind = tgb_fast_dda_phase(&fr, &r, dda_steps, &n_steps);
        }
        else if (kind == TGB_FAST_TREE) kind = tgb_fast_tree_phase(&fr, &r, tree_reps, &n_visits);
    }
    /* [2] boxes, [3] DDA steps of this frame (the exact kernel adds its own); [14] rays handed over */
    n_visits = __reduce_add_sync(0xFFFFFFFFu, n_visits);
    n_steps = __reduce_add_sync(0xFFFFFFFFu, n_steps);
    n_exact = __reduce_add_sync(0xFFFFFFFFu, n_exact);
    if (lane == 0)
    {
        atomicAdd(reinterpret_cast<unsigned long long*>(p_q_count) + 1, (unsigned long long)n_visits);
        atomicAdd(reinterpret_cast<unsigned long long*>(p_q_count) + 2, (unsigned long long)n_steps);
        atomicAdd(&p_q_count[14], n_exact);
    }
}

/*
 * One band of rays (called by tgbd__shade_launch, tgb_shade.cu, after k_shade queued them): the fast walk over the whole queue, then
 * the exact kernel over the slots it handed over (p_q_count[12] of them, read on the device).
 */
extern "C" b32 tgbd_gi_fast_trace(struct tgb_device* d, f32 far_plane)
{
    const u32 ctas_per_sm = (u32)max(1, min(16, tgbd_env_int("TGB_GI_FAST_CTAS_PER_SM", 8)));
    const u32 service_lanes = (u32)max(1, tgbd_env_int("TGB_GI_FAST_SERVICE_LANES", 12));
    const u32 tree_reps = (u32)max(1, tgbd_env_int("TGB_GI_FAST_TREE_REPS", 4));
    const u32 dda_steps = (u32)max(1, tgbd_env_int("TGB_GI_FAST_DDA_STEPS", 16));
    const u32 dda_bias = (u32)tgbd_env_int("TGB_GI_FAST_DDA_BIAS", 0);
    tgb_gi_frame fr;
    tgb_gi_frame_init(&fr, d->svo.bmin, d->svo.bmax, far_plane, d->svo.d_top_grid, d->svo.d_voxels);
    k_set_words<<<1, 32, 0, d->stream>>>(d->d_gi_count + 12, 2, 0u); /* handed over / fetched by the exact kernel */
    TGB_LAUNCH_CHECK(d);
    k_gi_trace_fast<<<d->n_sms * ctas_per_sm, TGB_FAST_THREADS, 0, d->stream>>>(fr, d->d_gi_q0, d->d_gi_q1, d->d_gi_q2, d->d_gi_count, d->d_gi_exact, d->d_radiance,
                                                                                 service_lanes, tree_reps, dda_steps, dda_bias);
    TGB_LAUNCH_CHECK(d);
    return tgbd_gi_pool_trace_list(d, far_plane, d->d_gi_exact, 12u);
}

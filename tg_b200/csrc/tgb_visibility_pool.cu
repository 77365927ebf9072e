/*
 * tgb_visibility_pool.cu -- K1 with SEVERAL PIXELS PER LANE.
 *
 *   fragment     assets/shaders/raytracer/visibility.frag:22-208 (per-ray pieces: tgb_k1_walk.cuh)
 *   coverage     visibility.vert:19-26 + cluster_functions.inc:1-38, instanced draw tgvk_raytracer.c:1275-1317 -- replaced by the
 *                object cull + per-tile object lists of tgb_visibility.cu (k_cull_objects / k_sort_frames), shared with k_visibility
 *
 * k_visibility (tgb_visibility.cu) gives every lane one pixel. A ray needs 1-3 rounds of { find the next cluster it can enter,
 * march through its 8^3 voxels }, two thirds of the rays are done after the first round, and the lanes that are done wait for
 * the stragglers of their warp: the march, two thirds of the kernel's instructions, ran with 6.6-11.4 of 32 lanes
 * (profiles/r01p_k1_regions.txt).
 *
 * Here a lane owns K pixels (one 8 x 4K pixel block per warp, a 16 x 16K tile per CTA). The state of a pixel's walk (16 words:
 * tgb_k1_walk packed, best word, t_skip, pending candidate) lives in shared memory, one column per word ([word][k][thread]:
 * conflict-free), registers only hold the working set of the piece being run. Every pixel is in one of the states
 *   NEXT  needs its next object from the tile's front-to-back list (rectangle / depth-bound tests, then tgb_k1_setup)
 *   ENUM  walks an object and needs its next candidate cluster (tgb_k1_next_candidate)
 *   CAND  holds a candidate cluster to march through (tgb_cluster_march), then ENUM again
 * and each warp iteration runs the piece for which most lanes have a pixel waiting (one REDUX), every such lane taking one of
 * its waiting pixels: stragglers of one pixel overlap with the first rounds of the lane's other pixels, and pixels do not wait
 * for each other at object boundaries. Per pixel the sequence of clusters and the arithmetic are those of k_visibility (the
 * same functions), the per-pixel minimum is order-free, so the buffer is bit-identical.
 */
#include "tgb_device.cuh"
#include "tgb_k1_walk.cuh"

#define TGB_K1P_THREADS 256
#define TGB_K1P_WORDS   16
#define TGB_K1P_TILE_W  16
#define TGB_FULL_MASK   0xFFFFFFFFu

enum { P_DX = 0, P_DY, P_DZ, P_TDX, P_TDY, P_TDZ, P_TIN, P_TOUT, P_ENTER, P_BEST_LO, P_BEST_HI, P_TSKIP, P_I0, P_I1, P_I2, P_I3 };
/* pixel states (4 bits each in the lane's `kinds` register) */
enum { ST_NEXT = 0, ST_ENUM = 1, ST_CAND = 2, ST_WINDOW_DONE = 3, ST_FINISHED = 4 };

/* P_I0 = s | n_slices << 16, P_I1 = cu | cu0 << 16, P_I2 = cu1 | cv << 16, P_I3 = cv1 | flags << 16 (all signed 16 bit);
 * flags = axis (2) | negative (1) | exotic (1) | list cursor j (9): the object being walked is list entry j - 1 */
__device__ __forceinline__ u32 tgb_pack16(i32 lo, i32 hi) { return ((u32)lo & 0xFFFFu) | ((u32)hi << 16); }
__device__ __forceinline__ i32 tgb_lo16(u32 w) { return (i32)(short)(w & 0xFFFFu); }
__device__ __forceinline__ i32 tgb_hi16(u32 w) { return (i32)(short)(w >> 16); }

template <int K, int MIN_CTAS>
__global__ void __launch_bounds__(TGB_K1P_THREADS, MIN_CTAS) k_visibility_pool(const tgb_object_frame* __restrict__ p_frames, const u32* __restrict__ p_count,
                                                                               tg_camera_rays cam, u32 w, u32 h,
                                                                               const u32* __restrict__ p_cluster_pointers, const u32* __restrict__ p_masks,
                                                                               u32 global_pointer_base, u64* __restrict__ p_vis, u32 n_ranks, u32 tile_rows)
{
    __shared__ u32 s_list[TGB_K1P_THREADS];
    __shared__ u32 s_warp_count[TGB_K1P_THREADS / 32];
    extern __shared__ u32 s_pool[];
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
#define S(word, k) s_pool[((u32)(word) * K + (u32)(k)) * TGB_K1P_THREADS + tid]
#define SF(word, k) __uint_as_float(S(word, k))

    const u32 n_visible = p_count[0];
    if (n_visible == 0) return;
    if (p_count[2] != 0) return; /* an object with more than 32767 clusters along an axis: the packed iterator cannot hold it, k_visibility runs */
    const bool sorted = p_count[1] != 0;

    constexpr u32 TILE_H = 16u * K;
    const i32 tx0 = (i32)(blockIdx.x * TGB_K1P_TILE_W), ty0 = (i32)(blockIdx.y * TILE_H);
    const i32 tx1 = (i32)min(blockIdx.x * TGB_K1P_TILE_W + TGB_K1P_TILE_W - 1u, w - 1u), ty1 = (i32)min(blockIdx.y * TILE_H + TILE_H - 1u, h - 1u);
    const u32 px = blockIdx.x * TGB_K1P_TILE_W + (warp & 1u) * 8u + (lane & 7u);
    const u32 py0 = blockIdx.y * TILE_H + (warp >> 1) * (4u * K) + (lane >> 3); /* pixel k of the lane: (px, py0 + 4 k) */

    u32 kinds = 0;
#pragma unroll
    for (int k = 0; k < K; k++)
    {
        S(P_BEST_LO, k) = 0xFFFFFFFFu; S(P_BEST_HI, k) = 0xFFFFFFFFu;
        S(P_TSKIP, k) = __float_as_uint(TG_F32_MAX);
        const bool in_screen = px < w && py0 + 4u * k < h;
        kinds |= (in_screen ? (u32)ST_NEXT : (u32)ST_FINISHED) << (4 * k);
    }

    for (u32 base = 0; base < n_visible; base += TGB_K1P_THREADS)
    {
        /* order-preserving compaction of this window's objects that overlap the tile */
        const u32 i = base + tid;
        bool overlaps = false;
        if (i < n_visible)
        {
            const tgb_object_frame& f = p_frames[i];
            overlaps = !(f.x1 < tx0 || f.x0 > tx1 || f.y1 < ty0 || f.y0 > ty1);
        }
        const u32 ballot = __ballot_sync(TGB_FULL_MASK, overlaps);
        if (lane == 0) s_warp_count[warp] = (u32)__popc(ballot);
        __syncthreads();
        u32 offset = 0, n_listed = 0;
#pragma unroll
        for (u32 k = 0; k < TGB_K1P_THREADS / 32; k++)
        {
            const u32 c = s_warp_count[k];
            offset += k < warp ? c : 0u;
            n_listed += c;
        }
        if (overlaps) s_list[offset + (u32)__popc(ballot & ((1u << lane) - 1u))] = i;
        __syncthreads();

        /* every pixel that is not finished starts this window's list */
#pragma unroll
        for (int k = 0; k < K; k++)
        {
            if (((kinds >> (4 * k)) & 15u) != ST_FINISHED)
            {
                kinds = (kinds & ~(15u << (4 * k))) | ((u32)ST_NEXT << (4 * k));
                S(P_I3, k) = 0u; /* cursor 0 */
            }
        }

        if (n_listed != 0)
        for (;;)
        {
            u32 has_next = 0, has_enum = 0, has_cand = 0;
#pragma unroll
            for (int k = 0; k < K; k++)
            {
                const u32 st = (kinds >> (4 * k)) & 15u;
                has_next |= st == ST_NEXT ? 1u : 0u;
                has_enum |= st == ST_ENUM ? 1u : 0u;
                has_cand |= st == ST_CAND ? 1u : 0u;
            }
            const u32 counts = __reduce_add_sync(TGB_FULL_MASK, has_next | (has_enum << 8) | (has_cand << 16));
            const u32 n_next = counts & 0xFFu, n_enum = (counts >> 8) & 0xFFu, n_cand = counts >> 16;
            if (counts == 0) break;

            if (n_cand >= n_enum && n_cand >= n_next)
            {
                /* ---- march: second half of the fragment for the candidate a pixel holds ---- */
                if (has_cand)
                {
                    u32 k = 0;
#pragma unroll
                    for (int j = K - 1; j >= 0; j--) if (((kinds >> (4 * j)) & 15u) == ST_CAND) k = (u32)j;
                    const u32 i0 = S(P_I0, k), i1 = S(P_I1, k), i2 = S(P_I2, k), i3 = S(P_I3, k);
                    const u32 flags = i3 >> 16, axis = flags & 3u;
                    const i32 s = tgb_lo16(i0), cu = tgb_lo16(i1), cv = tgb_hi16(i2);
                    const u32 cx = (u32)TGB_K1_K(axis, s, cv, cu), cy = (u32)TGB_K1_K(axis, cu, s, cv), cz = (u32)TGB_K1_K(axis, cv, cu, s);
                    tgb_ray_in_object r;
                    tgb_ray_in_object_restore(&r, tgb_v3(SF(P_DX, k), SF(P_DY, k), SF(P_DZ, k)), SF(P_TDX, k), SF(P_TDY, k), SF(P_TDZ, k), (flags & 8u) != 0);
                    u64 best = ((u64)S(P_BEST_HI, k) << 32) | (u64)S(P_BEST_LO, k);
                    f32 t_skip = SF(P_TSKIP, k);
                    const tgb_object_frame& f = p_frames[s_list[(flags >> 4) - 1u]];
                    tgb_cluster_march(f, r, cx, cy, cz, SF(P_ENTER, k), cam.far_plane, p_cluster_pointers, p_masks, global_pointer_base, best, t_skip);
                    S(P_BEST_LO, k) = (u32)best; S(P_BEST_HI, k) = (u32)(best >> 32);
                    S(P_TSKIP, k) = __float_as_uint(t_skip);
                    kinds = (kinds & ~(15u << (4 * k))) | ((u32)ST_ENUM << (4 * k));
                }
            }
            else if (n_enum >= n_next)
            {
                /* ---- enumerate: advance a pixel's walk to its next candidate cluster ---- */
                if (has_enum)
                {
                    u32 k = 0;
#pragma unroll
                    for (int j = K - 1; j >= 0; j--) if (((kinds >> (4 * j)) & 15u) == ST_ENUM) k = (u32)j;
                    const u32 i0 = S(P_I0, k), i1 = S(P_I1, k), i2 = S(P_I2, k), i3 = S(P_I3, k);
                    const u32 flags = i3 >> 16;
                    tgb_k1_walk wk;
                    wk.d = tgb_v3(SF(P_DX, k), SF(P_DY, k), SF(P_DZ, k));
                    wk.t_delta_x = SF(P_TDX, k); wk.t_delta_y = SF(P_TDY, k); wk.t_delta_z = SF(P_TDZ, k);
                    wk.t_in = SF(P_TIN, k); wk.t_out = SF(P_TOUT, k);
                    wk.s = tgb_lo16(i0); wk.n_slices = tgb_hi16(i0);
                    wk.cu = tgb_lo16(i1); wk.cu0 = tgb_hi16(i1);
                    wk.cu1 = tgb_lo16(i2); wk.cv = tgb_hi16(i2);
                    wk.cv1 = tgb_lo16(i3);
                    wk.axis = flags & 3u; wk.negative = (flags >> 2) & 1u; wk.exotic = (flags >> 3) & 1u;
                    const tgb_object_frame& f = p_frames[s_list[(flags >> 4) - 1u]];
                    u32 cx, cy, cz; f32 enter = 0.0f;
                    const bool have = tgb_k1_next_candidate(f, &wk, SF(P_TSKIP, k), &cx, &cy, &cz, &enter);
                    S(P_I0, k) = tgb_pack16(wk.s, wk.n_slices);
                    S(P_I1, k) = tgb_pack16(wk.cu, wk.cu0);
                    S(P_I2, k) = tgb_pack16(wk.cu1, wk.cv);
                    S(P_I3, k) = tgb_pack16(wk.cv1, (i32)flags);
                    S(P_ENTER, k) = __float_as_uint(enter);
                    kinds = (kinds & ~(15u << (4 * k))) | ((have ? (u32)ST_CAND : (u32)ST_NEXT) << (4 * k));
                }
            }
            else if (has_next)
            {
                /* ---- next object of the tile's list this pixel has to walk: depth bound, rectangle, ray vs object ---- */
                u32 k = 0;
#pragma unroll
                for (int j = K - 1; j >= 0; j--) if (((kinds >> (4 * j)) & 15u) == ST_NEXT) k = (u32)j;
                const u32 py = py0 + 4u * k;
                const u32 best_depth24 = S(P_BEST_HI, k) >> (TG_VIS_DEPTH_SHIFT - 32);
                u32 j = S(P_I3, k) >> 20;
                const v3 dir_ws = tgb_pixel_direction(&cam, w, h, px, py);
                u32 st = ST_WINDOW_DONE;
                while (j < n_listed)
                {
                    const tgb_object_frame& f = p_frames[s_list[j]];
                    j++;
                    if (f.min_depth24 > best_depth24)
                    {
                        /* the list is front to back: every later object is behind the best word too */
                        if (sorted) { st = ST_FINISHED; break; }
                        continue;
                    }
                    if ((i32)px < f.x0 || (i32)px > f.x1 || (i32)py < f.y0 || (i32)py > f.y1) continue;
                    tgb_k1_walk wk;
                    if (tgb_k1_setup(f, dir_ws, &wk))
                    {
                        S(P_DX, k) = __float_as_uint(wk.d.x); S(P_DY, k) = __float_as_uint(wk.d.y); S(P_DZ, k) = __float_as_uint(wk.d.z);
                        S(P_TDX, k) = __float_as_uint(wk.t_delta_x); S(P_TDY, k) = __float_as_uint(wk.t_delta_y); S(P_TDZ, k) = __float_as_uint(wk.t_delta_z);
                        S(P_TIN, k) = __float_as_uint(wk.t_in); S(P_TOUT, k) = __float_as_uint(wk.t_out);
                        S(P_I0, k) = tgb_pack16(wk.s, wk.n_slices);
                        S(P_I1, k) = tgb_pack16(wk.cu, wk.cu0);
                        S(P_I2, k) = tgb_pack16(wk.cu1, wk.cv);
                        st = ST_ENUM;
                        S(P_I3, k) = tgb_pack16(wk.cv1, (i32)(wk.axis | (wk.negative << 2) | (wk.exotic << 3) | (j << 4)));
                        break;
                    }
                }
                if (st != ST_ENUM) S(P_I3, k) = (j << 4) << 16;
                kinds = (kinds & ~(15u << (4 * k))) | (st << (4 * k));
            }
        }
        if (base + TGB_K1P_THREADS < n_visible) __syncthreads(); /* s_list is rewritten by the next window */
    }

    /* the buffer keeps rows in virtual order (tgb_rows.h; the identity on one GPU) */
#pragma unroll
    for (int k = 0; k < K; k++)
    {
        const u32 py = py0 + 4u * k;
        const u64 best = ((u64)S(P_BEST_HI, k) << 32) | (u64)S(P_BEST_LO, k);
        if (px < w && py < h && best != TG_VIS_CLEAR)
            atomicMin((unsigned long long*)&p_vis[(u64)tgb_row_to_virtual(py, n_ranks, tile_rows) * w + px], (unsigned long long)best);
    }
#undef S
#undef SF
}

template <int K, int MIN_CTAS>
static b32 tgbd__k1_pool_launch(struct tgb_device* d, const tg_camera_rays* p_cam)
{
    const size_t smem = (size_t)TGB_K1P_WORDS * K * TGB_K1P_THREADS * sizeof(u32);
    static bool attr_set = false;
    if (!attr_set)
    {
        TGB_CUDA(cudaFuncSetAttribute(k_visibility_pool<K, MIN_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const dim3 grid((d->width + TGB_K1P_TILE_W - 1) / TGB_K1P_TILE_W, (d->height + 16 * K - 1) / (16 * K));
    k_visibility_pool<K, MIN_CTAS><<<grid, TGB_K1P_THREADS, smem, d->stream>>>(d->d_frames_sorted, d->d_visible_count, *p_cam, d->width, d->height,
                                                                               d->d_cluster_pointers, d->d_masks, d->global_pointer_base, d->d_vis, d->n_ranks, d->tile_rows);
    TGB_LAUNCH_CHECK(d);
    return TG_TRUE;
}

/* launched after cull + sort (tgbd_render_visibility, tgb_visibility.cu); returns at once when the frame holds an object the packed
 * iterator cannot walk (d_visible_count[2] != 0), in which case the caller's k_visibility launch does the work */
extern "C" b32 tgbd_k1_pool_render(struct tgb_device* d, const tg_camera_rays* p_cam)
{
    const int k = tgbd_env_int("TGB_K1_PIXELS_PER_LANE", 2), min_ctas = tgbd_env_int("TGB_K1_POOL_MIN_CTAS", 0);
    switch (k)
    {
    case 1:  return min_ctas == 3 ? tgbd__k1_pool_launch<1, 3>(d, p_cam) : tgbd__k1_pool_launch<1, 4>(d, p_cam);
    case 3:  return min_ctas == 4 ? tgbd__k1_pool_launch<3, 4>(d, p_cam) : tgbd__k1_pool_launch<3, 3>(d, p_cam);
    case 4:  return min_ctas == 2 ? tgbd__k1_pool_launch<4, 2>(d, p_cam) : tgbd__k1_pool_launch<4, 3>(d, p_cam);
    default: return min_ctas == 3 ? tgbd__k1_pool_launch<2, 3>(d, p_cam) : tgbd__k1_pool_launch<2, 4>(d, p_cam);
    }
}

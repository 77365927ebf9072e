/*
 * tgb_visibility.cu -- K1, the compute-only visibility-buffer pass.
 *
 * Replaces the reference's rasterised cluster-box pass
 *   clear            assets/shaders/raytracer/clear.comp:15-21
 *   coverage         assets/shaders/raytracer/visibility.vert:19-26 + cluster_functions.inc:1-38
 *                    (instanced draw of one box per cluster, tgvk_raytracer.c:1275-1317)
 *   fragment         assets/shaders/raytracer/visibility.frag:22-208
 * whose result is, per pixel, the 64-bit minimum over ALL clusters of the fragment function
 * (coverage only prunes). Here rays are bound to screen tiles, objects are culled
 * and ordered front to back once per frame (k_cull_objects / k_sort_frames), and every ray walks
 * the cluster grid of each surviving object slice by slice along its dominant axis, visiting a
 * CONSERVATIVE SUPERSET of the clusters whose own slab test can succeed. For each visited cluster
 * the arithmetic that produces the written word is the reference's, operation for operation
 * (tgb_hoist.h, tgb_math.h; this TU is compiled with -fmad=false); everything that merely selects
 * candidates may be approximate because a superset yields the identical minimum.
 * Early-outs compare the quantised 24-bit depth of an entry point with the best word so far and
 * skip only on STRICTLY greater (a tie must still run: the lower pointer / voxel wins).
 */
#include <string.h>

#include "tgb_device.cuh"
#include "tgb_k1_walk.cuh"

#define TGB_TILE_W        16
#define TGB_TILE_H        16
#define TGB_SORT_MAX      4096
#define TGB_FULL_MASK     0xFFFFFFFFu

/* ------------------------------------------------------------------------------------------- */
/* clear.comp:19                                                                                */
/* ------------------------------------------------------------------------------------------- */
__global__ void k_clear_visibility(ulonglong2* __restrict__ p_vis2, u64 n_pairs, u64* __restrict__ p_vis, u64 n)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pairs) p_vis2[i] = make_ulonglong2(TG_VIS_CLEAR, TG_VIS_CLEAR);
    if (i == 0 && (n & 1)) p_vis[n - 1] = TG_VIS_CLEAR;
}

extern "C" b32 tgbd_clear(struct tgb_device* d)
{
    TGB_CUDA(cudaSetDevice(d->device));
    if (!tgbd_flush_objects(d)) return TG_FALSE; /* the frame's transform uploads: one copy, in front of the "inputs complete" event below */
    if (d->p2p_ready)
    {
        /* merge over peer memory (tgb_peer.cu): the peers may still read last frame's words, this frame goes to the other pair */
        d->vis_flip ^= 1u;
        d->d_vis = d->d_vis_pair[d->vis_flip];
        d->d_mat = d->d_mat_pair[d->vis_flip];
        d->frame_seq++; /* what this rank publishes when its K1 is done, and what it waits for on its peers (tgb_peer.cu) */
    }
    d->tiles_flagged = TG_FALSE; d->objects_gathered = TG_FALSE;
    /* the frame's inputs are queued (uploads precede the clear in every caller of this library): K2 may start behind this point on its own stream */
    TGB_CUDA(cudaEventRecord(d->ev_inputs, d->stream));
    d->ev_inputs_valid = TG_TRUE; d->inputs_changed_since_clear = TG_FALSE;
    d->vis_merged = TG_FALSE; d->tile_merged = TG_FALSE;
    const u64 n = (u64)d->width * d->tile_rows * d->n_ranks; /* the padded frame */
    TGB_CUDA(cudaEventRecord(d->ev[0], d->stream));
    k_clear_visibility<<<(u32)((n / 2 + 255) / 256) + 1, 256, 0, d->stream>>>((ulonglong2*)d->d_vis, n / 2, d->d_vis, n);
    TGB_LAUNCH_CHECK(d);
    TGB_CUDA(cudaEventRecord(d->ev[1], d->stream));
    d->ev_clear = TG_TRUE;
    return TG_TRUE;
}

/* ------------------------------------------------------------------------------------------- */
/* Object culling: one thread per object slot                                                   */
/* ------------------------------------------------------------------------------------------- */
struct tgb_pinhole
{
    f64 inv[9];  /* inverse of [br-bl | tl-bl | bl], row-major */
    f64 cam[3];
    f64 rx[3], uy[3], fz[3]; /* orthonormal camera frame: right (bl->br), up (bl->tl), forward */
    f64 c0, bx, by, lu, lv;  /* image-plane distance, bl in the frame, |br-bl|, |tl-bl| */
    i32 ok;
};

static void tgb_pinhole_init(const tg_camera_rays* c, tgb_pinhole* p)
{
    const f64 bl[3] = { c->ray_bl.x, c->ray_bl.y, c->ray_bl.z };
    const f64 u[3]  = { c->ray_br.x - bl[0], c->ray_br.y - bl[1], c->ray_br.z - bl[2] };
    const f64 v[3]  = { c->ray_tl.x - bl[0], c->ray_tl.y - bl[1], c->ray_tl.z - bl[2] };
    const f64 a = u[0], b = v[0], cc = bl[0];
    const f64 d = u[1], e = v[1], f = bl[1];
    const f64 g = u[2], h = v[2], i = bl[2];
    const f64 det = a * (e * i - f * h) - b * (d * i - f * g) + cc * (d * h - e * g);
    p->ok = det != 0.0;
    const f64 id = p->ok ? 1.0 / det : 0.0;
    p->inv[0] = (e * i - f * h) * id; p->inv[1] = (cc * h - b * i) * id; p->inv[2] = (b * f - cc * e) * id;
    p->inv[3] = (f * g - d * i) * id; p->inv[4] = (a * i - cc * g) * id; p->inv[5] = (cc * d - a * f) * id;
    p->inv[6] = (d * h - e * g) * id; p->inv[7] = (b * g - a * h) * id;  p->inv[8] = (a * e - b * d) * id;
    p->cam[0] = c->camera.x; p->cam[1] = c->camera.y; p->cam[2] = c->camera.z;
    p->lu = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    p->lv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (!(p->lu > 0.0) || !(p->lv > 0.0)) { p->ok = 0; return; }
    for (int k = 0; k < 3; k++) { p->rx[k] = u[k] / p->lu; p->uy[k] = v[k] / p->lv; }
    p->fz[0] = p->rx[1] * p->uy[2] - p->rx[2] * p->uy[1];
    p->fz[1] = p->rx[2] * p->uy[0] - p->rx[0] * p->uy[2];
    p->fz[2] = p->rx[0] * p->uy[1] - p->rx[1] * p->uy[0];
    f64 lf = sqrt(p->fz[0] * p->fz[0] + p->fz[1] * p->fz[1] + p->fz[2] * p->fz[2]);
    if (!(lf > 0.0)) { p->ok = 0; return; }
    p->c0 = (bl[0] * p->fz[0] + bl[1] * p->fz[1] + bl[2] * p->fz[2]) / lf;
    if (p->c0 < 0.0) { lf = -lf; p->c0 = -p->c0; }
    for (int k = 0; k < 3; k++) p->fz[k] /= lf;
    p->bx = bl[0] * p->rx[0] + bl[1] * p->rx[1] + bl[2] * p->rx[2];
    p->by = bl[0] * p->uy[0] + bl[1] * p->uy[1] + bl[2] * p->uy[2];
    if (!(p->c0 > 0.0)) p->ok = 0;
}

/* Range of tan(angle) over a circle of radius r around (a, z) seen from the origin of that plane; a side is
 * reported only when it is safely inside the forward half plane. */
__device__ void tgb_tangent_bounds(f64 a, f64 z, f64 r, f64* p_lo, f64* p_hi, bool* p_lo_ok, bool* p_hi_ok)
{
    const f64 d2 = a * a + z * z;
    *p_lo_ok = false; *p_hi_ok = false; *p_lo = 0.0; *p_hi = 0.0;
    if (d2 <= r * r * 1.0001) return;
    const f64 theta = atan2(a, z), alpha = asin(r / sqrt(d2));
    const f64 lo = theta - alpha, hi = theta + alpha, lim = 1.5607963267948966; /* pi/2 - 0.01 */
    if (lo > -lim && lo < lim) { *p_lo = tan(lo); *p_lo_ok = true; }
    if (hi > -lim && hi < lim) { *p_hi = tan(hi); *p_hi_ok = true; }
}

__global__ void k_cull_objects(const tg_object_data* __restrict__ p_objects, u32 object_capacity, tg_camera_rays cam, tgb_pinhole pin,
                               u32 w, u32 h, tgb_object_frame* __restrict__ p_frames, u32* __restrict__ p_count)
{
    const u32 object_idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (object_idx >= object_capacity) return;
    const tg_object_data o = p_objects[object_idx];
    /* tgvk_raytracer.c:1069-1077: initialised iff all dims != 0 */
    if (o.n_cluster_pointers_per_dim.x == 0 || o.n_cluster_pointers_per_dim.y == 0 || o.n_cluster_pointers_per_dim.z == 0) return;

    tgb_object_frame f;
    const v3 camera = tgb_v3(cam.camera.x, cam.camera.y, cam.camera.z);
    tgb_hoist_object(&o, camera, &f);
    f.object_idx = object_idx;

    tgb_frame_conservative(&f, &o, camera); /* og, eps */
    const f32 ext[3] = { 8.0f * (f32)f.nx, 8.0f * (f32)f.ny, 8.0f * (f32)f.nz };

    /* distance from the camera to the (inflated) box, in the grid frame (rigid transform) */
    f64 dist2 = 0.0;
    for (int k = 0; k < 3; k++)
    {
        const f64 x = f.og[k];
        const f64 lo = -(f64)f.eps, hi = (f64)ext[k] + (f64)f.eps;
        const f64 dd = x < lo ? lo - x : (x > hi ? x - hi : 0.0);
        dist2 += dd * dd;
    }
    const f64 dist = sqrt(dist2) * (1.0 - 1e-5) - 0.02;
    const f64 far_plane = (f64)cam.far_plane;
    if (dist > far_plane * (1.0 + 1e-5)) return; /* every hit would have d > 1 (visibility.frag:194) */
    {
        f64 q = dist <= 0.0 ? 0.0 : floor(dist / far_plane * 16777215.0) - 2.0;
        if (q < 0.0) q = 0.0;
        if (q > 16777215.0) q = 16777215.0;
        f.min_depth24 = (u32)q;
    }

    /* conservative screen rectangle of the inflated box */
    const f64 m = (f64)f.eps + 0.0625;
    f64 minx = 1e300, miny = 1e300, maxx = -1e300, maxy = -1e300;
    bool full = !pin.ok;
    int n_behind = 0;
    for (int k = 0; k < 8; k++)
    {
        const f64 lx = ((k & 1) ? (f64)ext[0] + m : -m) - (f64)f.half[0];
        const f64 ly = ((k & 2) ? (f64)ext[1] + m : -m) - (f64)f.half[1];
        const f64 lz = ((k & 4) ? (f64)ext[2] + m : -m) - (f64)f.half[2];
        const f64 X = (f64)o.rotation.m00 * lx + (f64)o.rotation.m01 * ly + (f64)o.rotation.m02 * lz + (f64)o.translation.x - pin.cam[0];
        const f64 Y = (f64)o.rotation.m10 * lx + (f64)o.rotation.m11 * ly + (f64)o.rotation.m12 * lz + (f64)o.translation.y - pin.cam[1];
        const f64 Z = (f64)o.rotation.m20 * lx + (f64)o.rotation.m21 * ly + (f64)o.rotation.m22 * lz + (f64)o.translation.z - pin.cam[2];
        const f64 a = pin.inv[0] * X + pin.inv[1] * Y + pin.inv[2] * Z;
        const f64 b = pin.inv[3] * X + pin.inv[4] * Y + pin.inv[5] * Z;
        const f64 c = pin.inv[6] * X + pin.inv[7] * Y + pin.inv[8] * Z;
        const f64 len = fabs(X) + fabs(Y) + fabs(Z);
        if (!(c > 1e-4 * len) || !(c > 1e-9))
        {
            full = true;
            if (c < -1e-4 * len) n_behind++;
            continue;
        }
        const f64 px = (a / c) * (f64)w - 0.5;
        const f64 py = (1.0 - b / c) * (f64)h - 0.5;
        minx = fmin(minx, px); maxx = fmax(maxx, px);
        miny = fmin(miny, py); maxy = fmax(maxy, py);
    }
    if (n_behind == 8) return; /* entirely behind the image plane: exit <= 0 for every ray */
    if (full)
    {
        /* some corner is beside / behind the eye: bound the projection of the bounding sphere instead */
        f64 x0 = 0.0, y0 = 0.0, x1 = (f64)w - 1.0, y1 = (f64)h - 1.0;
        if (pin.ok)
        {
            const f64 Cx = (f64)o.translation.x - pin.cam[0], Cy = (f64)o.translation.y - pin.cam[1], Cz = (f64)o.translation.z - pin.cam[2];
            const f64 hx = (f64)f.half[0] + m, hy = (f64)f.half[1] + m, hz = (f64)f.half[2] + m;
            const f64 r = sqrt(hx * hx + hy * hy + hz * hz) * (1.0 + 1e-6) + 1e-3;
            const f64 X = Cx * pin.rx[0] + Cy * pin.rx[1] + Cz * pin.rx[2];
            const f64 Y = Cx * pin.uy[0] + Cy * pin.uy[1] + Cz * pin.uy[2];
            const f64 Z = Cx * pin.fz[0] + Cy * pin.fz[1] + Cz * pin.fz[2];
            if (Z + r < 0.0) return;
            f64 lo, hi; bool lo_ok, hi_ok;
            tgb_tangent_bounds(X, Z, r, &lo, &hi, &lo_ok, &hi_ok);
            if (lo_ok) x0 = fmax(x0, floor(((pin.c0 * lo - pin.bx) / pin.lu) * (f64)w - 0.5) - 2.0);
            if (hi_ok) x1 = fmin(x1, ceil(((pin.c0 * hi - pin.bx) / pin.lu) * (f64)w - 0.5) + 2.0);
            tgb_tangent_bounds(Y, Z, r, &lo, &hi, &lo_ok, &hi_ok);
            if (hi_ok) y0 = fmax(y0, floor((1.0 - (pin.c0 * hi - pin.by) / pin.lv) * (f64)h - 0.5) - 2.0);
            if (lo_ok) y1 = fmin(y1, ceil((1.0 - (pin.c0 * lo - pin.by) / pin.lv) * (f64)h - 0.5) + 2.0);
        }
        if (x1 < x0 || y1 < y0) return;
        f.x0 = (i32)x0; f.y0 = (i32)y0; f.x1 = (i32)x1; f.y1 = (i32)y1;
    }
    else
    {
        minx = floor(minx) - 2.0; miny = floor(miny) - 2.0; maxx = ceil(maxx) + 2.0; maxy = ceil(maxy) + 2.0;
        if (minx < 0.0) minx = 0.0;
        if (miny < 0.0) miny = 0.0;
        if (maxx > (f64)w - 1.0) maxx = (f64)w - 1.0;
        if (maxy > (f64)h - 1.0) maxy = (f64)h - 1.0;
        if (maxx < minx || maxy < miny) return;
        f.x0 = (i32)minx; f.y0 = (i32)miny; f.x1 = (i32)maxx; f.y1 = (i32)maxy;
    }

    if (f.nx > 32767u || f.ny > 32767u || f.nz > 32767u) atomicOr(&p_count[2], 1u); /* k_visibility_pool packs its iterator into 16-bit fields: this frame takes k_visibility */
    const u32 slot = atomicAdd(p_count, 1u);
    p_frames[slot] = f;
}

/* Front-to-back order of the surviving objects: one CTA, bitonic sort of (min_depth24, object, slot) keys in shared memory. */
__global__ void __launch_bounds__(1024) k_sort_frames(const tgb_object_frame* __restrict__ p_frames, tgb_object_frame* __restrict__ p_sorted, u32* __restrict__ p_count)
{
    __shared__ u64 s_keys[TGB_SORT_MAX];
    const u32 n = p_count[0];
    if (n > TGB_SORT_MAX)
    {
        /* too many survivors for the single-CTA sort: keep arrival order, K1 then must not break early */
        const u32 n_words = n * (u32)(sizeof(tgb_object_frame) / 4);
        const u32* p_src = (const u32*)p_frames;
        u32* p_dst = (u32*)p_sorted;
        for (u32 i = threadIdx.x; i < n_words; i += blockDim.x) p_dst[i] = p_src[i];
        if (threadIdx.x == 0) p_count[1] = 0;
        return;
    }
    u32 n_pow2 = 1;
    while (n_pow2 < n) n_pow2 <<= 1;
    for (u32 i = threadIdx.x; i < n_pow2; i += blockDim.x)
    {
        s_keys[i] = i < n ? (((u64)p_frames[i].min_depth24 << 40) | ((u64)p_frames[i].object_idx << 16) | (u64)i) : ~0ull;
    }
    __syncthreads();
    for (u32 k = 2; k <= n_pow2; k <<= 1)
    {
        for (u32 j = k >> 1; j > 0; j >>= 1)
        {
            for (u32 i = threadIdx.x; i < n_pow2; i += blockDim.x)
            {
                const u32 ixj = i ^ j;
                if (ixj > i)
                {
                    const u64 a = s_keys[i], b = s_keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { s_keys[i] = b; s_keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    const u32 words_per = (u32)(sizeof(tgb_object_frame) / 4);
    const u32* p_src = (const u32*)p_frames;
    u32* p_dst = (u32*)p_sorted;
    for (u32 i = threadIdx.x; i < n * words_per; i += blockDim.x)
    {
        const u32 dst_obj = i / words_per, word = i % words_per;
        const u32 slot = (u32)(s_keys[dst_obj] & 0xFFFFu);
        p_dst[i] = p_src[slot * words_per + word];
    }
    if (threadIdx.x == 0) p_count[1] = 1;
}

/* ------------------------------------------------------------------------------------------- */
/* K1                                                                                           */
/* ------------------------------------------------------------------------------------------- */

/*
 * One ray against one object: enumerate, slice by slice along the dominant axis of d (front to
 * back), every cluster whose box inflated by eps the ray can touch, and visit each.
 */
/* 16 bytes global -> shared without passing through registers (LDGSTS) */
__device__ __forceinline__ void tgb_k1_cp_async16(void* p_shared, const void* p_global)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((u32)__cvta_generic_to_shared(p_shared)), "l"(p_global) : "memory");
}

/*
 * STAGE (TGB_K1_STAGE_MASKS=1, the measured alternative north_star (a) names; off by default): the 64-byte mask of the cluster about to be
 * marched is copied into shared memory by cp.async -- once per group of lanes that march the SAME cluster (__match_any_sync on the cluster
 * index: neighbouring rays mostly do), by the group's first lane -- and the march reads its z-slices from there. Bit-identical by
 * construction (pure data movement).
 */
template <bool REGROUP, bool DEFER, bool STAGE>
__device__ __forceinline__ void tgb_trace_object(const tgb_object_frame& f, v3 dir_ws, f32 far_plane,
                                                 const u32* __restrict__ p_cluster_pointers, const u32* __restrict__ p_masks,
                                                 u32 global_pointer_base, u64& best, f32& t_skip, u32* p_warp_masks)
{
    const f32 e = f.eps;
    /* ws2ms * (dir_ws, 0) before normalisation (tgb_hoist_direction) */
    v3 raw;
    raw.x = (dir_ws.x * f.c[0] + dir_ws.y * f.c[1]) + dir_ws.z * f.c[2];
    raw.y = (dir_ws.x * f.c[3] + dir_ws.y * f.c[4]) + dir_ws.z * f.c[5];
    raw.z = (dir_ws.x * f.c[6] + dir_ws.y * f.c[7]) + dir_ws.z * f.c[8];
    {
        /* Cheap reject before the exact set-up (a square root and six IEEE divisions): slab test of the object's box with the
         * un-normalised direction (the test is scale-free) and approximate reciprocals. The box is inflated by 2 eps plus
         * 1e-4 of the magnitudes involved, three orders above the error of the approximate quotients and of either path. */
        const f32 ex = 8.0f * (f32)f.nx, ey = 8.0f * (f32)f.ny, ez = 8.0f * (f32)f.nz;
        const f32 m = 2.0f * e + 1e-4f * (fabsf(f.og[0]) + fabsf(f.og[1]) + fabsf(f.og[2]) + ex + ey + ez);
        f32 t0 = 0.0f, t1 = TG_F32_MAX;
        bool out = false;
        if (fabsf(raw.x) > 1e-20f) { const f32 i = __fdividef(1.0f, raw.x), a = (-m - f.og[0]) * i, b2 = (ex + m - f.og[0]) * i; t0 = fmaxf(t0, fminf(a, b2)); t1 = fminf(t1, fmaxf(a, b2)); }
        else out = out || f.og[0] < -m || f.og[0] > ex + m;
        if (fabsf(raw.y) > 1e-20f) { const f32 i = __fdividef(1.0f, raw.y), a = (-m - f.og[1]) * i, b2 = (ey + m - f.og[1]) * i; t0 = fmaxf(t0, fminf(a, b2)); t1 = fminf(t1, fmaxf(a, b2)); }
        else out = out || f.og[1] < -m || f.og[1] > ey + m;
        if (fabsf(raw.z) > 1e-20f) { const f32 i = __fdividef(1.0f, raw.z), a = (-m - f.og[2]) * i, b2 = (ez + m - f.og[2]) * i; t0 = fmaxf(t0, fminf(a, b2)); t1 = fminf(t1, fmaxf(a, b2)); }
        else out = out || f.og[2] < -m || f.og[2] > ez + m;
        if (out || t0 > t1 * 1.0001f + 1e-30f) return;
    }
    tgb_ray_in_object r;
    r.d = tgb_normalize(raw); /* exact d_ms (tgb_hoist_direction), shared by all clusters of the object */
    const v3 d = r.d;
    const f32 adx = fabsf(d.x), ady = fabsf(d.y), adz = fabsf(d.z);
    /* visibility.frag:105-136: t_delta = 1 / d or 1 / -d, absent axis F32_MAX */
    r.t_delta_x = adx != 0.0f ? __frcp_rn(adx) : TG_F32_MAX;
    r.t_delta_y = ady != 0.0f ? __frcp_rn(ady) : TG_F32_MAX;
    r.t_delta_z = adz != 0.0f ? __frcp_rn(adz) : TG_F32_MAX;
    r.rx = d.x < 0.0f ? -r.t_delta_x : r.t_delta_x;
    r.ry = d.y < 0.0f ? -r.t_delta_y : r.t_delta_y;
    r.rz = d.z < 0.0f ? -r.t_delta_z : r.t_delta_z;
    r.exotic = (adx != 0.0f && adx < 1e-30f) || (ady != 0.0f && ady < 1e-30f) || (adz != 0.0f && adz < 1e-30f);

    /* permute so that axis k is the dominant one */
    const int k = (adx >= ady && adx >= adz) ? 0 : (ady >= adz ? 1 : 2);
    const f32 dk = k == 0 ? d.x : (k == 1 ? d.y : d.z);
    const f32 du = k == 0 ? d.y : (k == 1 ? d.z : d.x);
    const f32 dv = k == 0 ? d.z : (k == 1 ? d.x : d.y);
    const f32 ok = k == 0 ? f.og[0] : (k == 1 ? f.og[1] : f.og[2]);
    const f32 ou = k == 0 ? f.og[1] : (k == 1 ? f.og[2] : f.og[0]);
    const f32 ov = k == 0 ? f.og[2] : (k == 1 ? f.og[0] : f.og[1]);
    const i32 nk = (i32)(k == 0 ? f.nx : (k == 1 ? f.ny : f.nz));
    const i32 nu = (i32)(k == 0 ? f.ny : (k == 1 ? f.nz : f.nx));
    const i32 nv = (i32)(k == 0 ? f.nz : (k == 1 ? f.nx : f.ny));
    if (!(fabsf(dk) > 0.5f)) return; /* |d| == 1 => dominant component >= 0.577; false only for NaN directions */

    /* conservative slab of the inflated object box; u / v slabs only when the ray is not parallel to them */
    const f32 inv_dk = k == 0 ? r.rx : (k == 1 ? r.ry : r.rz);
    f32 t_in, t_out;
    {
        const f32 ta = (-e - ok) * inv_dk, tb = (8.0f * (f32)nk + e - ok) * inv_dk;
        t_in = fminf(ta, tb); t_out = fmaxf(ta, tb);
    }
    if (fabsf(du) > 1e-20f)
    {
        const f32 inv = k == 0 ? r.ry : (k == 1 ? r.rz : r.rx);
        const f32 ta = (-e - ou) * inv, tb = (8.0f * (f32)nu + e - ou) * inv;
        t_in = fmaxf(t_in, fminf(ta, tb)); t_out = fminf(t_out, fmaxf(ta, tb));
    }
    else if (ou < -e || ou > 8.0f * (f32)nu + e) return;
    if (fabsf(dv) > 1e-20f)
    {
        const f32 inv = k == 0 ? r.rz : (k == 1 ? r.rx : r.ry);
        const f32 ta = (-e - ov) * inv, tb = (8.0f * (f32)nv + e - ov) * inv;
        t_in = fmaxf(t_in, fminf(ta, tb)); t_out = fminf(t_out, fmaxf(ta, tb));
    }
    else if (ov < -e || ov > 8.0f * (f32)nv + e) return;
    /* slack: relative 2^-16 of |t| plus eps (positions move by at most |t|*2^-16 + eps) */
    t_in  -= e + 1.52587890625e-5f * fabsf(t_in);
    t_out += e + 1.52587890625e-5f * fabsf(t_out);
    t_in = fmaxf(t_in, 0.0f);
    if (!(t_in <= t_out)) return;

    const f32 pk_in = ok + t_in * dk, pk_out = ok + t_out * dk;
    const i32 sgn = dk > 0.0f ? 1 : -1;
    i32 s     = (i32)floorf((pk_in  - (f32)sgn * (2.0f * e)) * 0.125f);
    i32 s_end = (i32)floorf((pk_out + (f32)sgn * (2.0f * e)) * 0.125f);
    s     = max(0, min(nk - 1, s));
    s_end = max(0, min(nk - 1, s_end));

    if (REGROUP)
    {
        /*
         * The same enumeration as a per-lane iterator: every lane first advances to its NEXT cluster that passes the first
         * half (the inner loop; cheap), the lanes reconverge where it ends, and the march -- two thirds of K1's instructions --
         * runs with all the lanes that hold a cluster instead of those whose k-th enumerated cluster happens to be one
         * (measured: 7 of 32). Same clusters, same order per lane, same arithmetic.
         */
        i32 n_slices = (s_end - s) * sgn + 1;
        s -= sgn;
        i32 cu = 0, cu0 = 0, cu1 = -1, cv = 0, cv1 = -1;
        /*
         * DEFER (off by default, see the launcher): the word of a voxel found (a slab test with an IEEE division, the depth quantisation,
         * the packing: a seventh of the kernel's instructions) normally runs right after each march, with the 6.6 lanes of 32 that have
         * just found one. Deferred, the walk goes on with an UPPER bound of the t_skip that word would give (tgb_cluster_t_skip_bound; a looser bound only admits
         * more candidates, the minimum is the same), and the exact word is computed once the object is finished -- by all the lanes
         * that found a voxel in it, together. A second voxel found meanwhile (a tie, a rounding sliver) first settles the pending one.
         */
        const bool can_defer = DEFER && ((f.nx | f.ny | f.nz) < 65536u);
        u32 pend0 = 0, pend1 = 0; /* cx | cy << 16, cz | voxel << 16 | 1 << 31 */
        for (;;)
        {
            f32 enter = 0.0f;
            u32 cx = 0, cy = 0, cz = 0;
            bool have = false;
            for (;;)
            {
                if (cu < cu1) cu++;
                else if (cv < cv1) { cv++; cu = cu0; }
                else
                {
                    if (n_slices <= 0) break;
                    n_slices--;
                    s += sgn;
                    cu1 = -1; cv1 = -1; cu = 0; cv = 0; /* empty until the ranges are known */
                    const f32 ta = (8.0f * (f32)s - 2.0f * e - ok) * inv_dk, tb = (8.0f * (f32)(s + 1) + 2.0f * e - ok) * inv_dk;
                    const f32 t0 = fmaxf(fminf(ta, tb), t_in), t1 = fminf(fmaxf(ta, tb), t_out);
                    if (!(t0 <= t1)) continue;
                    /* slices are visited with non-decreasing t0: once even the slice entry is behind the best hit, stop */
                    if (t0 - (4.0f * e + 3.0517578125e-5f * fabsf(t0)) > t_skip) { n_slices = 0; break; }
                    const f32 ua = ou + t0 * du, ub = ou + t1 * du;
                    const f32 va = ov + t0 * dv, vb = ov + t1 * dv;
                    const f32 pad = 2.0f * e + 3.0517578125e-5f * (fabsf(ou) + fabsf(ov) + t1);
                    cu0 = max(0, (i32)floorf((fminf(ua, ub) - pad) * 0.125f));
                    const i32 u1 = min(nu - 1, (i32)floorf((fmaxf(ua, ub) + pad) * 0.125f));
                    const i32 v0 = max(0, (i32)floorf((fminf(va, vb) - pad) * 0.125f));
                    const i32 v1 = min(nv - 1, (i32)floorf((fmaxf(va, vb) + pad) * 0.125f));
                    if (cu0 > u1 || v0 > v1) continue;
                    cu = cu0; cu1 = u1; cv = v0; cv1 = v1;
                }
                cx = (u32)(k == 0 ? s : (k == 1 ? cv : cu));
                cy = (u32)(k == 0 ? cu : (k == 1 ? s : cv));
                cz = (u32)(k == 0 ? cv : (k == 1 ? cu : s));
                if (tgb_cluster_candidate(f, r, cx, cy, cz, t_skip, &enter)) { have = true; break; }
            }
            if (!have) break;
            i32 voxel;
            if (STAGE)
            {
                const u32 lane = threadIdx.x & 31u;
                const u32 cluster_idx = __ldg(&p_cluster_pointers[f.first_cluster_pointer + cx + f.nx * (cy + f.ny * cz)]);
                const u32 peers = __match_any_sync(__activemask(), cluster_idx); /* the lanes about to march this cluster */
                const u32 leader = (u32)(__ffs(peers) - 1);
                u32* p_slot = p_warp_masks + leader * TG_CLUSTER_MASK_WORDS;
                if (lane == leader)
                {
                    const u32* p_src = p_masks + (u64)cluster_idx * TG_CLUSTER_MASK_WORDS;
                    tgb_k1_cp_async16(p_slot, p_src); tgb_k1_cp_async16(p_slot + 4, p_src + 4);
                    tgb_k1_cp_async16(p_slot + 8, p_src + 8); tgb_k1_cp_async16(p_slot + 12, p_src + 12);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp(peers);
                voxel = tgb_cluster_find_in<true>(f, r, cx, cy, cz, enter, reinterpret_cast<const tgb_slice*>(p_slot));
                __syncwarp(peers); /* the leader's next candidate reuses its slot */
            }
            else voxel = tgb_cluster_find(f, r, cx, cy, cz, enter, p_cluster_pointers, p_masks);
            if (voxel >= 0)
            {
                if (!can_defer) tgb_cluster_word(f, r, cx, cy, cz, voxel, far_plane, global_pointer_base, best, t_skip);
                else
                {
                    if (pend1 >> 31) tgb_cluster_word(f, r, pend0 & 0xFFFFu, pend0 >> 16, pend1 & 0xFFFFu, (i32)((pend1 >> 16) & 511u), far_plane, global_pointer_base, best, t_skip);
                    pend0 = cx | (cy << 16); pend1 = cz | ((u32)voxel << 16) | 0x80000000u;
                    t_skip = fminf(t_skip, tgb_cluster_t_skip_bound(f, r, cx, cy, cz, voxel, far_plane));
                }
            }
        }
        if (pend1 >> 31) tgb_cluster_word(f, r, pend0 & 0xFFFFu, pend0 >> 16, pend1 & 0xFFFFu, (i32)((pend1 >> 16) & 511u), far_plane, global_pointer_base, best, t_skip);
    }
    else
    for (;; s += sgn)
    {
        const f32 ta = (8.0f * (f32)s - 2.0f * e - ok) * inv_dk, tb = (8.0f * (f32)(s + 1) + 2.0f * e - ok) * inv_dk;
        const f32 t0 = fmaxf(fminf(ta, tb), t_in), t1 = fminf(fmaxf(ta, tb), t_out);
        if (t0 <= t1)
        {
            /* slices are visited with non-decreasing t0: once even the slice entry is behind the best hit, stop */
            if (t0 - (4.0f * e + 3.0517578125e-5f * fabsf(t0)) > t_skip) return;
            const f32 ua = ou + t0 * du, ub = ou + t1 * du;
            const f32 va = ov + t0 * dv, vb = ov + t1 * dv;
            const f32 pad = 2.0f * e + 3.0517578125e-5f * (fabsf(ou) + fabsf(ov) + t1);
            const i32 cu0 = max(0,      (i32)floorf((fminf(ua, ub) - pad) * 0.125f));
            const i32 cu1 = min(nu - 1, (i32)floorf((fmaxf(ua, ub) + pad) * 0.125f));
            const i32 cv0 = max(0,      (i32)floorf((fminf(va, vb) - pad) * 0.125f));
            const i32 cv1 = min(nv - 1, (i32)floorf((fmaxf(va, vb) + pad) * 0.125f));
            for (i32 cv = cv0; cv <= cv1; cv++)
            {
                for (i32 cu = cu0; cu <= cu1; cu++)
                {
                    const u32 cx = (u32)(k == 0 ? s : (k == 1 ? cv : cu));
                    const u32 cy = (u32)(k == 0 ? cu : (k == 1 ? s : cv));
                    const u32 cz = (u32)(k == 0 ? cv : (k == 1 ? cu : s));
                    f32 enter;
                    if (tgb_cluster_candidate(f, r, cx, cy, cz, t_skip, &enter))
                        tgb_cluster_march(f, r, cx, cy, cz, enter, far_plane, p_cluster_pointers, p_masks, global_pointer_base, best, t_skip);
                }
            }
        }
        if (s == s_end) break;
    }
}

/*
 * One CTA = one 16x16 pixel tile, one warp = one 8x4 pixel block (coherent rays), one lane = one ray. The CTA first
 * compacts, window by window, the (front-to-back sorted) objects whose screen rectangle overlaps its tile into shared
 * memory, so a warp only walks the few objects near it. A lane retires from the object loop once the next object's
 * lower depth bound exceeds its best word, a warp once all its lanes did. The resolve is a single 64-bit atomicMin per
 * hit pixel (visibility.frag:206) so that other passes / shards may target the same buffer.
 */
#define TGB_K1_THREADS 256
/*
 * SHARDED (one process per GPU, merge over peer memory, tgb_peer.cu): the CTA records whether its 16x16 tile has any hit. The
 * material pass (k_resolve_material_tiles) and the peers' k_merge_tile skip the tiles that have none -- with the objects dealt out
 * over N ranks that is most of a rank's frame. (Resolving the material right here, in K1's epilogue, was measured: the four
 * dependent loads at the end of every CTA's life cost more, +0.07 ms at N = 2, than the separate pass over the flagged tiles.)
 */
struct tgb_k1_shard_args
{
    u32* __restrict__ p_tile_flags; /* [virtual band][tile column] */
    u32 tiles_x;
};

template <int MIN_CTAS, bool REGROUP, bool SHARDED, bool DEFER, bool STAGE = false>
__global__ void __launch_bounds__(TGB_K1_THREADS, MIN_CTAS) k_visibility(const tgb_object_frame* __restrict__ p_frames, const u32* __restrict__ p_count,
                                                               tg_camera_rays cam, u32 w, u32 h,
                                                               const u32* __restrict__ p_cluster_pointers, const u32* __restrict__ p_masks,
                                                               u32 global_pointer_base, u64* __restrict__ p_vis, u32 n_ranks, u32 tile_rows, u32 fallback_only,
                                                               const tgb_k1_shard_args sh)
{
    __shared__ u32 s_list[TGB_K1_THREADS];
    __shared__ u32 s_warp_count[TGB_K1_THREADS / 32];
    __shared__ __align__(16) u32 s_staged_masks[STAGE ? TGB_K1_THREADS * TG_CLUSTER_MASK_WORDS : 4]; /* STAGE: one 64-byte slot per lane */

    const u32 n_visible = p_count[0];
    if (n_visible == 0)
    {
        if (SHARDED && threadIdx.x == 0) sh.p_tile_flags[(tgb_row_to_virtual(blockIdx.y * TGB_TILE_H, n_ranks, tile_rows) / TGB_BAND_ROWS) * sh.tiles_x + blockIdx.x] = 0u;
        return;
    }
    if ((fallback_only & 1u) && p_count[2] == 0) return; /* k_visibility_pool (tgb_visibility_pool.cu) rendered this frame */
    const bool sorted = p_count[1] != 0;

    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const i32 tx0 = (i32)(blockIdx.x * TGB_TILE_W), ty0 = (i32)(blockIdx.y * TGB_TILE_H);
    const i32 tx1 = (i32)min(blockIdx.x * TGB_TILE_W + TGB_TILE_W - 1u, w - 1u), ty1 = (i32)min(blockIdx.y * TGB_TILE_H + TGB_TILE_H - 1u, h - 1u);
    const u32 wx0 = blockIdx.x * TGB_TILE_W + (warp & 1u) * 8u;
    const u32 wy0 = blockIdx.y * TGB_TILE_H + (warp >> 1) * 4u;
    const bool warp_on_screen = wx0 < w && wy0 < h; /* warp-uniform */
    const u32 px = wx0 + (lane & 7u), py = wy0 + (lane >> 3);
    const bool in_screen = px < w && py < h;
    const i32 wx1 = (i32)min(wx0 + 7u, w - 1u), wy1 = (i32)min(wy0 + 3u, h - 1u);

    /* A tile no visible object's rectangle touches (most tiles of a rank that holds 1 / N of the objects) leaves after one barrier,
     * before anything per-ray is computed. One window covers up to TGB_K1_THREADS visible objects. */
    if (n_visible <= TGB_K1_THREADS && !(fallback_only & 2u)) /* bit 1: TGB_K1_EARLY_EXIT=0, the measured alternative */
    {
        bool touched = false;
        if (threadIdx.x < n_visible)
        {
            const tgb_object_frame& f = p_frames[threadIdx.x];
            touched = !(f.x1 < tx0 || f.x0 > tx1 || f.y1 < ty0 || f.y0 > ty1);
        }
        if (!__syncthreads_or(touched ? 1 : 0))
        {
            if (SHARDED && threadIdx.x == 0) sh.p_tile_flags[(tgb_row_to_virtual(blockIdx.y * TGB_TILE_H, n_ranks, tile_rows) / TGB_BAND_ROWS) * sh.tiles_x + blockIdx.x] = 0u;
            return;
        }
    }

    const v3 dir_ws = tgb_pixel_direction(&cam, w, h, in_screen ? px : min(wx0, w - 1u), in_screen ? py : min(wy0, h - 1u));
    u64 best = TG_VIS_CLEAR;
    f32 t_skip = TG_F32_MAX;
    bool warp_done = !warp_on_screen;

    for (u32 base = 0; base < n_visible; base += TGB_K1_THREADS)
    {
        /* order-preserving compaction of this window's objects that overlap the tile */
        const u32 i = base + threadIdx.x;
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
        for (u32 k = 0; k < TGB_K1_THREADS / 32; k++)
        {
            const u32 c = s_warp_count[k];
            offset += k < warp ? c : 0u;
            n_listed += c;
        }
        if (overlaps) s_list[offset + (u32)__popc(ballot & ((1u << lane) - 1u))] = i;
        __syncthreads();

        if (!warp_done)
        {
            for (u32 j = 0; j < n_listed; j++)
            {
                const tgb_object_frame& f = p_frames[s_list[j]];
                const bool behind_best = !in_screen || (u64)f.min_depth24 > (best >> TG_VIS_DEPTH_SHIFT);
                if (sorted && __all_sync(TGB_FULL_MASK, behind_best)) { warp_done = true; break; }
                if (f.x1 < (i32)wx0 || f.x0 > wx1 || f.y1 < (i32)wy0 || f.y0 > wy1) continue; /* warp-uniform */
                if (behind_best || (i32)px < f.x0 || (i32)px > f.x1 || (i32)py < f.y0 || (i32)py > f.y1) continue;
                tgb_trace_object<REGROUP, DEFER, STAGE>(f, dir_ws, cam.far_plane, p_cluster_pointers, p_masks, global_pointer_base, best, t_skip,
                                                        s_staged_masks + (STAGE ? warp * 32u * TG_CLUSTER_MASK_WORDS : 0u));
            }
        }
        if (base + TGB_K1_THREADS < n_visible) __syncthreads(); /* s_list is rewritten by the next window */
    }

    /* the buffer keeps rows in virtual order (tgb_rows.h; the identity on one GPU) */
    const bool hit = in_screen && best != TG_VIS_CLEAR;
    const u64 pixel = (u64)tgb_row_to_virtual(in_screen ? py : 0u, n_ranks, tile_rows) * w + px;
    if (hit) atomicMin((unsigned long long*)&p_vis[pixel], (unsigned long long)best);
    if (SHARDED)
    {
        const int any_hit = __syncthreads_or(hit ? 1 : 0);
        if (threadIdx.x == 0) sh.p_tile_flags[(tgb_row_to_virtual(blockIdx.y * TGB_TILE_H, n_ranks, tile_rows) / TGB_BAND_ROWS) * sh.tiles_x + blockIdx.x] = any_hit ? 1u : 0u;
    }
}

extern "C" b32 tgbd_render_visibility(struct tgb_device* d, const tg_camera_rays* p_cam, u32 object_capacity)
{
    TGB_CUDA(cudaSetDevice(d->device));
    if (!tgbd_flush_objects(d)) return TG_FALSE;
    tgb_pinhole pin;
    tgb_pinhole_init(p_cam, &pin);

    TGB_CUDA(cudaEventRecord(d->ev[2], d->stream));
    k_set_words<<<1, 32, 0, d->stream>>>(d->d_visible_count, 4, 0u);
    TGB_LAUNCH_CHECK(d);
    k_cull_objects<<<(object_capacity + 127) / 128, 128, 0, d->stream>>>(d->d_objects, object_capacity, *p_cam, pin, d->width, d->height, d->d_frames, d->d_visible_count);
    TGB_LAUNCH_CHECK(d);
    k_sort_frames<<<1, 1024, 0, d->stream>>>(d->d_frames, d->d_frames_sorted, d->d_visible_count);
    TGB_LAUNCH_CHECK(d);
    TGB_CUDA(cudaEventRecord(d->ev[3], d->stream));

    /* TGB_K1_KERNEL: 1 (default) = one pixel per lane; 2 = several pixels per lane, walk state in shared memory (tgb_visibility_pool.cu):
     * bit-identical, measured SLOWER (c2: 1.06 ms with 2 pixels per lane, 0.92 with 1, against 0.73 here) -- the march's trip-count variance
     * inside a phase, not the number of lanes that start it, is what idles the lanes (profiles/r02d_k1pool_*); kept as the measured record */
    const int k1_kernel = tgbd_env_int("TGB_K1_KERNEL", 1);
    if (k1_kernel == 2 && !tgbd_k1_pool_render(d, p_cam)) return TG_FALSE;
    const u32 fallback_only = (k1_kernel == 2 ? 1u : 0u) | (tgbd_env_int("TGB_K1_EARLY_EXIT", 1) ? 0u : 2u); /* bit 0, after the pool kernel: only frames it declined (an object too large for its packed iterator) */
    const dim3 grid((d->width + TGB_TILE_W - 1) / TGB_TILE_W, (d->height + TGB_TILE_H - 1) / TGB_TILE_H);
    /* register budget: 5 CTAs per SM = 48 registers + ~30 spilled words; measured 0.711 ms against 0.737 with 4 CTAs (64 registers), 0.715 with 6 (40), 0.843 with 3
     * (80): the kernel is issue-bound and more resident warps buy more than the spills cost (TGB_K1_MIN_CTAS selects the other builds; tuning only) */
    const int min_ctas = tgbd_env_int("TGB_K1_MIN_CTAS", 5), regroup = tgbd_env_int("TGB_K1_REGROUP", 1);
    /* sharded frame on the peer-memory path: K1 flags the tiles in which this rank has a hit */
    const bool sharded = d->p2p_ready && d->n_ranks > 1 && k1_kernel != 2;
    tgb_k1_shard_args sh;
    memset(&sh, 0, sizeof(sh));
    if (sharded)
    {
        sh.p_tile_flags = tgbd_mat_tile_flags(d, d->d_mat);
        sh.tiles_x = tgbd_tiles_x(d);
    }
#define TGB_K1_LAUNCH(C, R, S, D) k_visibility<C, R, S, D><<<grid, TGB_K1_THREADS, 0, d->stream>>>(d->d_frames_sorted, d->d_visible_count, *p_cam, d->width, d->height, \
                                                                                       d->d_cluster_pointers, d->d_masks, d->global_pointer_base, d->d_vis, d->n_ranks, d->tile_rows, fallback_only, sh)
    /* TGB_K1_DEFER_WORD=1: the word of a voxel found is computed at the end of the object by all lanes together instead of right after the march.
     * Bit-identical, measured SLOWER (c2: 0.80 ms against 0.72; c2far 1.12 against 0.98): the bound, the second origin evaluation and the candidates the
     * looser bound admits cost more than the idle lanes of the immediate form. Off; kept as the measured record. */
    const int defer = tgbd_env_int("TGB_K1_DEFER_WORD", 0);
    /* TGB_K1_STAGE_MASKS=1: cluster masks staged in shared memory by cp.async, one copy per group of lanes that march the same cluster
     * (north_star (a) as written). Bit-identical; measured against the default in profiles/r03p_*: the 8-byte slice loads it replaces already hit L1
     * 96 % of the time, so the copy's latency (waited for before the march can start) is all it adds. Off; kept as the measured record. */
    if (tgbd_env_int("TGB_K1_STAGE_MASKS", 0) && !sharded && k1_kernel != 2)
        k_visibility<5, true, false, false, true><<<grid, TGB_K1_THREADS, 0, d->stream>>>(d->d_frames_sorted, d->d_visible_count, *p_cam, d->width, d->height,
                                                                                           d->d_cluster_pointers, d->d_masks, d->global_pointer_base, d->d_vis, d->n_ranks, d->tile_rows, fallback_only, sh);
    else
    if (sharded)      { if (defer) TGB_K1_LAUNCH(5, true, true, true); else TGB_K1_LAUNCH(5, true, true, false); }
    else if (regroup)
    {
        if (defer) { if (min_ctas >= 6) TGB_K1_LAUNCH(6, true, false, true); else if (min_ctas == 5) TGB_K1_LAUNCH(5, true, false, true); else TGB_K1_LAUNCH(4, true, false, true); }
        else       { if (min_ctas >= 6) TGB_K1_LAUNCH(6, true, false, false); else if (min_ctas == 5) TGB_K1_LAUNCH(5, true, false, false); else if (min_ctas == 4) TGB_K1_LAUNCH(4, true, false, false); else TGB_K1_LAUNCH(3, true, false, false); }
    }
    else              { if (min_ctas >= 4) TGB_K1_LAUNCH(4, false, false, false); else TGB_K1_LAUNCH(3, false, false, false); }
#undef TGB_K1_LAUNCH
    TGB_LAUNCH_CHECK(d);
    d->tiles_flagged = sharded ? TG_TRUE : TG_FALSE;
    TGB_CUDA(cudaEventRecord(d->ev[4], d->stream));
    d->ev_vis = TG_TRUE;
    return TG_TRUE;
}

/*
 * tgb_k1_walk.cuh -- one primary ray against one object, as resumable pieces.
 *
 *   fragment     assets/shaders/raytracer/visibility.frag:22-208 (collide.inc:3-24 slab test, :83-191 8^3 Amanatides-Woo,
 *                :194-206 depth quantisation, packing, atomicMin)
 * The reference evaluates that fragment for every (pixel, cluster) pair the rasteriser covers; the result per pixel is the
 * 64-bit minimum over ALL clusters. Here a ray walks the cluster grid of an object slice by slice along its
 * dominant axis, visiting a CONSERVATIVE SUPERSET of the clusters whose own slab test can succeed, front to back. For each
 * visited cluster the arithmetic that produces the written word is the reference's, operation for operation (tgb_hoist.h,
 * tgb_math.h, no FMA contraction); everything that merely selects candidates may be approximate because a superset yields the
 * identical minimum. Early-outs compare quantised 24-bit depths and skip only on STRICTLY greater (a tie must still run: the
 * lower pointer / voxel wins).
 *
 * The walk is cut into three pieces so that a kernel can suspend a ray between them (k_visibility_pool, tgb_visibility.cu:
 * several rays per lane, state in shared memory, each warp iteration runs the piece most lanes are waiting for):
 *   tgb_k1_setup           ray vs object: cheap reject, exact cluster-space direction, slab of the whole grid, first slice
 *   tgb_k1_next_candidate  advance the (slice, v, u) iterator to the next cluster that passes the first half of the fragment
 *   tgb_cluster_march      second half of the fragment: the 8^3 march, depth, packed word, best / t_skip update
 * Host-compilable: tests/cpu_sim drives the same pieces pixel by pixel and compares the words with the oracle.
 */
#ifndef TGB_K1_WALK_CUH
#define TGB_K1_WALK_CUH

#include "tgb_math.h"
#include "tgb_hoist.h"
#include "tgb_gi_walk.cuh" /* TGB_LDG / TGB_RCP_RN / TGB_FDIVIDEF */

/* one z-slice of a cluster mask: 64 bits, bit 8y + x (one aligned 8-byte load) */
#ifdef __CUDACC__
typedef uint2 tgb_slice;
#else
struct tgb_slice { u32 x, y; };
#endif

/* visibility.frag:194-201 quantisation of a depth in [0,1] (or beyond) */
TGB_HD u64 tgb_depth24(f32 t, f32 far_plane)
{
    const f32 dq = tgb_max(0.0f, t / far_plane) * TG_VIS_DEPTH_SCALE;
    return (u64)dq; /* cvt.rzi.u64.f32: truncation, saturating, NaN -> 0 */
}

/*
 * What one ray keeps per object: the exact cluster-space direction d (shared by all clusters of the object,
 * tgb_hoist.h), the DDA increments 1 / |d| (visibility.frag:105-136; rcp.rn is the IEEE quotient 1 / x) and their signed
 * twins r = 1 / d, which double as APPROXIMATE reciprocals: n * r is within 2^-22 of the IEEE quotient n / d, so it can
 * rank slab quotients and decide clear-cut comparisons, and an IEEE division is spent only on the value that is kept.
 * `exotic` (a non-zero component below 1e-30, whose reciprocal overflows) switches every short cut off.
 */
struct tgb_ray_in_object
{
    v3  d;
    f32 t_delta_x, t_delta_y, t_delta_z;
    f32 rx, ry, rz;
    bool exotic;
};

TGB_HD void tgb_ray_in_object_init(tgb_ray_in_object* r, v3 d)
{
    r->d = d;
    const f32 adx = fabsf(d.x), ady = fabsf(d.y), adz = fabsf(d.z);
    /* visibility.frag:105-136: t_delta = 1 / d or 1 / -d, absent axis F32_MAX */
    r->t_delta_x = adx != 0.0f ? TGB_RCP_RN(adx) : TG_F32_MAX;
    r->t_delta_y = ady != 0.0f ? TGB_RCP_RN(ady) : TG_F32_MAX;
    r->t_delta_z = adz != 0.0f ? TGB_RCP_RN(adz) : TG_F32_MAX;
    r->rx = d.x < 0.0f ? -r->t_delta_x : r->t_delta_x;
    r->ry = d.y < 0.0f ? -r->t_delta_y : r->t_delta_y;
    r->rz = d.z < 0.0f ? -r->t_delta_z : r->t_delta_z;
    r->exotic = (adx != 0.0f && adx < 1e-30f) || (ady != 0.0f && ady < 1e-30f) || (adz != 0.0f && adz < 1e-30f);
}

/* the same from stored d and 1 / |d| (no reciprocal is recomputed when a suspended ray is resumed) */
TGB_HD void tgb_ray_in_object_restore(tgb_ray_in_object* r, v3 d, f32 t_delta_x, f32 t_delta_y, f32 t_delta_z, bool exotic)
{
    r->d = d;
    r->t_delta_x = t_delta_x; r->t_delta_y = t_delta_y; r->t_delta_z = t_delta_z;
    r->rx = d.x < 0.0f ? -t_delta_x : t_delta_x;
    r->ry = d.y < 0.0f ? -t_delta_y : t_delta_y;
    r->rz = d.z < 0.0f ? -t_delta_z : t_delta_z;
    r->exotic = exotic;
}

/*
 * `enter` of collide.inc:3-24 for the box [lo, lo + size]^3 when the ray is already known to meet the box: the largest
 * of the three near-plane quotients. min((lo - o) / d, (hi - o) / d) is the quotient of the plane the ray meets first
 * (IEEE division by d is monotone), a zero component contributes -F32_MAX, and an axis whose approximate quotient is
 * clearly below the largest one cannot be the maximum (rounding is monotone), so only the axes within 1e-5 of it are
 * divided -- almost always one.
 */
TGB_HD f32 tgb_slab_enter(const tgb_ray_in_object& r, f32 nx, f32 ny, f32 nz, f32 ex, f32 ey, f32 ez, f32 e_max)
{
    const f32 floor_e = e_max - (1e-5f * fabsf(e_max) + 1e-30f);
    const bool cx = ex >= floor_e, cy = ey >= floor_e, cz = ez >= floor_e;
    const f32 num = cx ? nx : (cy ? ny : nz), den = cx ? r.d.x : (cy ? r.d.y : r.d.z);
    f32 enter = num / den;
    if ((u32)cx + (u32)cy + (u32)cz != 1u)
    {
        enter = TG_F32_MIN;
        if (cx && r.d.x != 0.0f) enter = tgb_max(enter, nx / r.d.x);
        if (cy && r.d.y != 0.0f) enter = tgb_max(enter, ny / r.d.y);
        if (cz && r.d.z != 0.0f) enter = tgb_max(enter, nz / r.d.z);
    }
    return enter;
}

/*
 * One cluster, exactly visibility.frag:71-207 with the ray (o, d) in cluster space, in two halves.
 *
 * First half, visibility.frag:71-81 = collide.inc:3-24 against [0,8]^3: does the ray meet the cluster (exit > 0 &&
 * enter <= exit), and can a hit inside still beat `best`? `t_skip` is a ray parameter beyond which no hit can
 * (depth24(t) > depth24(best) for every t >= t_skip). Returns the shader's `enter`.
 */
TGB_HD bool tgb_cluster_candidate(const tgb_object_frame& f, const tgb_ray_in_object& r, u32 cx, u32 cy, u32 cz, f32 t_skip, f32* p_enter)
{
    const v3 o = tgb_hoist_cluster_origin(&f, cx, cy, cz);
    const v3 d = r.d;
    const f32 nx = (d.x > 0.0f ? 0.0f : 8.0f) - o.x, fx = (d.x > 0.0f ? 8.0f : 0.0f) - o.x;
    const f32 ny = (d.y > 0.0f ? 0.0f : 8.0f) - o.y, fy = (d.y > 0.0f ? 8.0f : 0.0f) - o.y;
    const f32 nz = (d.z > 0.0f ? 0.0f : 8.0f) - o.z, fz = (d.z > 0.0f ? 8.0f : 0.0f) - o.z;
    const f32 ex = d.x != 0.0f ? nx * r.rx : TG_F32_MIN, xx = d.x != 0.0f ? fx * r.rx : TG_F32_MAX;
    const f32 ey = d.y != 0.0f ? ny * r.ry : TG_F32_MIN, xy = d.y != 0.0f ? fy * r.ry : TG_F32_MAX;
    const f32 ez = d.z != 0.0f ? nz * r.rz : TG_F32_MIN, xz = d.z != 0.0f ? fz * r.rz : TG_F32_MAX;
    const f32 e_max = fmaxf(fmaxf(ex, ey), ez), x_min = fminf(fminf(xx, xy), xz);
    const f32 tol = 1e-5f * (fabsf(e_max) + fabsf(x_min)) + 1e-30f;
    f32 enter;
    if (!r.exotic && ((x_min > tol) & (e_max + tol < x_min)))
    {
        /* clear hit; a cluster entered beyond t_skip cannot win (voxel_enter >= enter, depth24 is monotone) */
        if (e_max - tol > t_skip) return false;
        enter = tgb_slab_enter(r, nx, ny, nz, ex, ey, ez, e_max);
    }
    else
    {
        if (!r.exotic && ((x_min < -tol) | (e_max - tol > x_min))) return false; /* clear miss */
        f32 exit;
        if (!tgb_ray_aabb(o, d, tgb_v3(0.0f, 0.0f, 0.0f), tgb_v3(8.0f, 8.0f, 8.0f), &enter, &exit)) return false;
    }
    /* voxel_enter >= enter (same o, d, nested boxes, monotone rounding) => depth24(hit) >= depth24(enter) > depth24(best) */
    if (enter > t_skip) return false;
    *p_enter = enter;
    return true;
}

/*
 * Second half, visibility.frag:83-191: the 8^3 Amanatides-Woo march from `enter`. Returns the index 64 z + 8 y + x of the first solid
 * voxel the shader's DDA meets, or -1 when the ray leaves the cluster without meeting one.
 */
/* STAGED: p_slices points at a copy of the mask in shared memory (k_visibility's TGB_K1_STAGE_MASKS variant), read with plain loads */
template <bool STAGED>
TGB_HD i32 tgb_cluster_find_in(const tgb_object_frame& f, const tgb_ray_in_object& r, u32 cx, u32 cy, u32 cz, f32 enter, const tgb_slice* p_slices)
{
    const v3 o = tgb_hoist_cluster_origin(&f, cx, cy, cz);
    const v3 d = r.d;

    /* visibility.frag:83-137 */
    v3 hit;
    if (enter > 0.0f) { hit.x = o.x + enter * d.x; hit.y = o.y + enter * d.y; hit.z = o.z + enter * d.z; }
    else              { hit = o; }
    i32 x = (i32)tgb_clamp(floorf(hit.x), 0.0f, 8.0f - 1.0f);
    i32 y = (i32)tgb_clamp(floorf(hit.y), 0.0f, 8.0f - 1.0f);
    i32 z = (i32)tgb_clamp(floorf(hit.z), 0.0f, 8.0f - 1.0f);

    i32 step_x = 0, step_y = 0, step_z = 0;
    f32 t_max_x = TG_F32_MAX, t_max_y = TG_F32_MAX, t_max_z = TG_F32_MAX;
    if (d.x > 0.0f)      { step_x = 1;  t_max_x = enter + ((f32)(x + 1) - hit.x) / d.x; }
    else if (d.x < 0.0f) { step_x = -1; t_max_x = enter + (hit.x - (f32)x) / -d.x; }
    if (d.y > 0.0f)      { step_y = 1;  t_max_y = enter + ((f32)(y + 1) - hit.y) / d.y; }
    else if (d.y < 0.0f) { step_y = -1; t_max_y = enter + (hit.y - (f32)y) / -d.y; }
    if (d.z > 0.0f)      { step_z = 1;  t_max_z = enter + ((f32)(z + 1) - hit.z) / d.z; }
    else if (d.z < 0.0f) { step_z = -1; t_max_z = enter + (hit.z - (f32)z) / -d.z; }

    /* visibility.frag:141-191; the 64-bit z-slice (words 2z, 2z+1) is fetched once per z */
    i32 z_cached = -1;
    u32 lo = 0, hi = 0;
    i32 found = -1;
    for (;;)
    {
        if (z != z_cached)
        {
            const tgb_slice s = STAGED ? p_slices[z] : TGB_LDG(&p_slices[z]);
            lo = s.x; hi = s.y; z_cached = z;
        }
        const u32 word = (y & 4) ? hi : lo;
        if ((word >> (((y & 3) << 3) + x)) & 1u) { found = 64 * z + 8 * y + x; break; }
        if (t_max_x < t_max_y)
        {
            if (t_max_x < t_max_z) { t_max_x += r.t_delta_x; x += step_x; if (x < 0 || x >= 8) break; }
            else                   { t_max_z += r.t_delta_z; z += step_z; if (z < 0 || z >= 8) break; }
        }
        else
        {
            if (t_max_y < t_max_z) { t_max_y += r.t_delta_y; y += step_y; if (y < 0 || y >= 8) break; }
            else                   { t_max_z += r.t_delta_z; z += step_z; if (z < 0 || z >= 8) break; }
        }
    }
    return found;
}

TGB_HD i32 tgb_cluster_find(const tgb_object_frame& f, const tgb_ray_in_object& r, u32 cx, u32 cy, u32 cz, f32 enter,
                            const u32* p_cluster_pointers, const u32* p_masks)
{
    const u32 cluster_pointer = f.first_cluster_pointer + cx + f.nx * (cy + f.ny * cz);
    const u32 cluster_idx = TGB_LDG(&p_cluster_pointers[cluster_pointer]);
    return tgb_cluster_find_in<false>(f, r, cx, cy, cz, enter, reinterpret_cast<const tgb_slice*>(p_masks + (u64)cluster_idx * TG_CLUSTER_MASK_WORDS));
}

/*
 * visibility.frag:151-157, 194-207 for the voxel tgb_cluster_find returned: depth from the slab test against the voxel (only its
 * `enter` is used), quantisation, the packed word, best / t_skip update.
 */
TGB_HD void tgb_cluster_word(const tgb_object_frame& f, const tgb_ray_in_object& r, u32 cx, u32 cy, u32 cz, i32 voxel, f32 far_plane,
                             u32 global_pointer_base, u64& best, f32& t_skip)
{
    const v3 o = tgb_hoist_cluster_origin(&f, cx, cy, cz);
    const v3 d = r.d;
    const u32 cluster_pointer = f.first_cluster_pointer + cx + f.nx * (cy + f.ny * cz);
    const i32 x = voxel & 7, y = (voxel >> 3) & 7, z = voxel >> 6;
    f32 voxel_enter;
    {
        const f32 vx = (f32)(d.x > 0.0f ? x : x + 1) - o.x, vy = (f32)(d.y > 0.0f ? y : y + 1) - o.y, vz = (f32)(d.z > 0.0f ? z : z + 1) - o.z;
        if (!r.exotic)
        {
            const f32 qx = d.x != 0.0f ? vx * r.rx : TG_F32_MIN, qy = d.y != 0.0f ? vy * r.ry : TG_F32_MIN, qz = d.z != 0.0f ? vz * r.rz : TG_F32_MIN;
            voxel_enter = tgb_slab_enter(r, vx, vy, vz, qx, qy, qz, fmaxf(fmaxf(qx, qy), qz));
        }
        else
        {
            f32 voxel_exit;
            tgb_ray_aabb(o, d, tgb_v3((f32)x, (f32)y, (f32)z), tgb_v3((f32)(x + 1), (f32)(y + 1), (f32)(z + 1)), &voxel_enter, &voxel_exit);
        }
    }
    const f32 depth = tgb_max(0.0f, voxel_enter / far_plane);
    if (depth <= 1.0f)
    {
        const f32 dq = depth * TG_VIS_DEPTH_SCALE;
        const u64 word = ((u64)dq << TG_VIS_DEPTH_SHIFT)
                       | ((u64)(cluster_pointer + global_pointer_base) << TG_VIS_POINTER_SHIFT)
                       | (u64)(u32)voxel;
        if (word < best)
        {
            best = word;
            /* t / far * 16777215 >= trunc(dq) + 1 puts depth24(t) above the best depth: t_skip with a 1e-5 relative cushion */
            t_skip = (truncf(dq) + 1.0f) * (far_plane * (1.00001f / TG_VIS_DEPTH_SCALE));
        }
    }
}

/*
 * An UPPER bound of the t_skip tgb_cluster_word would leave for this voxel, from approximate quotients only (no division): lets
 * the walk go on -- and stop early -- while the exact word is computed later, together with the other lanes' (k_visibility).
 * The approximate near-plane quotients are within 2e-7 of the exact ones, so inflating the largest by 1e-6 bounds the exact
 * `enter`; the quantised depth of that bound, plus 2 instead of 1 (one unit for the rounding of this evaluation), times the same
 * cushioned scale is never below (trunc(dq) + 1) * far * 1.00001 / scale. A bound that is too LARGE only lets more candidates
 * through (a superset yields the same minimum); `exotic` rays get no bound.
 */
TGB_HD f32 tgb_cluster_t_skip_bound(const tgb_object_frame& f, const tgb_ray_in_object& r, u32 cx, u32 cy, u32 cz, i32 voxel, f32 far_plane)
{
    if (r.exotic) return TG_F32_MAX;
    const v3 o = tgb_hoist_cluster_origin(&f, cx, cy, cz);
    const v3 d = r.d;
    const i32 x = voxel & 7, y = (voxel >> 3) & 7, z = voxel >> 6;
    const f32 vx = (f32)(d.x > 0.0f ? x : x + 1) - o.x, vy = (f32)(d.y > 0.0f ? y : y + 1) - o.y, vz = (f32)(d.z > 0.0f ? z : z + 1) - o.z;
    const f32 qx = d.x != 0.0f ? vx * r.rx : TG_F32_MIN, qy = d.y != 0.0f ? vy * r.ry : TG_F32_MIN, qz = d.z != 0.0f ? vz * r.rz : TG_F32_MIN;
    const f32 q = fmaxf(fmaxf(qx, qy), qz);
    const f32 q_up = q + (1e-6f * fabsf(q) + 1e-30f);
    const f32 dq_up = fmaxf(0.0f, q_up / far_plane) * TG_VIS_DEPTH_SCALE;
    if (!(dq_up < 3.0e7f)) return TG_F32_MAX; /* beyond the far plane (or NaN): no bound */
    return (floorf(dq_up) + 2.0f) * (far_plane * (1.00001f / TG_VIS_DEPTH_SCALE));
}

/* both halves in one go */
TGB_HD void tgb_cluster_march(const tgb_object_frame& f, const tgb_ray_in_object& r, u32 cx, u32 cy, u32 cz, f32 enter, f32 far_plane,
                              const u32* p_cluster_pointers, const u32* p_masks, u32 global_pointer_base, u64& best, f32& t_skip)
{
    const i32 voxel = tgb_cluster_find(f, r, cx, cy, cz, enter, p_cluster_pointers, p_masks);
    if (voxel >= 0) tgb_cluster_word(f, r, cx, cy, cz, voxel, far_plane, global_pointer_base, best, t_skip);
}

/* ---- the walk over one object's cluster grid, resumable ------------------------------------------------------------------ */

/*
 * Iterator state of one (ray, object) pair. `axis` = dominant axis k of d (the slices are perpendicular to it); u, v are the two
 * other axes in cyclic order. The range fields are empty (cu1 < cu, cv1 < cv) until the first slice is opened.
 */
struct tgb_k1_walk
{
    v3  d;                                  /* exact cluster-space direction */
    f32 t_delta_x, t_delta_y, t_delta_z;    /* 1 / |d| */
    f32 t_in, t_out;                        /* conservative slab of the inflated grid box along the ray */
    i32 s, n_slices;                        /* current slice along the dominant axis, slices still to open */
    i32 cu, cu0, cu1, cv, cv1;              /* (u, v) cluster ranges of the current slice and the cursor in them */
    u32 axis, negative, exotic;             /* dominant axis; d[axis] < 0; see tgb_ray_in_object */
};

/* the three permuted views of a per-axis triple */
#define TGB_K1_K(a, x, y, z) ((a) == 0 ? (x) : ((a) == 1 ? (y) : (z)))
#define TGB_K1_U(a, x, y, z) ((a) == 0 ? (y) : ((a) == 1 ? (z) : (x)))
#define TGB_K1_V(a, x, y, z) ((a) == 0 ? (z) : ((a) == 1 ? (x) : (y)))

/*
 * Ray vs object. Returns false when the ray cannot touch the (inflated) grid box; otherwise the walk is positioned before the
 * first slice. Cheap reject first (slab test of the box with the un-normalised direction and approximate reciprocals; the box
 * is inflated by 2 eps plus 1e-4 of the magnitudes involved, three orders above the error of the approximate quotients and of
 * either path), then the exact set-up: normalised direction (tgb_hoist_direction), 1 / |d|, dominant axis, conservative slab.
 */
TGB_HD bool tgb_k1_setup(const tgb_object_frame& f, v3 dir_ws, tgb_k1_walk* w)
{
    const f32 e = f.eps;
    /* ws2ms * (dir_ws, 0) before normalisation (tgb_hoist_direction) */
    v3 raw;
    raw.x = (dir_ws.x * f.c[0] + dir_ws.y * f.c[1]) + dir_ws.z * f.c[2];
    raw.y = (dir_ws.x * f.c[3] + dir_ws.y * f.c[4]) + dir_ws.z * f.c[5];
    raw.z = (dir_ws.x * f.c[6] + dir_ws.y * f.c[7]) + dir_ws.z * f.c[8];
    {
        const f32 ex = 8.0f * (f32)f.nx, ey = 8.0f * (f32)f.ny, ez = 8.0f * (f32)f.nz;
        const f32 m = 2.0f * e + 1e-4f * (fabsf(f.og[0]) + fabsf(f.og[1]) + fabsf(f.og[2]) + ex + ey + ez);
        f32 t0 = 0.0f, t1 = TG_F32_MAX;
        bool out = false;
        if (fabsf(raw.x) > 1e-20f) { const f32 i = TGB_FDIVIDEF(1.0f, raw.x), a = (-m - f.og[0]) * i, b2 = (ex + m - f.og[0]) * i; t0 = fmaxf(t0, fminf(a, b2)); t1 = fminf(t1, fmaxf(a, b2)); }
        else out = out || f.og[0] < -m || f.og[0] > ex + m;
        if (fabsf(raw.y) > 1e-20f) { const f32 i = TGB_FDIVIDEF(1.0f, raw.y), a = (-m - f.og[1]) * i, b2 = (ey + m - f.og[1]) * i; t0 = fmaxf(t0, fminf(a, b2)); t1 = fminf(t1, fmaxf(a, b2)); }
        else out = out || f.og[1] < -m || f.og[1] > ey + m;
        if (fabsf(raw.z) > 1e-20f) { const f32 i = TGB_FDIVIDEF(1.0f, raw.z), a = (-m - f.og[2]) * i, b2 = (ez + m - f.og[2]) * i; t0 = fmaxf(t0, fminf(a, b2)); t1 = fminf(t1, fmaxf(a, b2)); }
        else out = out || f.og[2] < -m || f.og[2] > ez + m;
        if (out || t0 > t1 * 1.0001f + 1e-30f) return false;
    }
    tgb_ray_in_object r;
    tgb_ray_in_object_init(&r, tgb_normalize(raw)); /* exact d_ms (tgb_hoist_direction), shared by all clusters of the object */
    const v3 d = r.d;
    const f32 adx = fabsf(d.x), ady = fabsf(d.y), adz = fabsf(d.z);

    /* permute so that axis k is the dominant one */
    const u32 k = (adx >= ady && adx >= adz) ? 0u : (ady >= adz ? 1u : 2u);
    const f32 dk = TGB_K1_K(k, d.x, d.y, d.z), du = TGB_K1_U(k, d.x, d.y, d.z), dv = TGB_K1_V(k, d.x, d.y, d.z);
    const f32 ok = TGB_K1_K(k, f.og[0], f.og[1], f.og[2]), ou = TGB_K1_U(k, f.og[0], f.og[1], f.og[2]), ov = TGB_K1_V(k, f.og[0], f.og[1], f.og[2]);
    const i32 nk = (i32)TGB_K1_K(k, f.nx, f.ny, f.nz), nu = (i32)TGB_K1_U(k, f.nx, f.ny, f.nz), nv = (i32)TGB_K1_V(k, f.nx, f.ny, f.nz);
    if (!(fabsf(dk) > 0.5f)) return false; /* |d| == 1 => dominant component >= 0.577; false only for NaN directions */

    /* conservative slab of the inflated object box; u / v slabs only when the ray is not parallel to them */
    const f32 inv_dk = TGB_K1_K(k, r.rx, r.ry, r.rz);
    f32 t_in, t_out;
    {
        const f32 ta = (-e - ok) * inv_dk, tb = (8.0f * (f32)nk + e - ok) * inv_dk;
        t_in = fminf(ta, tb); t_out = fmaxf(ta, tb);
    }
    if (fabsf(du) > 1e-20f)
    {
        const f32 inv = TGB_K1_U(k, r.rx, r.ry, r.rz);
        const f32 ta = (-e - ou) * inv, tb = (8.0f * (f32)nu + e - ou) * inv;
        t_in = fmaxf(t_in, fminf(ta, tb)); t_out = fminf(t_out, fmaxf(ta, tb));
    }
    else if (ou < -e || ou > 8.0f * (f32)nu + e) return false;
    if (fabsf(dv) > 1e-20f)
    {
        const f32 inv = TGB_K1_V(k, r.rx, r.ry, r.rz);
        const f32 ta = (-e - ov) * inv, tb = (8.0f * (f32)nv + e - ov) * inv;
        t_in = fmaxf(t_in, fminf(ta, tb)); t_out = fminf(t_out, fmaxf(ta, tb));
    }
    else if (ov < -e || ov > 8.0f * (f32)nv + e) return false;
    /* slack: relative 2^-16 of |t| plus eps (positions move by at most |t|*2^-16 + eps) */
    t_in  -= e + 1.52587890625e-5f * fabsf(t_in);
    t_out += e + 1.52587890625e-5f * fabsf(t_out);
    t_in = fmaxf(t_in, 0.0f);
    if (!(t_in <= t_out)) return false;

    const f32 pk_in = ok + t_in * dk, pk_out = ok + t_out * dk;
    const i32 sgn = dk > 0.0f ? 1 : -1;
    i32 s     = (i32)floorf((pk_in  - (f32)sgn * (2.0f * e)) * 0.125f);
    i32 s_end = (i32)floorf((pk_out + (f32)sgn * (2.0f * e)) * 0.125f);
    s     = s < 0 ? 0 : (s > nk - 1 ? nk - 1 : s);
    s_end = s_end < 0 ? 0 : (s_end > nk - 1 ? nk - 1 : s_end);

    w->d = d;
    w->t_delta_x = r.t_delta_x; w->t_delta_y = r.t_delta_y; w->t_delta_z = r.t_delta_z;
    w->t_in = t_in; w->t_out = t_out;
    w->n_slices = (s_end - s) * sgn + 1;
    w->s = s - sgn;
    w->cu = 0; w->cu0 = 0; w->cu1 = -1; w->cv = 0; w->cv1 = -1;
    w->axis = k; w->negative = dk > 0.0f ? 0u : 1u; w->exotic = r.exotic ? 1u : 0u;
    return true;
}

/*
 * Advances the walk to the ray's NEXT cluster that passes the first half of the fragment (tgb_cluster_candidate) and can still
 * beat the best word (t_skip). Returns false when the object is exhausted. On true the cluster is (cx, cy, cz), `enter` the
 * shader's entry parameter; the walk is left ON that cluster (the next call moves past it).
 */
TGB_HD bool tgb_k1_next_candidate(const tgb_object_frame& f, tgb_k1_walk* w, f32 t_skip, u32* p_cx, u32* p_cy, u32* p_cz, f32* p_enter)
{
    tgb_ray_in_object r;
    tgb_ray_in_object_restore(&r, w->d, w->t_delta_x, w->t_delta_y, w->t_delta_z, w->exotic != 0);
    const f32 e = f.eps;
    const u32 k = w->axis;
    const i32 sgn = w->negative ? -1 : 1;
    const f32 dk = TGB_K1_K(k, r.d.x, r.d.y, r.d.z), du = TGB_K1_U(k, r.d.x, r.d.y, r.d.z), dv = TGB_K1_V(k, r.d.x, r.d.y, r.d.z);
    const f32 ok = TGB_K1_K(k, f.og[0], f.og[1], f.og[2]), ou = TGB_K1_U(k, f.og[0], f.og[1], f.og[2]), ov = TGB_K1_V(k, f.og[0], f.og[1], f.og[2]);
    const i32 nu = (i32)TGB_K1_U(k, f.nx, f.ny, f.nz), nv = (i32)TGB_K1_V(k, f.nx, f.ny, f.nz);
    const f32 inv_dk = TGB_K1_K(k, r.rx, r.ry, r.rz);
    (void)dk;
    i32 s = w->s, n_slices = w->n_slices, cu = w->cu, cu0 = w->cu0, cu1 = w->cu1, cv = w->cv, cv1 = w->cv1;
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
            const f32 t0 = fmaxf(fminf(ta, tb), w->t_in), t1 = fminf(fmaxf(ta, tb), w->t_out);
            if (!(t0 <= t1)) continue;
            /* slices are visited with non-decreasing t0: once even the slice entry is behind the best hit, stop */
            if (t0 - (4.0f * e + 3.0517578125e-5f * fabsf(t0)) > t_skip) { n_slices = 0; break; }
            const f32 ua = ou + t0 * du, ub = ou + t1 * du;
            const f32 va = ov + t0 * dv, vb = ov + t1 * dv;
            const f32 pad = 2.0f * e + 3.0517578125e-5f * (fabsf(ou) + fabsf(ov) + t1);
            const i32 u0 = (i32)floorf((fminf(ua, ub) - pad) * 0.125f), u1r = (i32)floorf((fmaxf(ua, ub) + pad) * 0.125f);
            const i32 v0r = (i32)floorf((fminf(va, vb) - pad) * 0.125f), v1r = (i32)floorf((fmaxf(va, vb) + pad) * 0.125f);
            cu0 = u0 < 0 ? 0 : u0;
            const i32 u1 = u1r > nu - 1 ? nu - 1 : u1r;
            const i32 v0 = v0r < 0 ? 0 : v0r;
            const i32 v1 = v1r > nv - 1 ? nv - 1 : v1r;
            if (cu0 > u1 || v0 > v1) continue;
            cu = cu0; cu1 = u1; cv = v0; cv1 = v1;
        }
        const u32 cx = (u32)TGB_K1_K(k, s, cv, cu) , cy = (u32)TGB_K1_K(k, cu, s, cv), cz = (u32)TGB_K1_K(k, cv, cu, s);
        if (tgb_cluster_candidate(f, r, cx, cy, cz, t_skip, p_enter)) { *p_cx = cx; *p_cy = cy; *p_cz = cz; have = true; break; }
    }
    w->s = s; w->n_slices = n_slices; w->cu = cu; w->cu0 = cu0; w->cu1 = cu1; w->cv = cv; w->cv1 = cv1;
    return have;
}

/* the cluster the walk currently stands on (after tgb_k1_next_candidate returned true) */
TGB_HD void tgb_k1_current_cluster(const tgb_k1_walk* w, u32* p_cx, u32* p_cy, u32* p_cz)
{
    const u32 k = w->axis;
    *p_cx = (u32)TGB_K1_K(k, w->s, w->cv, w->cu);
    *p_cy = (u32)TGB_K1_K(k, w->cu, w->s, w->cv);
    *p_cz = (u32)TGB_K1_K(k, w->cv, w->cu, w->s);
}

/*
 * Conservative per-object data shared by the cull kernel and the walk: the camera in the object's grid frame and the
 * candidate-enumeration slack eps = 2^-5 + 2^-17 * magnitude (>> accumulated rounding of either path).
 */
TGB_HD void tgb_frame_conservative(tgb_object_frame* f, const tg_object_data* p_object, v3 camera)
{
    const v3 og = tgb_hoist_cluster_origin(f, 0, 0, 0);
    f->og[0] = og.x; f->og[1] = og.y; f->og[2] = og.z;
    const f32 ex = 8.0f * (f32)f->nx, ey = 8.0f * (f32)f->ny, ez = 8.0f * (f32)f->nz;
    const f32 mag = fmaxf(fmaxf(fabsf(p_object->translation.x), fabsf(p_object->translation.y)), fabsf(p_object->translation.z))
                  + fmaxf(fmaxf(fabsf(camera.x), fabsf(camera.y)), fabsf(camera.z))
                  + fmaxf(fmaxf(ex, ey), ez);
    f->eps = 0.03125f + 7.62939453125e-6f * mag;
}

#endif

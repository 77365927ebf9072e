/*
 * tgb_gi_walk.cuh -- the per-ray pieces of tg_svo_traverse (assets/shaders/raytracer/svo_functions.inc:1-329) over the
 * flattened tree, shared by the GI kernels (tgb_shade.cu: k_gi_trace, k_gi_trace_flat; tgb_gi_pool.cu: k_gi_trace_pool)
 * and written so that a plain C++ compiler can run them too: tests/cpu_sim/ drives the same state machine ray by ray on
 * the host (no CUDA needed) and compares its hit / miss decisions with the oracle's transcription of the shader.
 *
 * Why the flattened tree is exact is argued above k_gi_trace_flat (tgb_shade.cu); in short: the shader's stack only
 * serves to find, after every advance, the terminal node (invalid octant or leaf) around `position`, which is a
 * function of `position` and the shader's comparison rules alone and is tabulated per 32^3 cell by k_svo_flatten.
 * Everything that moves `position` or decides a comparison below is the shader's operation, in the shader's order
 * (all translation units are built without FMA contraction).
 */
#ifndef TGB_GI_WALK_CUH
#define TGB_GI_WALK_CUH

#include "tgb_math.h"

#ifdef __CUDA_ARCH__
#define TGB_LDG(p)          __ldg(p)
#define TGB_RCP_RN(x)       __frcp_rn(x)        /* == IEEE 1 / x */
#define TGB_FDIVIDEF(a, b)  __fdividef((a), (b)) /* 2-ulp quotient: ranks candidates only, never kept */
#else
#define TGB_LDG(p)          (*(p))
#define TGB_RCP_RN(x)       (1.0f / (x))
#define TGB_FDIVIDEF(a, b)  ((a) / (b))
#endif

/* the flattened tree (k_svo_flatten, tgb_svo.cu): one word per 32^3 cell of the 1024^3 box */
#define TGB_TOP_GRID_DIM      32u
#define TGB_TOP_GRID_CELLS    (TGB_TOP_GRID_DIM * TGB_TOP_GRID_DIM * TGB_TOP_GRID_DIM)
#define TGB_TOP_HAS_DATA      0x80000000u /* the terminal node is a leaf with n != 0: bits 0..27 = data_pointer */
#define TGB_TOP_LEVEL_SHIFT   28u         /* bits 28..30: depth of the inner node whose child is terminal (child side = 512 >> level) */
#define TGB_TOP_POINTER_MASK  0x0FFFFFFFu

/* 16-bit form of a cell: leaf with data -> 0x8000 | data pointer (< 32768 leaves, level 4), otherwise the level */
#define TGB_TOP16_HAS_DATA    0x8000u
TGB_HD u32 tgb_top16_pack(u32 entry) { return (entry & TGB_TOP_HAS_DATA) ? (TGB_TOP16_HAS_DATA | (entry & 0x7FFFu)) : ((entry >> TGB_TOP_LEVEL_SHIFT) & 7u); }
TGB_HD u32 tgb_top16_unpack(u32 e16) { return (e16 & TGB_TOP16_HAS_DATA) ? (TGB_TOP_HAS_DATA | (4u << TGB_TOP_LEVEL_SHIFT) | (e16 & 0x7FFFu)) : (e16 << TGB_TOP_LEVEL_SHIFT); }

#define TGB_TRAVERSE_MAX_ITERS 4096u /* Q9: cap that valid input never reaches */

/*
 * svo_functions.inc:283-292, the distance to the far border of a box: per axis the shader evaluates a = (min - p) / d and
 * b = (max - p) / d, takes max(a, b), then the minimum over the axes. Because min < max and IEEE subtraction / division
 * are monotone and sign-symmetric, max(a, b) is the quotient towards the FAR plane of the axis, q = num / |d| with
 * num = d > 0 ? max - p : p - min, bit for bit; d == 0 gives max(-F32_MAX, F32_MAX) = F32_MAX.
 * Rounding is monotone, so an axis whose quotient is clearly larger than the smallest one cannot be the minimum: a
 * 2-ulp approximate quotient ranks the axes and IEEE division is spent only on the axes within 1e-5 of the smallest
 * (almost always one). The value returned is the shader's.
 */
TGB_HD f32 tgb_exit_distance(v3 bmin, v3 bmax, v3 position, v3 d)
{
    const f32 nx = d.x > 0.0f ? bmax.x - position.x : position.x - bmin.x, ax = fabsf(d.x);
    const f32 ny = d.y > 0.0f ? bmax.y - position.y : position.y - bmin.y, ay = fabsf(d.y);
    const f32 nz = d.z > 0.0f ? bmax.z - position.z : position.z - bmin.z, az = fabsf(d.z);
    const f32 qx = ax != 0.0f ? TGB_FDIVIDEF(nx, ax) : TG_F32_MAX;
    const f32 qy = ay != 0.0f ? TGB_FDIVIDEF(ny, ay) : TG_F32_MAX;
    const f32 qz = az != 0.0f ? TGB_FDIVIDEF(nz, az) : TG_F32_MAX;
    const f32 q_min = fminf(fminf(qx, qy), qz);
    const f32 limit = q_min + (1e-5f * fabsf(q_min) + 1e-30f);
    /* common case, branch-free: exactly one axis is within the margin of the smallest quotient -> one IEEE division */
    const bool cx = qx <= limit, cy = qy <= limit, cz = qz <= limit;
    const f32 num = cx ? nx : (cy ? ny : nz), den = cx ? ax : (cy ? ay : az);
    f32 exit = num / den;
    /* __fdividef is only specified for 2^-126 <= |d| <= 2^126 and finite operands: anything unusual takes the exact path */
    const bool odd = !(q_min == q_min) || fabsf(q_min) > 1e30f || (ax != 0.0f && ax < 1e-30f) || (ay != 0.0f && ay < 1e-30f) || (az != 0.0f && az < 1e-30f);
    if (((u32)cx + (u32)cy + (u32)cz != 1u) || odd)
    {
        exit = TG_F32_MAX;
        if (ax != 0.0f && (odd || cx)) exit = tgb_min(exit, nx / ax);
        if (ay != 0.0f && (odd || cy)) exit = tgb_min(exit, ny / ay);
        if (az != 0.0f && (odd || cz)) exit = tgb_min(exit, nz / az);
    }
    return exit;
}

/*
 * exit_distance(box) > F32_EPSILON (the pop test, svo_functions.inc:296-324) without dividing: q = num / |d| exceeds
 * epsilon for sure when num > 2.5 eps |d| and is below it for sure when num < 0.5 eps |d| (this includes a ray that
 * is on or past the border, num <= 0); only the sliver in between needs the quotient itself. Branch-free unless a
 * component sits in the sliver.
 */
TGB_HD bool tgb_still_inside(v3 bmin, v3 bmax, v3 position, v3 d)
{
    const f32 nx = d.x > 0.0f ? bmax.x - position.x : position.x - bmin.x, ax = fabsf(d.x);
    const f32 ny = d.y > 0.0f ? bmax.y - position.y : position.y - bmin.y, ay = fabsf(d.y);
    const f32 nz = d.z > 0.0f ? bmax.z - position.z : position.z - bmin.z, az = fabsf(d.z);
    const bool in_x = (ax == 0.0f) | (nx > 2.5f * TG_F32_EPSILON * ax), out_x = (ax != 0.0f) & (nx < 0.5f * TG_F32_EPSILON * ax);
    const bool in_y = (ay == 0.0f) | (ny > 2.5f * TG_F32_EPSILON * ay), out_y = (ay != 0.0f) & (ny < 0.5f * TG_F32_EPSILON * ay);
    const bool in_z = (az == 0.0f) | (nz > 2.5f * TG_F32_EPSILON * az), out_z = (az != 0.0f) & (nz < 0.5f * TG_F32_EPSILON * az);
    if (in_x & in_y & in_z) return true;
    if (out_x | out_y | out_z) return false;
    return (in_x || nx / ax > TG_F32_EPSILON) && (in_y || ny / ay > TG_F32_EPSILON) && (in_z || nz / az > TG_F32_EPSILON);
}

/*
 * Index along one axis of the 32^3 cell the shader's octant rule (:63-80: upper half iff mid < p || (p == mid && d > 0))
 * selects for p. The box corners are multiples of 32 here, p / 32 is exact (a power of two) and so is its floor: the
 * cell is floor(p / 32) - min / 32, one lower when p sits exactly on a cell border and the ray does not move up; a
 * position outside the box takes the outermost cell like the shader's comparisons do.
 */
TGB_HD u32 tgb_cell_axis(f32 p, f32 d, i32 box_min_cell)
{
    const f32 q = p * 0.03125f, fl = floorf(q);
    const i32 c = (i32)fl - box_min_cell - (((q == fl) & !(d > 0.0f)) ? 1 : 0);
    return (u32)(c < 0 ? 0 : (c > 31 ? 31 : c));
}

/*
 * tgb_exit_distance with the ray's exact reciprocals 1 / |d| (the DDA increments, :139-176) as the approximate
 * quotients that rank the axes: num * RN(1 / |d|) is within 2^-22 of num / |d|, far inside the 1e-5 margin, and the
 * value returned is still the IEEE quotient of the winning axis (see tgb_exit_distance for why that is the shader's
 * value). `exotic` rays (a non-zero component below 1e-30, whose reciprocal overflows) always take the exact path.
 */
TGB_HD f32 tgb_exit_distance_rcp(v3 bmin, f32 size, v3 position, v3 d, f32 rx, f32 ry, f32 rz, bool exotic)
{
    const f32 nx = d.x > 0.0f ? (bmin.x + size) - position.x : position.x - bmin.x, ax = fabsf(d.x);
    const f32 ny = d.y > 0.0f ? (bmin.y + size) - position.y : position.y - bmin.y, ay = fabsf(d.y);
    const f32 nz = d.z > 0.0f ? (bmin.z + size) - position.z : position.z - bmin.z, az = fabsf(d.z);
    const f32 qx = ax != 0.0f ? nx * rx : TG_F32_MAX;
    const f32 qy = ay != 0.0f ? ny * ry : TG_F32_MAX;
    const f32 qz = az != 0.0f ? nz * rz : TG_F32_MAX;
    const f32 q_min = fminf(fminf(qx, qy), qz);
    const f32 limit = q_min + (1e-5f * fabsf(q_min) + 1e-30f);
    const bool cx = qx <= limit, cy = qy <= limit, cz = qz <= limit;
    const f32 num = cx ? nx : (cy ? ny : nz), den = cx ? ax : (cy ? ay : az);
    f32 exit = num / den;
    const bool odd = exotic | !(fabsf(q_min) < 1e30f);
    if (((u32)cx + (u32)cy + (u32)cz != 1u) | odd)
    {
        exit = TG_F32_MAX;
        if (ax != 0.0f && (odd || cx)) exit = tgb_min(exit, nx / ax);
        if (ay != 0.0f && (odd || cy)) exit = tgb_min(exit, ny / ay);
        if (az != 0.0f && (odd || cz)) exit = tgb_min(exit, nz / az);
    }
    return exit;
}

/* ---- the traversal as a resumable per-ray state machine (k_gi_trace_pool) --------------------------------------- */

/* what a ray waits for */
enum { TGB_RAY_IDLE = 0, TGB_RAY_TREE = 1, TGB_RAY_DDA = 2, TGB_RAY_HIT = 3, TGB_RAY_MISS = 4 };

/* per-frame constants of the traversal */
struct tgb_gi_frame
{
    v3  bmin, bmax;   /* the SVO box (corners on the 32-unit lattice) */
    v3  center;       /* svo_functions.inc:3-8: the ray origin is taken relative to it */
    v3  box_mid;
    i32 min_cell_x, min_cell_y, min_cell_z;
    f32 far_plane;
    const u32* p_grid;   /* [32^3 + 1] flattened tree */
    const unsigned short* p_grid16; /* [32^3] the same in 16 bits per cell (tgb_top16_*): half the cache footprint; global or shared memory */
    const u32* p_voxels; /* 1024 u32 per leaf */
};

TGB_HD void tgb_gi_frame_init(tgb_gi_frame* f, v3 bmin, v3 bmax, f32 far_plane, const u32* p_grid, const u32* p_voxels)
{
    f->p_grid16 = 0;
    f->bmin = bmin; f->bmax = bmax;
    const v3 extent = tgb_sub(bmax, bmin);
    f->center = tgb_add(tgb_scale(extent, 0.5f), bmin);
    f->box_mid = tgb_scale(tgb_add(bmin, bmax), 0.5f);
    f->min_cell_x = (i32)(bmin.x * 0.03125f); f->min_cell_y = (i32)(bmin.y * 0.03125f); f->min_cell_z = (i32)(bmin.z * 0.03125f);
    f->far_plane = far_plane;
    f->p_grid = p_grid; f->p_voxels = p_voxels;
}

/*
 * The state of one ray between phases, 16 words (one shared-memory column per word in k_gi_trace_pool):
 *   d, position, t_delta (1 / |d|, :139-176), t_max (leaf DDA), and four packed words:
 *   cell  = cx | cy << 5 | cz << 10 | level << 15 | iterations << 18   the terminal box of the last look-up (side 512 >> level,
 *           corner = cell coordinates with the low bits cleared) and the shader's iteration counter
 *   vox   = x | y << 5 | z << 10 | flags << 16                         voxel inside the leaf block + flags below
 *   data  = leaf data pointer (block = p_voxels + 1024 * data)
 *   slot  = index in the ray queue
 */
#define TGB_RF_ADVANCE  0x01u /* the next tree phase first advances to the far border of the terminal box (:279-294) */
#define TGB_RF_SETUP    0x02u /* the next DDA phase first sets the leaf DDA up (:111-176) */
#define TGB_RF_BORDER   0x04u /* MISS only: within one unit of a root face, the pop test of the root is still to be made */
#define TGB_RF_EXOTIC   0x08u /* a non-zero direction component below 1e-30 */
#define TGB_RF_STEP_SHIFT 4u  /* 2 bits per axis: 0 -> 0, 1 -> +1, 2 -> -1 */

TGB_HD u32 tgb_step_code(f32 d) { return d > 0.0f ? 1u : (d < 0.0f ? 2u : 0u); }
TGB_HD i32 tgb_step_decode(u32 code) { return (i32)(code & 1u) - (i32)(code >> 1); }

TGB_HD void tgb_cell_box(const tgb_gi_frame* f, u32 cell, v3* p_child_min, f32* p_child_size)
{
    const u32 cx = cell & 31u, cy = (cell >> 5) & 31u, cz = (cell >> 10) & 31u, level = (cell >> 15) & 7u;
    const u32 cells = 16u >> level;  /* side of the terminal box in cells */
    const u32 keep = ~(cells - 1u);
    *p_child_size = (f32)(cells << 5);
    *p_child_min = tgb_v3(f->bmin.x + (f32)((cx & keep) << 5), f->bmin.y + (f32)((cy & keep) << 5), f->bmin.z + (f32)((cz & keep) << 5));
}

/* a fresh ray from its queue record: origin.xyz | direction.xyz + `enter` of the slab test against the root (:27-31, made by k_shade) */
TGB_HD void tgb_gi_ray_start(const tgb_gi_frame* f, v3 origin, v3 dir, f32 root_enter, v3* p_position, v3* p_t_delta, u32* p_flags)
{
    v3 position = tgb_sub(origin, f->center);
    if (root_enter > 0.0f) position = tgb_add(position, tgb_scale(dir, root_enter));
    *p_position = position;
    /* :139-176: the DDA increments 1 / |d| depend on the ray only (rcp.rn == IEEE 1 / x) */
    const f32 ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
    p_t_delta->x = ax != 0.0f ? TGB_RCP_RN(ax) : TG_F32_MAX;
    p_t_delta->y = ay != 0.0f ? TGB_RCP_RN(ay) : TG_F32_MAX;
    p_t_delta->z = az != 0.0f ? TGB_RCP_RN(az) : TG_F32_MAX;
    const bool exotic = (ax != 0.0f && ax < 1e-30f) || (ay != 0.0f && ay < 1e-30f) || (az != 0.0f && az < 1e-30f);
    *p_flags = (exotic ? TGB_RF_EXOTIC : 0u) | ((tgb_step_code(dir.x) | (tgb_step_code(dir.y) << 2) | (tgb_step_code(dir.z) << 4)) << TGB_RF_STEP_SHIFT);
}

/*
 * Tree phase: up to `reps` times { advance to the far border of the terminal box (:279-294), end the ray when it has left
 * the root (:296-324), look the terminal node around the new position up (:44-110) }. Returns the ray's new kind:
 * TREE (budget used up), DDA (arrived in a leaf with data: *p_data set, SETUP flagged) or MISS (left the root, or within
 * one unit of a root face: BORDER flagged, the exact pop test is made by tgb_gi_border_test).
 */
/* GRID: 0 = the 32-bit table, 1 = the 16-bit table through the read-only path, 2 = the 16-bit table with plain loads (staged in shared memory) */
template <int GRID>
TGB_HD u32 tgb_gi_tree_phase_t(const tgb_gi_frame* f, v3 d, v3 t_delta, v3* p_position, u32* p_cell, u32* p_flags, u32* p_data, u32 reps,
                               u32* p_n_visits, u32* p_n_advances)
{
    v3 position = *p_position;
    u32 flags = *p_flags, kind = TGB_RAY_TREE;
    u32 iterations = *p_cell >> 18;
    /* the terminal box is kept unpacked across the repetitions: decoded once on entry, packed once on exit */
    u32 cx = *p_cell & 31u, cy = (*p_cell >> 5) & 31u, cz = (*p_cell >> 10) & 31u, level = (*p_cell >> 15) & 7u;
    v3 child_min; f32 child_size;
    tgb_cell_box(f, *p_cell, &child_min, &child_size);
    const bool exotic = (flags & TGB_RF_EXOTIC) != 0;
    for (u32 rep = 0; rep < reps && kind == TGB_RAY_TREE; rep++)
    {
        if (flags & TGB_RF_ADVANCE)
        {
            (*p_n_advances)++;
            const f32 exit = tgb_exit_distance_rcp(child_min, child_size, position, d, t_delta.x, t_delta.y, t_delta.z, exotic);
            position = tgb_add(position, tgb_scale(d, exit + TG_F32_EPSILON));
            /* a position at least one unit inside every face passes the pop test (exit >= 1 / |d| >= ~1 > epsilon) without evaluating it */
            const f32 off = fmaxf(fmaxf(fabsf(position.x - f->box_mid.x), fabsf(position.y - f->box_mid.y)), fabsf(position.z - f->box_mid.z));
            if (!(off < 0.5f * (f32)TG_SVO_SIDE_LENGTH - 1.0f)) { kind = TGB_RAY_MISS; flags |= TGB_RF_BORDER; }
        }
        flags |= TGB_RF_ADVANCE;
        if (kind == TGB_RAY_TREE)
        {
            if (++iterations > TGB_TRAVERSE_MAX_ITERS) kind = TGB_RAY_MISS;
            else
            {
                (*p_n_visits)++;
                cx = tgb_cell_axis(position.x, d.x, f->min_cell_x);
                cy = tgb_cell_axis(position.y, d.y, f->min_cell_y);
                cz = tgb_cell_axis(position.z, d.z, f->min_cell_z);
                const u32 cell_idx = (cz << 10) | (cy << 5) | cx;
                const u32 entry = GRID == 0 ? TGB_LDG(&f->p_grid[cell_idx]) : tgb_top16_unpack(GRID == 1 ? (u32)TGB_LDG(&f->p_grid16[cell_idx]) : (u32)f->p_grid16[cell_idx]);
                level = (entry >> TGB_TOP_LEVEL_SHIFT) & 7u;
                const u32 cells = 16u >> level, keep = ~(cells - 1u);
                child_size = (f32)(cells << 5);
                child_min = tgb_v3(f->bmin.x + (f32)((cx & keep) << 5), f->bmin.y + (f32)((cy & keep) << 5), f->bmin.z + (f32)((cz & keep) << 5));
                if (entry & TGB_TOP_HAS_DATA)
                {
                    *p_data = entry & TGB_TOP_POINTER_MASK;
                    flags |= TGB_RF_SETUP;
                    kind = TGB_RAY_DDA;
                }
            }
        }
    }
    *p_position = position;
    *p_cell = cx | (cy << 5) | (cz << 10) | (level << 15) | (iterations << 18);
    *p_flags = flags;
    return kind;
}

TGB_HD u32 tgb_gi_tree_phase(const tgb_gi_frame* f, v3 d, v3 t_delta, v3* p_position, u32* p_cell, u32* p_flags, u32* p_data, u32 reps,
                             u32* p_n_visits, u32* p_n_advances)
{
    return tgb_gi_tree_phase_t<0>(f, d, t_delta, p_position, p_cell, p_flags, p_data, reps, p_n_visits, p_n_advances);
}

/* MISS with BORDER flagged: the pop test of the root itself (:296-324); still inside -> back to the tree at the advanced position */
TGB_HD u32 tgb_gi_border_test(const tgb_gi_frame* f, v3 d, v3 position, u32* p_flags)
{
    *p_flags &= ~TGB_RF_BORDER;
    if (tgb_still_inside(f->bmin, f->bmax, position, d)) { *p_flags &= ~TGB_RF_ADVANCE; return TGB_RAY_TREE; }
    return TGB_RAY_MISS;
}

/* :111-176: voxel and t_max where the ray enters the leaf block */
TGB_HD void tgb_gi_dda_setup(v3 d, v3 position, v3 child_min, f32 child_size, i32* p_x, i32* p_y, i32* p_z, v3* p_t_max)
{
    v3 hit = position;
    v3 xyz = tgb_v3(tgb_clamp(floorf(hit.x), child_min.x, (child_min.x + child_size) - 1.0f),
                    tgb_clamp(floorf(hit.y), child_min.y, (child_min.y + child_size) - 1.0f),
                    tgb_clamp(floorf(hit.z), child_min.z, (child_min.z + child_size) - 1.0f));
    hit = tgb_sub(hit, child_min);
    xyz = tgb_sub(xyz, child_min);
    const i32 x = (i32)xyz.x, y = (i32)xyz.y, z = (i32)xyz.z;
    v3 t_max = tgb_v3(TG_F32_MAX, TG_F32_MAX, TG_F32_MAX);
    if (d.x > 0.0f)      t_max.x = ((f32)(x + 1) - hit.x) / d.x;
    else if (d.x < 0.0f) t_max.x = (hit.x - (f32)x) / -d.x;
    if (d.y > 0.0f)      t_max.y = ((f32)(y + 1) - hit.y) / d.y;
    else if (d.y < 0.0f) t_max.y = (hit.y - (f32)y) / -d.y;
    if (d.z > 0.0f)      t_max.z = ((f32)(z + 1) - hit.z) / d.z;
    else if (d.z < 0.0f) t_max.z = (hit.z - (f32)z) / -d.z;
    *p_x = x; *p_y = y; *p_z = z; *p_t_max = t_max;
}

/*
 * DDA phase: up to `steps` voxel steps of the leaf DDA (:178-257), written with selects (adding +0 to the other two t_max
 * leaves them bit-identical). A block row is one word: bit 1024 z + 32 y + x; it is re-read only when the row changes.
 * Returns DDA (budget used up), HIT (solid voxel at x, y, z: the shader's slab test decides, tgb_gi_hit_test) or TREE
 * (left the block; the ADVANCE flag is still set from the look-up, so the next tree phase moves past the leaf).
 */
/* STAGED: the block was copied to shared memory (k_gi_trace_list): plain loads */
template <bool STAGED>
TGB_HD u32 tgb_gi_dda_phase_t(const u32* p_block, v3 t_delta, u32 step_codes, v3* p_t_max, i32* p_x, i32* p_y, i32* p_z, u32 steps, u32* p_n_steps)
{
    const i32 step_x = tgb_step_decode(step_codes & 3u), step_y = tgb_step_decode((step_codes >> 2) & 3u), step_z = tgb_step_decode((step_codes >> 4) & 3u);
    f32 t_max_x = p_t_max->x, t_max_y = p_t_max->y, t_max_z = p_t_max->z;
    i32 x = *p_x, y = *p_y, z = *p_z;
    u32 kind = TGB_RAY_DDA;
    u32 bits = STAGED ? p_block[32 * z + y] : TGB_LDG(&p_block[32 * z + y]);
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (u32 k = 0; k < steps; k++)
    {
        (*p_n_steps)++;
        if ((bits >> x) & 1u) { kind = TGB_RAY_HIT; break; }
        const bool xy = t_max_x < t_max_y;
        const bool go_x = xy & (t_max_x < t_max_z);
        const bool go_y = !xy & (t_max_y < t_max_z);
        const bool go_z = !(go_x | go_y);
        t_max_x = go_x ? t_max_x + t_delta.x : t_max_x;
        t_max_y = go_y ? t_max_y + t_delta.y : t_max_y;
        t_max_z = go_z ? t_max_z + t_delta.z : t_max_z;
        x += go_x ? step_x : 0;
        y += go_y ? step_y : 0;
        z += go_z ? step_z : 0;
        if ((u32)(x | y | z) > 31u) { kind = TGB_RAY_TREE; break; } /* left the block: a coordinate is -1 or 32 */
        if (!go_x) bits = STAGED ? p_block[32 * z + y] : TGB_LDG(&p_block[32 * z + y]);
    }
    p_t_max->x = t_max_x; p_t_max->y = t_max_y; p_t_max->z = t_max_z;
    *p_x = x; *p_y = y; *p_z = z;
    return kind;
}

/*
 * The same DDA for a warp whose lanes all hold the SAME ray (k_gi_trace_list): branches are uniform there, so the step is written with
 * them -- one comparison chain, one addition, one coordinate, one bound -- a third of the instructions of the select form on the
 * dependent chain this kernel's duration consists of. Same comparisons, same additions: same t_max bits, same voxel sequence
 * (tests/test_gi_walk_cpu.py drives both forms side by side). The block is in shared memory.
 */
TGB_HD u32 tgb_gi_dda_phase_uniform(const u32* p_block, v3 t_delta, u32 step_codes, v3* p_t_max, i32* p_x, i32* p_y, i32* p_z, u32 steps, u32* p_n_steps)
{
    const i32 step_x = tgb_step_decode(step_codes & 3u), step_y = tgb_step_decode((step_codes >> 2) & 3u), step_z = tgb_step_decode((step_codes >> 4) & 3u);
    f32 t_max_x = p_t_max->x, t_max_y = p_t_max->y, t_max_z = p_t_max->z;
    i32 x = *p_x, y = *p_y, z = *p_z;
    u32 kind = TGB_RAY_DDA;
    u32 bits = p_block[32 * z + y];
    u32 k = 0;
    for (; k < steps; k++)
    {
        if ((bits >> x) & 1u) { kind = TGB_RAY_HIT; k++; break; }
        if (t_max_x < t_max_y)
        {
            if (t_max_x < t_max_z) { t_max_x += t_delta.x; x += step_x; if ((u32)x > 31u) { kind = TGB_RAY_TREE; k++; break; } continue; }
        }
        else if (t_max_y < t_max_z)
        {
            t_max_y += t_delta.y; y += step_y;
            if ((u32)y > 31u) { kind = TGB_RAY_TREE; k++; break; }
            bits = p_block[32 * z + y];
            continue;
        }
        t_max_z += t_delta.z; z += step_z;
        if ((u32)z > 31u) { kind = TGB_RAY_TREE; k++; break; }
        bits = p_block[32 * z + y];
    }
    *p_n_steps += k;
    p_t_max->x = t_max_x; p_t_max->y = t_max_y; p_t_max->z = t_max_z;
    *p_x = x; *p_y = y; *p_z = z;
    return kind;
}

TGB_HD u32 tgb_gi_dda_phase(const u32* p_block, v3 t_delta, u32 step_codes, v3* p_t_max, i32* p_x, i32* p_y, i32* p_z, u32 steps, u32* p_n_steps)
{
    return tgb_gi_dda_phase_t<false>(p_block, t_delta, step_codes, p_t_max, p_x, p_y, p_z, steps, p_n_steps);
}

/*
 * :219-256: result = enter / far of the slab test against the solid voxel. Only `enter` matters: the largest of the three
 * near-plane quotients, and min((lo - o) / d, (hi - o) / d) is the quotient of the plane the ray meets first (division by
 * d is monotone), so three divisions give the shader's value. Only result < 1 ends the shader's loop (occluded: IDLE),
 * otherwise it advances past the leaf (TREE).
 */
TGB_HD u32 tgb_gi_hit_test(const tgb_gi_frame* f, v3 origin, v3 d, v3 child_min, i32 x, i32 y, i32 z)
{
    const v3 o = tgb_sub(origin, f->center);
    const v3 lo = tgb_add(child_min, tgb_v3((f32)x, (f32)y, (f32)z));
    const v3 hi = tgb_add(child_min, tgb_v3((f32)(x + 1), (f32)(y + 1), (f32)(z + 1)));
    const f32 ex = d.x == 0.0f ? TG_F32_MIN : ((d.x > 0.0f ? lo.x : hi.x) - o.x) / d.x;
    const f32 ey = d.y == 0.0f ? TG_F32_MIN : ((d.y > 0.0f ? lo.y : hi.y) - o.y) / d.y;
    const f32 ez = d.z == 0.0f ? TG_F32_MIN : ((d.z > 0.0f ? lo.z : hi.z) - o.z) / d.z;
    const f32 enter = tgb_max(tgb_max(ex, ey), ez);
    return (enter / f->far_plane < 1.0f) ? TGB_RAY_IDLE : TGB_RAY_TREE;
}

#endif

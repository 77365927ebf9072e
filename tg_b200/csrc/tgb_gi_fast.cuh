/*
 * tgb_gi_fast.cuh -- the occluded / unoccluded decision of a secondary ray, CERTIFIED instead of transcribed.
 *
 *   SVO traversal      assets/shaders/raytracer/svo_functions.inc:1-329 (exact per-ray pieces: tgb_gi_walk.cuh)
 *
 * What the GI pass needs from tg_svo_traverse is one bit per ray: does the shader's traversal return a depth < 1
 * (occluded) or not. The exact kernels (k_gi_trace_pool, k_gi_trace_flat, k_gi_trace) obtain it by running the shader's own
 * arithmetic, operation for operation: an IEEE division per advanced cell, the shader's tie rules, its accumulated
 * `position` -- ~180 thread instructions per visited cell (profiles/r02b_gi_pool_regions.txt).
 *
 * The shader's traversal is a geometric one: it visits the voxels of the 1-bit SVO that the ray o + t d passes through, in
 * order, and stops at the first solid one (before the far plane). Its floating-point path differs from the ideal line only
 * by rounding: `position` is advanced by `position += (exit + eps) * d` at most a few dozen times at coordinates below 512
 * (half an ulp(512) = 3e-5 per advance and component; along-ray errors do not move the line), and inside a leaf the DDA's
 * t_max are sums of a few dozen small floats. In other words the shader decides like an exact traversal of a ray DISPLACED
 * sideways by less than ~3e-4 units. So a cheap traversal of the ideal line (FMA, reciprocal multiplies, no tie rules)
 * reaches the same decision whenever no decision it took depended on less than DELTA = 1e-3 units of displacement:
 *
 *   HIT   is certain when the ray stays inside the solid voxel for longer than 2 W (W = DELTA * sum 1 / |d_k|, the time
 *         a sideways displacement of DELTA can shift any plane crossing): every ray displaced by less than DELTA passes through
 *         that voxel too, so the shader's traversal meets it -- or an earlier solid voxel; either way it returns occluded.
 *         (Unless the voxel lies at the far plane: svo_functions.inc:219-256 only ends on enter / far < 1.)
 *   MISS  is certain when, in addition, no two consecutive plane crossings of the walk were closer than W in time (a displaced
 *         ray could have swapped them and visited a voxel this walk did not test) and the ray entered no box closer than
 *         DELTA-equivalent to a lattice line of that box's granularity (voxel planes for a leaf, 32-unit planes for an empty
 *         terminal box).
 *
 * Every ray that ends unoccluded with an uncertain event on its way, every ray with a direction component below 8 DELTA
 * (a ray can then linger next to a plane over several crossings) and every ray that runs into the iteration cap is handed
 * to the exact kernel (tgb_gi_pool.cu) through a list of queue slots. The decision taken here is therefore the shader's on every
 * ray; tests/test_gi_fast_cpu.py holds the host build of these functions against the exact state machine (itself held against
 * the oracle) on millions of rays, and sweeps DELTA down to where the first disagreement appears (tools/gi_fast_margin.py).
 * Host-compilable; the same IEEE operations on both sides (fmaf, 1 / x, floorf, rintf), so the host build predicts the
 * device's decisions AND its flags.
 */
#ifndef TGB_GI_FAST_CUH
#define TGB_GI_FAST_CUH

#include "tgb_gi_walk.cuh"

#define TGB_FAST_DELTA        1.0e-3f  /* sideways displacement (world units) a decision must survive */
#define TGB_FAST_SHALLOW      8.0e-3f  /* direction components below this (zero included) go to the exact kernel */
#define TGB_FAST_MAX_BOXES    200u     /* boxes a ray may enter (a straight line crosses < 96 cells); beyond: exact kernel */
#define TGB_FAST_FAR_FRACTION 0.99f    /* a solid voxel beyond this fraction of the far plane is not decided here */

/* kinds of the fast walk (TREE / DDA as in tgb_gi_walk.cuh) */
enum { TGB_FAST_IDLE = 0, TGB_FAST_TREE = 1, TGB_FAST_DDA = 2, TGB_FAST_OCCLUDED = 3, TGB_FAST_UNOCCLUDED = 4, TGB_FAST_EXACT = 5 };

struct tgb_fast_ray
{
    v3  o, d, inv;          /* origin relative to the box centre, direction, 1 / d */
    f32 w;                  /* W = DELTA * sum 1 / |d_k| (time units) */
    f32 t_cur;              /* ray parameter at which the current box was entered */
    v3  p;                  /* o + t_cur d */
    v3  t_max;              /* leaf DDA: time of the next plane crossing per axis, relative to t_cur */
    f32 m_cur;              /* leaf DDA: time the current voxel was entered, relative to t_cur */
    f32 w_leaf;             /* W scaled for the magnitude of t_cur */
    u32 cell;               /* cx | cy << 5 | cz << 10 of the 32^3 cell the ray stands in */
    u32 vox;                /* leaf DDA: x | y << 5 | z << 10 */
    u32 data;               /* leaf data pointer */
    u32 entry_axis;         /* axis through whose plane the current box was entered; bits 0..2 = axes NOT to check on entry */
    u32 uncertain;          /* an uncertain event happened on the way */
    u32 n_boxes;
};

/* distance of q to the nearest lattice plane of the given spacing (1 or 32) */
TGB_HD f32 tgb_fast_lattice_distance(f32 q, f32 spacing, f32 inv_spacing)
{
    const f32 s = q * inv_spacing;
    return fabsf(s - rintf(s)) * spacing;
}

/*
 * A fresh ray from its queue record (origin, direction, `enter` of the slab test against the root). Returns TREE, or EXACT for the
 * rays the fast walk does not take (a direction component below TGB_FAST_SHALLOW). `delta` is TGB_FAST_DELTA in the product; the
 * margin sweep (tools/gi_fast_margin.py) lowers it until the first disagreement with the exact walk appears.
 */
TGB_HD u32 tgb_fast_start(const tgb_gi_frame* f, v3 origin, v3 dir, f32 root_enter, f32 delta, tgb_fast_ray* r)
{
    const f32 ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
    if (!(ax >= TGB_FAST_SHALLOW && ay >= TGB_FAST_SHALLOW && az >= TGB_FAST_SHALLOW)) return TGB_FAST_EXACT; /* NaN included */
    r->o = tgb_sub(origin, f->center);
    r->d = dir;
    r->inv = tgb_v3(TGB_RCP_RN(dir.x), TGB_RCP_RN(dir.y), TGB_RCP_RN(dir.z));
    r->w = delta * ((fabsf(r->inv.x) + fabsf(r->inv.y)) + fabsf(r->inv.z));
    r->t_cur = root_enter > 0.0f ? root_enter : 0.0f;
    r->p = tgb_v3(fmaf(r->t_cur, dir.x, r->o.x), fmaf(r->t_cur, dir.y, r->o.y), fmaf(r->t_cur, dir.z, r->o.z));
    /* the cell around p; a ray that starts on (or outside) a root face is clamped into the outermost cell, and that axis has
     * nothing on its other side to be confused with */
    const f32 half = 0.5f * (f32)TG_SVO_SIDE_LENGTH;
    u32 skip = 0;
    skip |= fabsf(r->p.x - f->box_mid.x) > half - 0.01f ? 1u : 0u;
    skip |= fabsf(r->p.y - f->box_mid.y) > half - 0.01f ? 2u : 0u;
    skip |= fabsf(r->p.z - f->box_mid.z) > half - 0.01f ? 4u : 0u;
    const i32 cx = (i32)floorf((r->p.x - f->bmin.x) * 0.03125f), cy = (i32)floorf((r->p.y - f->bmin.y) * 0.03125f), cz = (i32)floorf((r->p.z - f->bmin.z) * 0.03125f);
    r->cell = (u32)(cx < 0 ? 0 : (cx > 31 ? 31 : cx)) | ((u32)(cy < 0 ? 0 : (cy > 31 ? 31 : cy)) << 5) | ((u32)(cz < 0 ? 0 : (cz > 31 ? 31 : cz)) << 10);
    r->entry_axis = skip;
    r->uncertain = 0;
    r->n_boxes = 0;
    return TGB_FAST_TREE;
}

/* is p, entering a box through the planes in `skip`, closer than W-equivalent to a lattice plane of the other axes? */
TGB_HD bool tgb_fast_entry_uncertain(const tgb_gi_frame* f, const tgb_fast_ray* r, f32 w, f32 spacing, f32 inv_spacing)
{
    const u32 skip = r->entry_axis;
    bool near = false;
    near = near || (!(skip & 1u) && tgb_fast_lattice_distance(r->p.x - f->bmin.x, spacing, inv_spacing) < w * fabsf(r->d.x));
    near = near || (!(skip & 2u) && tgb_fast_lattice_distance(r->p.y - f->bmin.y, spacing, inv_spacing) < w * fabsf(r->d.y));
    near = near || (!(skip & 4u) && tgb_fast_lattice_distance(r->p.z - f->bmin.z, spacing, inv_spacing) < w * fabsf(r->d.z));
    return near;
}

/*
 * Tree phase: up to `reps` boxes. Looks the terminal box around the ray's cell up; a leaf with data sets the DDA up (kind DDA),
 * an empty box is crossed to its far border. Returns TREE (budget used up), DDA, UNOCCLUDED (left the root) or EXACT (cap).
 */
TGB_HD u32 tgb_fast_tree_phase(const tgb_gi_frame* f, tgb_fast_ray* r, u32 reps, u32* p_n_visits)
{
    for (u32 rep = 0; rep < reps; rep++)
    {
        if (++r->n_boxes > TGB_FAST_MAX_BOXES) return TGB_FAST_EXACT;
        (*p_n_visits)++;
        const u32 cx = r->cell & 31u, cy = (r->cell >> 5) & 31u, cz = r->cell >> 10;
        const u32 entry = TGB_LDG(&f->p_grid[r->cell]);
        const f32 w = r->w * (1.0f + r->t_cur * 0.00390625f); /* the rounding of t grows with t */
        if (entry & TGB_TOP_HAS_DATA)
        {
            /* a leaf with data: voxel around p, next plane crossings relative to t_cur */
            if (tgb_fast_entry_uncertain(f, r, w, 1.0f, 1.0f)) r->uncertain = 1;
            const v3 lmin = tgb_v3(f->bmin.x + (f32)(cx << 5), f->bmin.y + (f32)(cy << 5), f->bmin.z + (f32)(cz << 5));
            const f32 hx = r->p.x - lmin.x, hy = r->p.y - lmin.y, hz = r->p.z - lmin.z;
            const f32 vx = tgb_clamp(floorf(hx), 0.0f, 31.0f), vy = tgb_clamp(floorf(hy), 0.0f, 31.0f), vz = tgb_clamp(floorf(hz), 0.0f, 31.0f);
            r->t_max.x = ((r->d.x > 0.0f ? vx + 1.0f : vx) - hx) * r->inv.x;
            r->t_max.y = ((r->d.y > 0.0f ? vy + 1.0f : vy) - hy) * r->inv.y;
            r->t_max.z = ((r->d.z > 0.0f ? vz + 1.0f : vz) - hz) * r->inv.z;
            r->vox = (u32)(i32)vx | ((u32)(i32)vy << 5) | ((u32)(i32)vz << 10);
            r->data = entry & TGB_TOP_POINTER_MASK;
            r->m_cur = 0.0f;
            r->w_leaf = w;
            return TGB_FAST_DDA;
        }
        /* an empty terminal box of 16 >> level cells: to its far border */
        if (tgb_fast_entry_uncertain(f, r, w, 32.0f, 0.03125f)) r->uncertain = 1;
        const u32 level = (entry >> TGB_TOP_LEVEL_SHIFT) & 7u;
        const u32 cells = 16u >> level, keep = ~(cells - 1u);
        const u32 bx = cx & keep, by = cy & keep, bz = cz & keep;
        const u32 fx = r->d.x > 0.0f ? bx + cells : bx, fy = r->d.y > 0.0f ? by + cells : by, fz = r->d.z > 0.0f ? bz + cells : bz;
        const f32 tx = ((f->bmin.x + (f32)(fx << 5)) - r->o.x) * r->inv.x;
        const f32 ty = ((f->bmin.y + (f32)(fy << 5)) - r->o.y) * r->inv.y;
        const f32 tz = ((f->bmin.z + (f32)(fz << 5)) - r->o.z) * r->inv.z;
        const bool xy = tx < ty;
        const bool go_x = xy & (tx < tz), go_y = !xy & (ty < tz), go_z = !(go_x | go_y);
        const f32 t_exit = go_x ? tx : (go_y ? ty : tz);
        r->t_cur = t_exit;
        r->p = tgb_v3(fmaf(t_exit, r->d.x, r->o.x), fmaf(t_exit, r->d.y, r->o.y), fmaf(t_exit, r->d.z, r->o.z));
        /* the next cell: exact along the exit axis, from p along the other two (clamped into the box: p is inside by construction) */
        i32 nx = (i32)floorf((r->p.x - f->bmin.x) * 0.03125f), ny = (i32)floorf((r->p.y - f->bmin.y) * 0.03125f), nz = (i32)floorf((r->p.z - f->bmin.z) * 0.03125f);
        nx = nx < (i32)bx ? (i32)bx : (nx > (i32)(bx + cells - 1u) ? (i32)(bx + cells - 1u) : nx);
        ny = ny < (i32)by ? (i32)by : (ny > (i32)(by + cells - 1u) ? (i32)(by + cells - 1u) : ny);
        nz = nz < (i32)bz ? (i32)bz : (nz > (i32)(bz + cells - 1u) ? (i32)(bz + cells - 1u) : nz);
        if (go_x) nx = r->d.x > 0.0f ? (i32)(bx + cells) : (i32)bx - 1;
        if (go_y) ny = r->d.y > 0.0f ? (i32)(by + cells) : (i32)by - 1;
        if (go_z) nz = r->d.z > 0.0f ? (i32)(bz + cells) : (i32)bz - 1;
        if ((u32)(nx | ny | nz) > 31u) return TGB_FAST_UNOCCLUDED; /* left the root */
        r->cell = (u32)nx | ((u32)ny << 5) | ((u32)nz << 10);
        r->entry_axis = go_x ? 1u : (go_y ? 2u : 4u);
    }
    return TGB_FAST_TREE;
}

/*
 * DDA phase: up to `steps` voxels of the leaf block. Returns DDA (budget used up), OCCLUDED (a solid voxel the ray certainly passes
 * through), TREE (left the block; the ray stands in the neighbouring cell), UNOCCLUDED (left the root).
 */
TGB_HD u32 tgb_fast_dda_phase(const tgb_gi_frame* f, tgb_fast_ray* r, u32 steps, u32* p_n_steps)
{
    const u32* p_block = f->p_voxels + (u64)r->data * TG_SVO_BLOCK_WORDS;
    const f32 rx = fabsf(r->inv.x), ry = fabsf(r->inv.y), rz = fabsf(r->inv.z);
    const i32 step_x = r->d.x > 0.0f ? 1 : -1, step_y = r->d.y > 0.0f ? 1 : -1, step_z = r->d.z > 0.0f ? 1 : -1;
    f32 tx = r->t_max.x, ty = r->t_max.y, tz = r->t_max.z, m_cur = r->m_cur;
    i32 x = (i32)(r->vox & 31u), y = (i32)((r->vox >> 5) & 31u), z = (i32)(r->vox >> 10);
    const f32 w = r->w_leaf, w2 = 2.0f * w;
    u32 uncertain = r->uncertain;
    u32 kind = TGB_FAST_DDA;
    u32 bits = TGB_LDG(&p_block[32 * z + y]);
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (u32 k = 0; k < steps; k++)
    {
        (*p_n_steps)++;
        const bool solid = ((bits >> x) & 1u) != 0;
        const f32 m_next = fminf(fminf(tx, ty), tz);
        const f32 gap = m_next - m_cur;
        if (gap < (solid ? w2 : w)) uncertain = 1;   /* a grazed solid voxel, or two crossings a displaced ray could swap */
        else if (solid)
        {
            /* certain unless it lies at the far plane (the shader skips the rest of the leaf there: left to the exact kernel) */
            if (r->t_cur + m_cur < TGB_FAST_FAR_FRACTION * f->far_plane) { kind = TGB_FAST_OCCLUDED; break; }
            uncertain = 1; kind = TGB_FAST_UNOCCLUDED; break;
        }
        const bool xy = tx < ty;
        const bool go_x = xy & (tx < tz), go_y = !xy & (ty < tz), go_z = !(go_x | go_y);
        tx = go_x ? tx + rx : tx;
        ty = go_y ? ty + ry : ty;
        tz = go_z ? tz + rz : tz;
        x += go_x ? step_x : 0;
        y += go_y ? step_y : 0;
        z += go_z ? step_z : 0;
        m_cur = m_next;
        if ((u32)(x | y | z) > 31u)
        {
            /* left the block through the plane of the axis that stepped: the neighbouring cell, entered at t_cur + m. A displaced ray
             * could cross the NEXT voxel plane before leaving and visit one more voxel of this block */
            if (fminf(fminf(tx, ty), tz) - m_cur < w) uncertain = 1;
            i32 cx = (i32)(r->cell & 31u), cy = (i32)((r->cell >> 5) & 31u), cz = (i32)(r->cell >> 10);
            cx += go_x ? step_x : 0; cy += go_y ? step_y : 0; cz += go_z ? step_z : 0;
            if ((u32)(cx | cy | cz) > 31u) { kind = TGB_FAST_UNOCCLUDED; break; }
            r->cell = (u32)cx | ((u32)cy << 5) | ((u32)cz << 10);
            r->entry_axis = go_x ? 1u : (go_y ? 2u : 4u);
            r->t_cur = r->t_cur + m_cur;
            r->p = tgb_v3(fmaf(r->t_cur, r->d.x, r->o.x), fmaf(r->t_cur, r->d.y, r->o.y), fmaf(r->t_cur, r->d.z, r->o.z));
            kind = TGB_FAST_TREE;
            break;
        }
        if (!go_x) bits = TGB_LDG(&p_block[32 * z + y]);
    }
    r->t_max = tgb_v3(tx, ty, tz);
    r->m_cur = m_cur;
    r->vox = ((u32)x & 31u) | (((u32)y & 31u) << 5) | (((u32)z & 31u) << 10);
    r->uncertain = uncertain;
    return kind;
}

#endif

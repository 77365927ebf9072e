/*
 * tgb_gi_fast.cuh -- the occluded / unoccluded decision of a secondary ray, CERTIFIED instead of transcribed.
 *
 *   SVO traversal      assets/shaders/raytracer/svo_functions.inc:1-329 (exact per-ray pieces: tgb_gi_walk.cuh)
 *
 * What the GI pass needs from tg_svo_traverse is one bit per ray: does the shader's traversal return a depth < 1
 * (occluded) or not. The exact kernels (k_gi_trace_pool, k_gi_trace_flat, k_gi_trace) obtain it by running the shader's own
 * arithmetic, operation for operation: an IEEE division per advanced cell, the shader's tie rules, its accumulated `position`,
 * and two different inner loops (tree cells, leaf voxels) that split every warp into two camps (12 - 17 of 32 lanes busy,
 * profiles/r02b_gi_pool_regions.txt).
 *
 * The shader's traversal is a geometric one: it visits the cells of the 1-bit SVO -- empty terminal boxes, voxels of the leaf
 * blocks -- that the ray o + t d passes through, in order, and stops at the first solid voxel (before the far plane). Its
 * floating-point path differs from the ideal line only by rounding: `position` is advanced by `position += (exit + eps) * d`
 * at most a few dozen times at coordinates below 512 (half an ulp(512) = 3e-5 per advance and component; along-ray errors do not
 * move the line; 2^-24 of the advanced distance per component for the product), and inside a leaf the DDA's t_max are sums of
 * a few dozen small floats. In other words, after n advances the shader decides like an exact traversal of a ray DISPLACED
 * sideways by less than n * 2.1e-5 units. So a cheap traversal of the ideal line reaches the same decision whenever no decision
 * it took depended on less than DELTA(n) = DELTA0 + n * DELTA1 units of displacement, DELTA0 = 2e-4 covering this walk's own
 * rounding (the origin is moved into the box's frame once: half an ulp(1024) = 6e-5; so is the position that selects the next
 * cell), DELTA1 = 3e-5 per box the ray has entered.
 *
 * The walk here has ONE kind of step for every cell, whatever its size (voxel 1, empty terminal box 32 .. 512): look the cell
 * around the ray's integer position up, take the times the ray crosses the cell's three near planes and three far planes
 * (one FMA each from the cell's corner), leave through the nearest far plane. All lanes of a warp run the same instructions
 * whether their rays cross empty space or walk a leaf. With W = DELTA(n) * sum 1 / |d_k| (the time by which a sideways displacement
 * of DELTA can shift any plane crossing, plus 2.5e-7 t for the rounding of the times themselves):
 *
 *   OCCLUDED   is certain when the cell is a solid voxel and the ray stays inside it for longer than 2 W: every ray displaced by
 *              less than DELTA passes through that voxel too, so the shader's traversal meets it -- or an earlier solid voxel;
 *              either way it returns occluded. (Unless the voxel lies at the far plane: svo_functions.inc:219-256 only ends on
 *              enter / far < 1; such rays are handed over.)
 *   UNOCCLUDED is certain when the ray left the root and no step on the way was UNCERTAIN. A displaced ray visits the same cells
 *              in the same order unless it passes an EDGE of a visited cell on the other side; a step is uncertain when the ray
 *              passes an edge of its cell within W in time: two near planes crossed within W of each other (entered next to an
 *              edge), two far planes within W (left next to an edge), or the far plane less than W after the near plane (cut a
 *              corner; for a solid voxel: 2 W, the grazed voxel then counts as uncertain and not as a hit). Where cells of
 *              different sizes meet, the finer cell's own check covers the lines that lie inside the coarser cell's face.
 *              A ray that starts inside a cell is checked against every plane of that cell on both sides.
 *
 * Every ray that ends unoccluded after an uncertain step, every ray with a direction component below TGB_FAST_SHALLOW (zero
 * included: W is then so large that nothing could be certified anyway) and every ray that runs into the step cap is handed to the exact
 * kernel (tgb_gi_pool.cu) through a list of queue slots. The decision taken here is therefore the shader's on every ray;
 * tests/test_gi_fast_cpu.py holds the host build of these functions against the exact state machine (itself held against the
 * oracle) on millions of rays, and tools/gi_fast_margin.py sweeps DELTA down to where the first disagreement appears (3e-6).
 * Host-compilable; the same IEEE operations on both sides (fmaf, 1 / x, floorf), so the host build predicts the device's
 * decisions AND its hand-overs.
 */
#ifndef TGB_GI_FAST_CUH
#define TGB_GI_FAST_CUH

#include "tgb_gi_walk.cuh"

#define TGB_FAST_DELTA        2.0e-4f  /* DELTA0: sideways displacement (world units) a decision must survive at the start of the ray */
#define TGB_FAST_DELTA_STEP   0.15f    /* DELTA1 / DELTA0: growth per box entered (3e-5 for DELTA0 = 2e-4) */
#define TGB_FAST_SHALLOW      1.0e-3f  /* direction components below this (zero included) go to the exact kernel */
#define TGB_FAST_MAX_STEPS    256u     /* cells a ray may enter here (median 5, 99.9 % below 220); the few that skim along a leaf layer for longer go to the exact kernel: */
#define TGB_FAST_MAX_STEPS_UNCERTAIN 64u /* one ray of 1,000 cells is a 0.2 ms dependent chain that the whole kernel waits for. A ray already uncertain can only still end occluded; it gets less */
#define TGB_FAST_FAR_FRACTION 0.99f    /* a solid voxel beyond this fraction of the far plane is not decided here */

/* kinds of the fast walk */
enum { TGB_FAST_IDLE = 0, TGB_FAST_WALK = 1, TGB_FAST_OCCLUDED = 2, TGB_FAST_UNOCCLUDED = 3, TGB_FAST_EXACT = 4 };

struct tgb_fast_ray
{
    v3  ob;                 /* origin relative to the box's min corner (in the shader's frame: origin - center - bmin) */
    v3  d, inv;             /* direction, 1 / d */
    v3  r;                  /* |inv| */
    v3  posf;               /* 1 where d_k > 0, else 0 */
    f32 w, w_step, w_t;     /* W = DELTA(n) * sum |inv_k| (time units), its growth per box and per unit of t */
    f32 t_cur;              /* ray parameter from which the ray is known to be in the current cell */
    i32 vx, vy, vz;         /* integer position in the box, 0 .. 1023 */
    u32 cell, entry;        /* the last 32^3 cell looked up and its table entry */
    u32 flags;              /* bit 0: an uncertain step happened; bit 1: first step of a ray that starts inside the box */
    u32 n_steps;
};

#define TGB_FAST_UNCERTAIN 1u
#define TGB_FAST_FIRST     2u

/* the fields that follow from d, inv and W(0) (a ready ray is stored without them, tgb_gi_fast.cu) */
TGB_HD void tgb_fast_derive(tgb_fast_ray* r, f32 w)
{
    r->r = tgb_v3(fabsf(r->inv.x), fabsf(r->inv.y), fabsf(r->inv.z));
    r->posf = tgb_v3(r->d.x > 0.0f ? 1.0f : 0.0f, r->d.y > 0.0f ? 1.0f : 0.0f, r->d.z > 0.0f ? 1.0f : 0.0f);
    r->w = w;
    r->w_step = TGB_FAST_DELTA_STEP * w;
    r->w_t = (2.5e-7f / TGB_FAST_DELTA) * w;
}

/*
 * A fresh ray from its queue record (origin, direction, `enter` of the slab test against the root). Returns WALK, or EXACT for the
 * rays the fast walk does not take (a direction component below TGB_FAST_SHALLOW). `delta` is TGB_FAST_DELTA in the product; the
 * margin sweep (tools/gi_fast_margin.py) lowers it until the first disagreement with the exact walk appears.
 */
TGB_HD u32 tgb_fast_start(const tgb_gi_frame* f, v3 origin, v3 dir, f32 root_enter, f32 delta, tgb_fast_ray* r)
{
    const f32 ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
    if (!(ax >= TGB_FAST_SHALLOW && ay >= TGB_FAST_SHALLOW && az >= TGB_FAST_SHALLOW)) return TGB_FAST_EXACT; /* NaN included */
    r->ob = tgb_sub(tgb_sub(origin, f->center), f->bmin);
    r->d = dir;
    r->inv = tgb_v3(TGB_RCP_RN(dir.x), TGB_RCP_RN(dir.y), TGB_RCP_RN(dir.z));
    tgb_fast_derive(r, delta * ((fabsf(r->inv.x) + fabsf(r->inv.y)) + fabsf(r->inv.z)));
    r->t_cur = root_enter > 0.0f ? root_enter : 0.0f;
    /* the voxel around the starting point; a ray that starts on (or outside) a root face is clamped into the outermost layer */
    const f32 px = fmaf(r->t_cur, dir.x, r->ob.x), py = fmaf(r->t_cur, dir.y, r->ob.y), pz = fmaf(r->t_cur, dir.z, r->ob.z);
    const f32 top = (f32)TG_SVO_SIDE_LENGTH - 1.0f;
    r->vx = (i32)tgb_clamp(floorf(px), 0.0f, top);
    r->vy = (i32)tgb_clamp(floorf(py), 0.0f, top);
    r->vz = (i32)tgb_clamp(floorf(pz), 0.0f, top);
    r->cell = 0xFFFFFFFFu;
    r->entry = 0;
    r->flags = root_enter > 0.0f ? 0u : TGB_FAST_FIRST;
    r->n_steps = 0;
    return TGB_FAST_WALK;
}

/*
 * Up to `steps` cells. Returns WALK (budget used up), OCCLUDED, UNOCCLUDED (left the root; certain only if no step was uncertain)
 * or EXACT (step cap).
 */
TGB_HD u32 tgb_fast_walk(const tgb_gi_frame* f, tgb_fast_ray* r, u32 steps, u32* p_n_cells, u32* p_n_voxels, u32 max_steps = TGB_FAST_MAX_STEPS, u32 max_steps_uncertain = TGB_FAST_MAX_STEPS_UNCERTAIN)
{
    i32 vx = r->vx, vy = r->vy, vz = r->vz;
    f32 t_cur = r->t_cur, w_n = r->w;
    u32 cell = r->cell, entry = r->entry, flags = r->flags;
    u32 kind = TGB_FAST_WALK;
    u32 k = 0;
    for (;;)
    {
        /* ---- the cell around (vx, vy, vz): a voxel of a leaf block, or the empty terminal box of the flattened tree ---- */
        const u32 c = (((u32)vz & 0x3E0u) << 5) | ((u32)vy & 0x3E0u) | ((u32)vx >> 5);
        const bool new_cell = c != cell;
        if (new_cell) { cell = c; entry = TGB_LDG(&f->p_grid[c]); }
        const bool leaf = (entry & TGB_TOP_HAS_DATA) != 0;
        u32 row = 0;
        if (leaf) row = TGB_LDG(&f->p_voxels[((entry & 0x3FFFFFu) << 10) | (((u32)vz & 31u) << 5) | ((u32)vy & 31u)]); /* < 2^22 leaves: the word index fits 32 bits */
        const bool solid = ((row >> ((u32)vx & 31u)) & 1u) != 0;
        if (p_n_cells) { if (leaf) (*p_n_voxels)++; else (*p_n_cells)++; }
        if (new_cell | !leaf) w_n += r->w_step; /* one more advance of the shader's `position` */
        const u32 size = leaf ? 1u : (512u >> ((entry >> TGB_TOP_LEVEL_SHIFT) & 7u));
        const u32 mask = 0u - size;
        const f32 size_f = (f32)size;
        /* crossing times of the far planes (difference form: exact when the ray is close to the plane) and of the near planes */
        const f32 fx = (fmaf(size_f, r->posf.x, (f32)((u32)vx & mask)) - r->ob.x) * r->inv.x;
        const f32 fy = (fmaf(size_f, r->posf.y, (f32)((u32)vy & mask)) - r->ob.y) * r->inv.y;
        const f32 fz = (fmaf(size_f, r->posf.z, (f32)((u32)vz & mask)) - r->ob.z) * r->inv.z;
        const f32 nx = fmaf(-size_f, r->r.x, fx), ny = fmaf(-size_f, r->r.y, fy), nz = fmaf(-size_f, r->r.z, fz);
        const f32 t_exit = fminf(fminf(fx, fy), fz);
        const f32 t_in = fmaxf(fmaxf(fmaxf(nx, ny), nz), t_cur);
        const f32 w = fmaf(t_cur, r->w_t, w_n);
        /* edges of the cell within W in time: near-near, far-far, near-far */
        const f32 t_lo = t_in - w, t_hi = t_exit + w;
        const bool bx = nx > t_lo, by = ny > t_lo, bz = nz > t_lo;
        const bool ex = fx < t_hi, ey = fy < t_hi, ez = fz < t_hi;
        bool uncertain = ((bx & by) | (bx & bz) | (by & bz)) | ((ex & ey) | (ex & ez) | (ey & ez));
        if (flags & TGB_FAST_FIRST) uncertain = uncertain | bx | by | bz; /* started inside the cell: any plane close behind */
        uncertain = uncertain | ((t_exit - t_in) < (solid ? w + w : w));
        if (solid && !uncertain)
        {
            if (t_in < TGB_FAST_FAR_FRACTION * f->far_plane) { kind = TGB_FAST_OCCLUDED; break; }
            flags |= TGB_FAST_UNCERTAIN; kind = TGB_FAST_UNOCCLUDED; break; /* at the far plane: the exact kernel decides */
        }
        flags = (flags & ~TGB_FAST_FIRST) | (uncertain ? TGB_FAST_UNCERTAIN : 0u);
        /* ---- leave through the nearest far plane(s): exact along the exit axis, from the position along the others ---- */
        /* time never runs backwards: where two planes are crossed within rounding of each other (an uncertain step anyway) the cell just
         * entered can claim to end before it began; evaluated at its own exit time the position would fall back behind the plane just
         * crossed and the walk would alternate between two cells (measured before this: 1 ray in 6,000 ran into the step cap, and those few
         * set the duration of the kernel). */
        const f32 t_next = fmaxf(t_exit, t_cur);
        const f32 px = fmaf(t_next, r->d.x, r->ob.x) + (fx == t_exit ? r->posf.x - 0.5f : 0.0f);
        const f32 py = fmaf(t_next, r->d.y, r->ob.y) + (fy == t_exit ? r->posf.y - 0.5f : 0.0f);
        const f32 pz = fmaf(t_next, r->d.z, r->ob.z) + (fz == t_exit ? r->posf.z - 0.5f : 0.0f);
        /* ... and no coordinate ever steps back against its direction of travel (a position within rounding of a plane just crossed can floor to the cell behind it) */
        const i32 qx = (i32)floorf(px), qy = (i32)floorf(py), qz = (i32)floorf(pz);
        vx = r->d.x > 0.0f ? (qx > vx ? qx : vx) : (qx < vx ? qx : vx);
        vy = r->d.y > 0.0f ? (qy > vy ? qy : vy) : (qy < vy ? qy : vy);
        vz = r->d.z > 0.0f ? (qz > vz ? qz : vz) : (qz < vz ? qz : vz);
        t_cur = t_next;
        if ((u32)(vx | vy | vz) >= (u32)TG_SVO_SIDE_LENGTH) { kind = TGB_FAST_UNOCCLUDED; break; } /* left the root */
        if (++k >= steps) break;
    }
    r->n_steps += k;
    if (kind == TGB_FAST_WALK && r->n_steps > ((flags & TGB_FAST_UNCERTAIN) ? max_steps_uncertain : max_steps)) kind = TGB_FAST_EXACT;
    r->vx = vx; r->vy = vy; r->vz = vz;
    r->t_cur = t_cur; r->w = w_n;
    r->cell = cell; r->entry = entry; r->flags = flags;
    return kind;
}

#endif

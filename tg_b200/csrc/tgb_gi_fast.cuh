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
#define TGB_FAST_SHALLOW_PER_AXIS 1.0e-6f /* the same for the walk with one margin per axis (tgb_fast_walk_tiled) */
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
/* `per_axis` (tgb_fast_walk_tiled): the ray carries DELTA itself, in world units, and the walk derives a time margin per axis from it;
 * directions down to TGB_FAST_SHALLOW_PER_AXIS are taken */
TGB_HD u32 tgb_fast_start(const tgb_gi_frame* f, v3 origin, v3 dir, f32 root_enter, f32 delta, tgb_fast_ray* r, bool per_axis = false)
{
    const f32 ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
    const f32 shallow = per_axis ? TGB_FAST_SHALLOW_PER_AXIS : TGB_FAST_SHALLOW;
    if (!(ax >= shallow && ay >= shallow && az >= shallow)) return TGB_FAST_EXACT; /* NaN included */
    r->ob = tgb_sub(tgb_sub(origin, f->center), f->bmin);
    r->d = dir;
    r->inv = tgb_v3(TGB_RCP_RN(dir.x), TGB_RCP_RN(dir.y), TGB_RCP_RN(dir.z));
    tgb_fast_derive(r, per_axis ? delta : delta * ((fabsf(r->inv.x) + fabsf(r->inv.y)) + fabsf(r->inv.z)));
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

/* ---- round 2, second half: the same walk over COARSER CELLS ------------------------------------------------------------------------
 *
 * The certificate above never uses that a cell is a node of the shader's octree: it needs the cells the ideal ray visits to be boxes
 * that tile the root without overlap, each either one voxel of a leaf block or free of solid voxels altogether. (Cells of different
 * sizes may meet anywhere: a line of the tiling that lies inside a coarser cell's face is an edge of the finer cell next to it, whose
 * own near-near / far-far / near-far check sees it.) Profiles of the walk over octree cells (profiles/r03e_*) and the host build on
 * the bench scene's rays say where its steps go: three rays of four leave the root, through ~9 empty terminal boxes and ~14 EMPTY
 * voxels of the leaf blocks around the objects each. So the empty space is re-tiled with larger boxes, once per SVO build:
 *
 *   top level   the 32^3 table cells that hold no leaf data are merged into boxes of cells: a cell's free run along x, runs of
 *               neighbouring rows along z with the same x extent, then along y with the same (x, z) extent (tgb_tile_*: three
 *               passes, every cell finds its box on its own, all cells of a box find the same one -- a tiling by construction);
 *   leaf level  a leaf block is 4^3 bricks of 8^3 voxels; bricks without a solid voxel are merged the same way inside their block;
 *   runs        inside a brick that holds solid voxels the free voxels are taken in runs along the ray's DOMINANT axis (a run = the free
 *               voxels of one voxel line between two solid ones, at most the brick's 8): one bit scan of the line's word, which the
 *               voxel test reads anyway -- from the block's rows when x dominates, from copies of the block with y or z as the bit index
 *               otherwise (tgb_fast_tiling::p_columns). Three tilings, then, and a ray stays with one.
 *
 * A step then enters one of: a voxel of a non-empty brick (side 1), a box of empty bricks (sides 8 .. 32), a box of empty table
 * cells (sides 32 .. 1024). What changes in the certificate is only the count of the shader's advances, which DELTA(n) grows with:
 * inside a box of table cells the shader advances once per terminal node it crosses, at most once per 32-unit plane the ray
 * crosses plus one (tgb_fast_advances), so W grows by that many steps when the box is entered.
 */
#define TGB_CELLS_LEAF       0x80000000u /* p_cells entry: leaf block with data, bits 0 .. 27 = data pointer; else the box of free cells around the cell, in cells:
                                            x0 | y0 << 5 | z0 << 10 | (sx - 1) << 15 | (sy - 1) << 20 | (sz - 1) << 25 */
#define TGB_BRICK_SOLID      0x80000000u /* p_bricks entry (64 per leaf block, brick = bz << 4 | by << 2 | bx): the brick holds a solid voxel; else the box of empty
                                            bricks around it in the same layout, in VOXELS of the block (corner a multiple of 8, sides 8 .. 32) */

/*
 * The tiling passes over an occupancy grid of side `side` (bit test `occ(x, y, z)`), one call per cell and pass. Runs are grown along
 * x first, then z, then y (the engine's up axis: the free space above a terrain becomes a few slabs):
 *   pass 1  -> x0 | x1 << 8                       the maximal free run along x through the cell
 *   pass 2  -> pass 1 | z0 << 16 | z1 << 24       the maximal run of z neighbours whose pass-1 value is the same
 *   pass 3  -> y0 | y1 << 8                       the maximal run of y neighbours whose pass-2 value is the same
 * Cells of one box compute identical values (a neighbour is joined only when its whole run equals this cell's), so the boxes tile.
 */
template <class OCC> TGB_HD u32 tgb_tile_pass1(OCC occ, u32 side, u32 x, u32 y, u32 z)
{
    u32 x0 = x, x1 = x;
    while (x0 > 0u && !occ(x0 - 1u, y, z)) x0--;
    while (x1 + 1u < side && !occ(x1 + 1u, y, z)) x1++;
    return x0 | (x1 << 8);
}
template <class OCC, class P1> TGB_HD u32 tgb_tile_pass2(OCC occ, P1 p1, u32 side, u32 x, u32 y, u32 z)
{
    const u32 mine = p1(x, y, z);
    u32 z0 = z, z1 = z;
    while (z0 > 0u && !occ(x, y, z0 - 1u) && p1(x, y, z0 - 1u) == mine) z0--;
    while (z1 + 1u < side && !occ(x, y, z1 + 1u) && p1(x, y, z1 + 1u) == mine) z1++;
    return mine | (z0 << 16) | (z1 << 24);
}
template <class OCC, class P2> TGB_HD u32 tgb_tile_pass3(OCC occ, P2 p2, u32 side, u32 x, u32 y, u32 z)
{
    const u32 mine = p2(x, y, z);
    u32 y0 = y, y1 = y;
    while (y0 > 0u && !occ(x, y0 - 1u, z) && p2(x, y0 - 1u, z) == mine) y0--;
    while (y1 + 1u < side && !occ(x, y1 + 1u, z) && p2(x, y1 + 1u, z) == mine) y1++;
    return y0 | (y1 << 8);
}
/* the entry of a free cell from its pass-2 and pass-3 values; BITS = 5 (table cells) or 2 (bricks) */
template <u32 BITS> TGB_HD u32 tgb_tile_entry(u32 p2, u32 p3)
{
    const u32 x0 = p2 & 0xFFu, x1 = (p2 >> 8) & 0xFFu, z0 = (p2 >> 16) & 0xFFu, z1 = p2 >> 24, y0 = p3 & 0xFFu, y1 = p3 >> 8;
    return x0 | (y0 << BITS) | (z0 << (2u * BITS)) | ((x1 - x0) << (3u * BITS)) | ((y1 - y0) << (4u * BITS)) | ((z1 - z0) << (5u * BITS));
}

/* the same for an empty brick (pass values in bricks) in voxels of its block */
TGB_HD u32 tgb_tile_brick_entry(u32 p2, u32 p3)
{
    const u32 x0 = p2 & 0xFFu, x1 = (p2 >> 8) & 0xFFu, z0 = (p2 >> 16) & 0xFFu, z1 = p2 >> 24, y0 = p3 & 0xFFu, y1 = p3 >> 8;
    return (x0 << 3) | (y0 << 8) | (z0 << 13) | ((((x1 - x0 + 1u) << 3) - 1u) << 15) | ((((y1 - y0 + 1u) << 3) - 1u) << 20) | ((((z1 - z0 + 1u) << 3) - 1u) << 25);
}

/* upper bound of the shader's advances while the ray crosses a box of table cells: one per terminal node, nodes are at least one cell
 * wide, so at most one per 32-unit plane crossed (per axis: distance / 32 + 1, and never more than the box has) plus one */
TGB_HD f32 tgb_fast_advances(f32 dt, f32 sum_abs_d, u32 planes_in_box)
{
    const f32 by_distance = floorf(dt * sum_abs_d * 0.03125f) + 3.0f;
    const f32 by_box = (f32)planes_in_box;
    return 1.0f + (by_distance < by_box ? by_distance : by_box);
}

struct tgb_fast_tiling
{
    const u32* p_cells;              /* [32^3] */
    const u32* p_bricks;             /* [n_leaves * 64] */
    const u32* p_columns;            /* [2][n_leaves * 1024] the blocks' voxels once more with y, then with z as the bit index (word 32 z + x / 32 y + x) */
    u64 columns_stride;              /* words between the two copies */
};

#ifdef __CUDA_ARCH__
#define tgb_clz(v) __clz((int)(v))
#define tgb_ffs(v) __ffs((int)(v))
#else
#define tgb_clz(v) __builtin_clz((unsigned)(v))
#define tgb_ffs(v) __builtin_ffs((int)(v))
#endif

/* the axis whose voxel lines the walk takes its runs from: 0 = x, 1 = y, 2 = z */
TGB_HD u32 tgb_fast_dominant_axis(v3 d)
{
    const f32 ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    return (ax >= ay && ax >= az) ? 0u : (ay >= az ? 1u : 2u);
}

/* is the voxel (x, y, z) free space -- outside the root, in a box of free cells, in an empty brick, or an empty voxel of a brick with solid ones? */
TGB_HD bool tgb_fast_voxel_free(const tgb_gi_frame* f, const tgb_fast_tiling* tl, i32 x, i32 y, i32 z)
{
    if ((u32)(x | y | z) >= (u32)TG_SVO_SIDE_LENGTH) return true;
    const u32 entry = TGB_LDG(&tl->p_cells[(((u32)z & 0x3E0u) << 5) | ((u32)y & 0x3E0u) | ((u32)x >> 5)]);
    if (!(entry & TGB_CELLS_LEAF)) return true;
    const u32 lp = entry & 0x0FFFFFFFu;
    const u32 be = TGB_LDG(&tl->p_bricks[(lp << 6) | (((u32)z & 24u) << 1) | (((u32)y & 24u) >> 1) | (((u32)x & 24u) >> 3)]);
    if (!(be & TGB_BRICK_SOLID)) return true;
    const u32 row = TGB_LDG(&f->p_voxels[(lp << 10) | (((u32)z & 31u) << 5) | ((u32)y & 31u)]);
    return ((row >> ((u32)x & 31u)) & 1u) == 0u;
}

/*
 * A step that passed an edge within the margins is not lost yet: what a displaced ray can do there is pass the edge on its other
 * side, through a cell the ideal ray does not visit. While the ideal ray is that close to the edge (an interval of at most 3 (w_x + w_y + w_z)
 * around the crossing at time t), every displaced ray stays inside the cube p(t) +- h, h_k = DELTA(n) + 3 (w_x + w_y + w_z) |d_k|; if every cell
 * that meets the cube is free space, nothing can be hit there whichever side is taken, and once past the edge the displaced ray is in
 * the cell the ideal ray is in. With h_k < 1/2 the cube meets at most the eight voxels around its corners.
 */
TGB_HD bool tgb_fast_cube_free(const tgb_gi_frame* f, const tgb_fast_tiling* tl, const tgb_fast_ray* r, f32 t, f32 m, f32 w_sum)
{
    const f32 hx = fmaf(3.0f * w_sum, fabsf(r->d.x), m), hy = fmaf(3.0f * w_sum, fabsf(r->d.y), m), hz = fmaf(3.0f * w_sum, fabsf(r->d.z), m);
    if (!(hx < 0.49f && hy < 0.49f && hz < 0.49f)) return false;
    const f32 px = fmaf(t, r->d.x, r->ob.x), py = fmaf(t, r->d.y, r->ob.y), pz = fmaf(t, r->d.z, r->ob.z);
    const i32 x0 = (i32)floorf(px - hx), x1 = (i32)floorf(px + hx), y0 = (i32)floorf(py - hy), y1 = (i32)floorf(py + hy), z0 = (i32)floorf(pz - hz), z1 = (i32)floorf(pz + hz);
    bool free_ = true;
    for (u32 i = 0; i < 8u; i++) free_ = free_ & tgb_fast_voxel_free(f, tl, (i & 1u) ? x1 : x0, (i & 2u) ? y1 : y0, (i & 4u) ? z1 : z0); /* no short cut: eight independent look-ups */
    return free_;
}

/*
 * tgb_fast_walk over the coarser tiling. Same contract; `r->entry` caches the p_cells entry of `r->cell`.
 *
 * ONE MARGIN PER AXIS. A sideways displacement of at most DELTA per coordinate moves the time at which the ray crosses a plane of axis k
 * by at most w_k = DELTA / |d_k| -- not by W = DELTA * sum 1 / |d_j|, which tgb_fast_walk charges to every plane alike (three to five times
 * w_k for a typical direction, and hopeless for a shallow one, which is why it does not take them). With n_k / f_k the crossing times
 * of the cell's near / far planes, a = the entry axis (n_a = max n_k), e = the exit axis (f_e = min f_k), a displaced ray crosses
 * plane k at n_k + s_k, |s_k| <= w_k, so it enters through the same face iff n_a - n_b > w_a + w_b for the other two axes b, leaves through
 * the same face iff f_b - f_e > w_b + w_e, and meets the cell at all iff f_e - n_a > w_a + w_e; a ray that starts inside its cell must have
 * every near plane b more than w_b behind it. The face it leaves through and the point where it does so (displaced by less than the
 * margins the next cell's own entry check demands) select the next cell, so a walk on which none of these inequalities failed visits
 * the cells every displaced ray visits. Each w_k also carries the rounding of the times themselves (2.5e-7 t, twice per comparison).
 * Here `r->w` is DELTA(n) itself, in world units (tgb_fast_start with per_axis).
 */
/* CUBE: with the cube check of steps that pass an edge (tgb_fast_cube_free). The bulk kernels run without it -- a rare, long, divergent piece of
 * code inside the hot loop cost them a fifth of their time (profiles/r04d_*) -- and hand such rays over; the second stage (k_gi_trace_list) walks
 * them again with it before it resorts to the shader's own arithmetic. */
template <bool CUBE>
TGB_HD u32 tgb_fast_walk_tiled(const tgb_gi_frame* f, const tgb_fast_tiling* tl, tgb_fast_ray* r, u32 steps, u32* p_n_cells, u32* p_n_voxels, u32 max_steps = TGB_FAST_MAX_STEPS, u32 max_steps_uncertain = TGB_FAST_MAX_STEPS_UNCERTAIN)
{
    i32 vx = r->vx, vy = r->vy, vz = r->vz;
    f32 t_cur = r->t_cur, w_n = r->w;
    u32 cell = r->cell, entry = r->entry, flags = r->flags;
    u32 kind = TGB_FAST_WALK;
    u32 k = 0;
    const f32 sum_abs_d = (fabsf(r->d.x) + fabsf(r->d.y)) + fabsf(r->d.z);
    const u32 axis = tgb_fast_dominant_axis(r->d);
    for (;;)
    {
        /* ---- the cell of the tiling around (vx, vy, vz): one decode for the three kinds. A box is (e, shift, base): corner and sides - 1 in
         * 5-bit fields of e, in units of 1 << shift, relative to base -- a box of free table cells (shift 5, base 0), a box of empty bricks or
         * a single voxel (shift 0, base = the block's corner). The brick entry and the voxel row are requested together. ---- */
        const u32 c = (((u32)vz & 0x3E0u) << 5) | ((u32)vy & 0x3E0u) | ((u32)vx >> 5);
        const bool new_cell = c != cell;
        if (new_cell) { cell = c; entry = TGB_LDG(&tl->p_cells[c]); }
        const bool leaf = (entry & TGB_CELLS_LEAF) != 0;
        u32 e = entry, shift = 5u, base_mask = 0u;
        bool solid = false;
        if (leaf)
        {
            const u32 lp = entry & 0x0FFFFFFFu;
            const u32 brick = (((u32)vz & 24u) << 1) | (((u32)vy & 24u) >> 1) | (((u32)vx & 24u) >> 3);
            const u32 be = TGB_LDG(&tl->p_bricks[(lp << 6) | brick]);
            /* the voxel line through (vx, vy, vz) along the dominant axis: word and bit index */
            const u32 lx = (u32)vx & 31u, ly = (u32)vy & 31u, lz = (u32)vz & 31u;
            const u32 line = axis == 0u ? ((lz << 5) | ly) : (axis == 1u ? ((lz << 5) | lx) : ((ly << 5) | lx));
            const u32 pos = axis == 0u ? lx : (axis == 1u ? ly : lz);
            const u32* p_lines = axis == 0u ? f->p_voxels : tl->p_columns + (axis == 1u ? 0u : tl->columns_stride);
            const u32 row = TGB_LDG(&p_lines[(lp << 10) | line]);
            const bool voxel = (be & TGB_BRICK_SOLID) != 0;
            solid = voxel & (((row >> pos) & 1u) != 0);
            /* the run of free voxels around pos inside the brick's 8 (a solid voxel is its own cell) */
            const u32 seg = (row >> (pos & 24u)) & 0xFFu, q = pos & 7u;
            const u32 below = seg & ((1u << q) - 1u), above = seg >> (q + 1u);
            const u32 lo = solid ? q : (below ? 32u - (u32)tgb_clz(below) : 0u);
            const u32 hi = solid ? q : (above ? q + (u32)tgb_ffs(above) - 1u : 7u);
            const u32 first = (pos & 24u) | lo, extra = hi - lo;
            const u32 run = axis == 0u ? (first | (ly << 5) | (lz << 10) | (extra << 15))
                          : (axis == 1u ? (lx | (first << 5) | (lz << 10) | (extra << 20)) : (lx | (ly << 5) | (first << 10) | (extra << 25)));
            e = voxel ? run : be;
            shift = 0u; base_mask = ~31u;
        }
        if (p_n_cells) { if (leaf) (*p_n_voxels)++; else (*p_n_cells)++; }
        const u32 mx = ((u32)vx & base_mask) + ((e & 31u) << shift), my = ((u32)vy & base_mask) + (((e >> 5) & 31u) << shift), mz = ((u32)vz & base_mask) + (((e >> 10) & 31u) << shift);
        const u32 ex = (e >> 15) & 31u, ey = (e >> 20) & 31u, ez = (e >> 25) & 31u;
        const u32 sx = (ex + 1u) << shift, sy = (ey + 1u) << shift, sz = (ez + 1u) << shift;
        const u32 planes = ex + ey + ez;   /* of a box of free cells: 32-unit planes inside it */
        const f32 sxf = (f32)sx, syf = (f32)sy, szf = (f32)sz;
        /* crossing times of the far planes (difference form: exact when the ray is close to the plane) and of the near planes */
        const f32 fx = (fmaf(sxf, r->posf.x, (f32)mx) - r->ob.x) * r->inv.x;
        const f32 fy = (fmaf(syf, r->posf.y, (f32)my) - r->ob.y) * r->inv.y;
        const f32 fz = (fmaf(szf, r->posf.z, (f32)mz) - r->ob.z) * r->inv.z;
        const f32 nx = fmaf(-sxf, r->r.x, fx), ny = fmaf(-syf, r->r.y, fy), nz = fmaf(-szf, r->r.z, fz);
        const f32 t_exit = fminf(fminf(fx, fy), fz);
        const f32 t_in = fmaxf(fmaxf(fmaxf(nx, ny), nz), t_cur);
        const f32 t_next = fmaxf(t_exit, t_cur);
        /* one more advance of the shader's `position` per leaf block entered, per terminal node crossed inside a box of free cells */
        if (leaf) { if (new_cell) w_n += r->w_step; }
        else w_n = fmaf(tgb_fast_advances(t_next - t_cur, sum_abs_d, planes), r->w_step, w_n);
        /* the margins: w_k = DELTA(n) / |d_k| + the rounding of a crossing time */
        const f32 e_t = t_next * r->w_t;
        const f32 wx = fmaf(w_n, r->r.x, e_t), wy = fmaf(w_n, r->r.y, e_t), wz = fmaf(w_n, r->r.z, e_t);
        const f32 n_max = fmaxf(fmaxf(nx, ny), nz);
        const f32 w_a = nx == n_max ? wx : (ny == n_max ? wy : wz), w_e = fx == t_exit ? wx : (fy == t_exit ? wy : wz);
        /* entered / left next to an edge: a second plane within the pair's margin of the entry / exit plane (the plane itself always counts) */
        const bool bx = n_max - nx < wx + w_a, by = n_max - ny < wy + w_a, bz = n_max - nz < wz + w_a;
        const bool cx = fx - t_exit < wx + w_e, cy = fy - t_exit < wy + w_e, cz = fz - t_exit < wz + w_e;
        bool near_edge = (bx & by) | (bx & bz) | (by & bz), far_edge = (cx & cy) | (cx & cz) | (cy & cz);
        /* started inside the cell: any near plane within its own margin behind the starting point */
        if (flags & TGB_FAST_FIRST) near_edge = near_edge | (t_cur - nx < wx) | (t_cur - ny < wy) | (t_cur - nz < wz);
        const f32 w = w_a + w_e;
        const bool brief = (t_exit - t_in) < (solid ? w + w : w);
        bool uncertain = near_edge | far_edge | brief;
        if (CUBE && uncertain && !solid)
        {
            /* the edges passed lie in free space on every side? (a cell met only briefly lies between its entry and its exit point) */
            const f32 w_sum = (wx + wy) + wz;
            uncertain = ((near_edge | brief) && !tgb_fast_cube_free(f, tl, r, t_in, w_n, w_sum)) || ((far_edge | brief) && !tgb_fast_cube_free(f, tl, r, t_next, w_n, w_sum));
        }
        if (solid && !uncertain)
        {
            if (t_in < TGB_FAST_FAR_FRACTION * f->far_plane) { kind = TGB_FAST_OCCLUDED; break; }
            flags |= TGB_FAST_UNCERTAIN; kind = TGB_FAST_UNOCCLUDED; break; /* at the far plane: the exact kernel decides */
        }
        flags = (flags & ~TGB_FAST_FIRST) | (uncertain ? TGB_FAST_UNCERTAIN : 0u);
        /* ---- leave through the nearest far plane(s); time and every coordinate monotone (see tgb_fast_walk) ---- */
        const f32 px = fmaf(t_next, r->d.x, r->ob.x) + (fx == t_exit ? r->posf.x - 0.5f : 0.0f);
        const f32 py = fmaf(t_next, r->d.y, r->ob.y) + (fy == t_exit ? r->posf.y - 0.5f : 0.0f);
        const f32 pz = fmaf(t_next, r->d.z, r->ob.z) + (fz == t_exit ? r->posf.z - 0.5f : 0.0f);
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

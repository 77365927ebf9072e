/*
 * TEST INFRASTRUCTURE. The per-ray traversal state machine of k_gi_trace_pool (tg_b200/csrc/tgb_gi_walk.cuh) compiled
 * by a plain C++ compiler and driven ray by ray on the host, so that its hit / miss decisions can be held against the
 * oracle's transcription of svo_functions.inc without a GPU (tests/test_gi_walk_cpu.py). The flattened tree is rebuilt
 * here with the rules of k_svo_flatten (tgb_svo.cu); phase budgets are parameters so that the test can vary them the
 * way the kernel's scheduler does (a ray may be suspended after any number of steps).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../../tg_b200/csrc/tgb_gi_walk.cuh"

extern "C" {

/* k_svo_flatten on the host: p_grid[32^3 + 1], last word non-zero = complete */
void tgbsim_flatten(const u32* p_nodes, const u32* p_leaf_data, u32 n_nodes, u32 n_leaves, u32* p_grid)
{
    p_grid[TGB_TOP_GRID_CELLS] = 1;
    for (u32 cell = 0; cell < TGB_TOP_GRID_CELLS; cell++)
    {
        const u32 cx = cell & 31u, cy = (cell >> 5) & 31u, cz = cell >> 10;
        u32 node = 0, entry = 0;
        bool ok = true, done = false;
        for (u32 level = 0; level < 5u && !done; level++)
        {
            const u32 shift = 4u - level;
            const u32 oct = ((cx >> shift) & 1u) | (((cy >> shift) & 1u) << 1) | (((cz >> shift) & 1u) << 2);
            const u32 node_data = p_nodes[node];
            const u32 child_pointer = node_data & 0xFFFFu, valid_mask = (node_data >> 16) & 0xFFu, leaf_mask = node_data >> 24;
            entry = level << TGB_TOP_LEVEL_SHIFT;
            if (((valid_mask >> oct) & 1u) == 0) { done = true; break; }
            const u32 child = node + child_pointer + (u32)__builtin_popcount(valid_mask & ((1u << oct) - 1u));
            if (child >= n_nodes) { ok = false; done = true; break; }
            if ((leaf_mask >> oct) & 1u)
            {
                if (level != 4u) ok = false;
                const u32 data_pointer = p_nodes[child];
                if (data_pointer >= n_leaves || data_pointer > TGB_TOP_POINTER_MASK) ok = false;
                else if (p_leaf_data[(uint64_t)data_pointer * 65u] != 0) entry |= TGB_TOP_HAS_DATA | data_pointer;
                done = true;
                break;
            }
            node = child;
        }
        if (!done) ok = false;
        p_grid[cell] = entry;
        if (!ok) p_grid[TGB_TOP_GRID_CELLS] = 0;
    }
}

/*
 * n rays (origin_ws, dir) through the flattened tree exactly as k_shade + k_gi_trace_pool handle them: root slab test
 * (svo_functions.inc:27-31; a miss is unoccluded), then service / tree / DDA phases until the ray is decided.
 * p_occluded[i] = 1 when the shader's traversal would return a depth < 1. Returns the number of rays that hit the
 * iteration cap (must be 0 on valid input).
 */
u32 tgbsim_gi_trace(const f32* p_bmin, const f32* p_bmax, f32 far_plane, const u32* p_grid, const u32* p_voxels, u32 n, const f32* p_origins, const f32* p_dirs,
                    u32 tree_reps, u32 dda_steps, u32 grid16, u8* p_occluded, u64* p_work /* [3]: look-ups, DDA steps, advances */)
{
    tgb_gi_frame fr;
    tgb_gi_frame_init(&fr, tgb_v3(p_bmin[0], p_bmin[1], p_bmin[2]), tgb_v3(p_bmax[0], p_bmax[1], p_bmax[2]), far_plane, p_grid, p_voxels);
    /* grid16: the kernel's 16-bit form of the table (tgb_top16_pack), read back through tgb_top16_unpack */
    const bool uniform_dda = (grid16 & 2u) != 0; /* the list kernel's branch form of the leaf DDA (tgb_gi_dda_phase_uniform) */
    grid16 &= 1u;
    unsigned short* p_grid16 = NULL;
    if (grid16)
    {
        p_grid16 = (unsigned short*)malloc(TGB_TOP_GRID_CELLS * sizeof(unsigned short));
        for (u32 c = 0; c < TGB_TOP_GRID_CELLS; c++) p_grid16[c] = (unsigned short)tgb_top16_pack(p_grid[c]);
        fr.p_grid16 = p_grid16;
    }
    u32 n_capped = 0;
    for (u32 i = 0; i < n; i++)
    {
        const v3 origin = tgb_v3(p_origins[3 * i], p_origins[3 * i + 1], p_origins[3 * i + 2]);
        const v3 d = tgb_v3(p_dirs[3 * i], p_dirs[3 * i + 1], p_dirs[3 * i + 2]);
        f32 e0, e1;
        if (!tgb_ray_aabb(tgb_sub(origin, fr.center), d, fr.bmin, fr.bmax, &e0, &e1)) { p_occluded[i] = 0; continue; } /* k_shade: not queued */
        v3 position, t_delta, t_max = tgb_v3(0, 0, 0);
        u32 flags, cell = 0, data = 0, kind = TGB_RAY_TREE;
        i32 x = 0, y = 0, z = 0;
        u32 n_visits = 0, n_steps = 0, n_advances = 0;
        tgb_gi_ray_start(&fr, origin, d, e0, &position, &t_delta, &flags);
        bool occluded = false;
        for (;;)
        {
            if (kind == TGB_RAY_TREE) kind = grid16 ? tgb_gi_tree_phase_t<2>(&fr, d, t_delta, &position, &cell, &flags, &data, tree_reps, &n_visits, &n_advances)
                                                    : tgb_gi_tree_phase(&fr, d, t_delta, &position, &cell, &flags, &data, tree_reps, &n_visits, &n_advances);
            else if (kind == TGB_RAY_DDA)
            {
                if (flags & TGB_RF_SETUP)
                {
                    flags &= ~TGB_RF_SETUP;
                    v3 child_min; f32 child_size;
                    tgb_cell_box(&fr, cell, &child_min, &child_size);
                    tgb_gi_dda_setup(d, position, child_min, child_size, &x, &y, &z, &t_max);
                }
                kind = uniform_dda ? tgb_gi_dda_phase_uniform(p_voxels + (uint64_t)data * TG_SVO_BLOCK_WORDS, t_delta, flags >> TGB_RF_STEP_SHIFT, &t_max, &x, &y, &z, dda_steps, &n_steps)
                                   : tgb_gi_dda_phase(p_voxels + (uint64_t)data * TG_SVO_BLOCK_WORDS, t_delta, flags >> TGB_RF_STEP_SHIFT, &t_max, &x, &y, &z, dda_steps, &n_steps);
                /* what the kernel stores between phases: 5 bits per coordinate */
                if (kind == TGB_RAY_DDA || kind == TGB_RAY_HIT) { x &= 31; y &= 31; z &= 31; }
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
                if (kind == TGB_RAY_MISS) { if ((cell >> 18) > TGB_TRAVERSE_MAX_ITERS) n_capped++; break; }
            }
        }
        p_occluded[i] = occluded ? 1 : 0;
        if (p_work) { p_work[0] += n_visits; p_work[1] += n_steps; p_work[2] += n_advances; }
    }
    free(p_grid16);
    return n_capped;
}

}

#include "../../tg_b200/csrc/tgb_svo_traverse.cuh"

extern "C" {

/* tgb_svo_traverse_stack (the BLOCKS view's traversal) for n rays: result bits, node index, voxel index, packed word */
void tgbsim_svo_traverse(const u32* p_nodes, const u32* p_leaf_data, const u32* p_voxels, const f32* p_bmin, const f32* p_bmax, f32 far_plane, u32 n,
                         const f32* p_origins, const f32* p_dirs, f32* p_result, u32* p_node_idx, u32* p_voxel_idx, u64* p_word)
{
    for (u32 i = 0; i < n; i++)
    {
        p_result[i] = tgb_svo_traverse_stack(p_nodes, p_leaf_data, p_voxels, tgb_v3(p_bmin[0], p_bmin[1], p_bmin[2]), tgb_v3(p_bmax[0], p_bmax[1], p_bmax[2]), far_plane,
                                             tgb_v3(p_origins[3 * i], p_origins[3 * i + 1], p_origins[3 * i + 2]), tgb_v3(p_dirs[3 * i], p_dirs[3 * i + 1], p_dirs[3 * i + 2]),
                                             &p_node_idx[i], &p_voxel_idx[i]);
        p_word[i] = tgb_svo_visibility_word(p_result[i], p_node_idx[i], p_voxel_idx[i]);
    }
}

}

#include "../../tg_b200/csrc/tgb_gi_fast.cuh"

extern "C" {

/*
 * The certified fast walk (tgb_gi_fast.cuh: what k_gi_trace_fast runs per ray) for n rays. p_result[i]: 0 = unoccluded, 1 = occluded,
 * 2 = handed to the exact kernel (uncertain and unoccluded, shallow direction, cap). Rays that miss the root are unoccluded
 * (k_shade does not queue them). p_work[3]: empty boxes entered, voxels entered, rays handed over.
 */
void tgbsim_gi_fast(const f32* p_bmin, const f32* p_bmax, f32 far_plane, const u32* p_grid, const u32* p_voxels, u32 n, const f32* p_origins, const f32* p_dirs,
                    u32 steps, f32 delta, u8* p_result, u64* p_work)
{
    tgb_gi_frame fr;
    tgb_gi_frame_init(&fr, tgb_v3(p_bmin[0], p_bmin[1], p_bmin[2]), tgb_v3(p_bmax[0], p_bmax[1], p_bmax[2]), far_plane, p_grid, p_voxels);
    for (u32 i = 0; i < n; i++)
    {
        const v3 origin = tgb_v3(p_origins[3 * i], p_origins[3 * i + 1], p_origins[3 * i + 2]);
        const v3 d = tgb_v3(p_dirs[3 * i], p_dirs[3 * i + 1], p_dirs[3 * i + 2]);
        f32 e0, e1;
        if (!tgb_ray_aabb(tgb_sub(origin, fr.center), d, fr.bmin, fr.bmax, &e0, &e1)) { p_result[i] = 0; continue; }
        tgb_fast_ray r;
        memset(&r, 0, sizeof r);
        u32 n_cells = 0, n_voxels = 0;
        u32 kind = tgb_fast_start(&fr, origin, d, e0, delta, &r);
        while (kind == TGB_FAST_WALK) kind = tgb_fast_walk(&fr, &r, steps, &n_cells, &n_voxels);
        if (kind == TGB_FAST_UNOCCLUDED && (r.flags & TGB_FAST_UNCERTAIN)) kind = TGB_FAST_EXACT;
        p_result[i] = kind == TGB_FAST_OCCLUDED ? 1 : (kind == TGB_FAST_UNOCCLUDED ? 0 : 2);
        if (p_work) { p_work[0] += n_cells; p_work[1] += n_voxels; p_work[2] += kind == TGB_FAST_EXACT ? 1 : 0; }
    }
}

}

extern "C" {
/* steps the fast walk takes per ray (diagnostics for tools/gi_fast_margin.py) */
void tgbsim_gi_fast_steps(const f32* p_bmin, const f32* p_bmax, f32 far_plane, const u32* p_grid, const u32* p_voxels, u32 n, const f32* p_origins, const f32* p_dirs, u32* p_steps)
{
    tgb_gi_frame fr;
    tgb_gi_frame_init(&fr, tgb_v3(p_bmin[0], p_bmin[1], p_bmin[2]), tgb_v3(p_bmax[0], p_bmax[1], p_bmax[2]), far_plane, p_grid, p_voxels);
    for (u32 i = 0; i < n; i++)
    {
        const v3 origin = tgb_v3(p_origins[3 * i], p_origins[3 * i + 1], p_origins[3 * i + 2]);
        const v3 d = tgb_v3(p_dirs[3 * i], p_dirs[3 * i + 1], p_dirs[3 * i + 2]);
        f32 e0, e1;
        p_steps[i] = 0;
        if (!tgb_ray_aabb(tgb_sub(origin, fr.center), d, fr.bmin, fr.bmax, &e0, &e1)) continue;
        tgb_fast_ray r;
        memset(&r, 0, sizeof r);
        u32 kind = tgb_fast_start(&fr, origin, d, e0, TGB_FAST_DELTA, &r);
        while (kind == TGB_FAST_WALK) kind = tgb_fast_walk(&fr, &r, 8, (u32*)0, (u32*)0);
        p_steps[i] = r.n_steps;
    }
}
}

extern "C" {

/*
 * The coarser tiling of the fast walk (tgb_gi_fast.cuh, second half) built on the host with the per-cell passes the kernels
 * k_fast_tile_* run (tgb_gi_fast.cu): p_cells[32^3] from the flattened tree, p_bricks[64 * n_leaves] from the leaf blocks' voxels.
 */
void tgbsim_fast_tiling(const u32* p_grid, const u32* p_voxels, u32 n_leaves, u32* p_cells, u32* p_bricks, u32* p_columns /* [2][n_leaves * 1024] */)
{
    /* the blocks once more with y / z as the bit index (k_fast_tile_columns) */
    for (u32 leaf = 0; leaf < n_leaves; leaf++)
    {
        const u32* block = p_voxels + (uint64_t)leaf * 1024u;
        u32* cy = p_columns + (uint64_t)leaf * 1024u, *cz = p_columns + (uint64_t)n_leaves * 1024u + (uint64_t)leaf * 1024u;
        for (u32 i = 0; i < 1024u; i++) { cy[i] = 0; cz[i] = 0; }
        for (u32 z = 0; z < 32u; z++) for (u32 y = 0; y < 32u; y++) for (u32 x = 0; x < 32u; x++)
            if ((block[32u * z + y] >> x) & 1u) { cy[32u * z + x] |= 1u << y; cz[32u * y + x] |= 1u << z; }
    }
    const u32 N = TGB_TOP_GRID_DIM;
    u32* p1 = (u32*)malloc(TGB_TOP_GRID_CELLS * sizeof(u32));
    u32* p2 = (u32*)malloc(TGB_TOP_GRID_CELLS * sizeof(u32));
    auto occ = [&](u32 x, u32 y, u32 z) { return (p_grid[(z << 10) | (y << 5) | x] & TGB_TOP_HAS_DATA) != 0; };
    auto g1 = [&](u32 x, u32 y, u32 z) { return p1[(z << 10) | (y << 5) | x]; };
    auto g2 = [&](u32 x, u32 y, u32 z) { return p2[(z << 10) | (y << 5) | x]; };
    for (u32 c = 0; c < TGB_TOP_GRID_CELLS; c++) p1[c] = tgb_tile_pass1(occ, N, c & 31u, (c >> 5) & 31u, c >> 10);
    for (u32 c = 0; c < TGB_TOP_GRID_CELLS; c++) p2[c] = tgb_tile_pass2(occ, g1, N, c & 31u, (c >> 5) & 31u, c >> 10);
    for (u32 c = 0; c < TGB_TOP_GRID_CELLS; c++)
    {
        const u32 x = c & 31u, y = (c >> 5) & 31u, z = c >> 10;
        if (occ(x, y, z)) p_cells[c] = TGB_CELLS_LEAF | (p_grid[c] & TGB_TOP_POINTER_MASK);
        else p_cells[c] = tgb_tile_entry<5>(p2[c], tgb_tile_pass3(occ, g2, N, x, y, z));
    }
    free(p1); free(p2);
    for (u32 leaf = 0; leaf < n_leaves; leaf++)
    {
        const u32* block = p_voxels + (uint64_t)leaf * TG_SVO_BLOCK_WORDS;
        u32 solid[64], b1[64], b2[64];
        for (u32 b = 0; b < 64u; b++)
        {
            const u32 bx = b & 3u, by = (b >> 2) & 3u, bz = b >> 4;
            u32 any = 0;
            for (u32 z = 0; z < 8u; z++) for (u32 y = 0; y < 8u; y++) any |= (block[32u * (8u * bz + z) + 8u * by + y] >> (8u * bx)) & 0xFFu;
            solid[b] = any != 0;
        }
        auto bocc = [&](u32 x, u32 y, u32 z) { return solid[(z << 4) | (y << 2) | x] != 0; };
        auto bg1 = [&](u32 x, u32 y, u32 z) { return b1[(z << 4) | (y << 2) | x]; };
        auto bg2 = [&](u32 x, u32 y, u32 z) { return b2[(z << 4) | (y << 2) | x]; };
        for (u32 b = 0; b < 64u; b++) b1[b] = tgb_tile_pass1(bocc, 4u, b & 3u, (b >> 2) & 3u, b >> 4);
        for (u32 b = 0; b < 64u; b++) b2[b] = tgb_tile_pass2(bocc, bg1, 4u, b & 3u, (b >> 2) & 3u, b >> 4);
        for (u32 b = 0; b < 64u; b++)
            p_bricks[leaf * 64u + b] = solid[b] ? TGB_BRICK_SOLID : tgb_tile_brick_entry(b2[b], tgb_tile_pass3(bocc, bg2, 4u, b & 3u, (b >> 2) & 3u, b >> 4));
    }
}

/* tgbsim_gi_fast over the coarser tiling; cube = 0: as the bulk kernels walk (k_shade, k_gi_trace_fast), 1: as the second stage walks the rays they hand
 * over (cube check of near-edge steps, no step caps to speak of: k_gi_trace_list); p_steps (optional): cells entered per ray */
void tgbsim_gi_fast_tiled(const f32* p_bmin, const f32* p_bmax, f32 far_plane, const u32* p_grid, const u32* p_voxels, const u32* p_cells, const u32* p_bricks, const u32* p_columns, u32 n_leaves,
                          u32 n, const f32* p_origins, const f32* p_dirs, u32 steps, f32 delta, u32 cube, u8* p_result, u64* p_work, u32* p_steps)
{
    tgb_gi_frame fr;
    tgb_gi_frame_init(&fr, tgb_v3(p_bmin[0], p_bmin[1], p_bmin[2]), tgb_v3(p_bmax[0], p_bmax[1], p_bmax[2]), far_plane, p_grid, p_voxels);
    tgb_fast_tiling tl; tl.p_cells = p_cells; tl.p_bricks = p_bricks; tl.p_columns = p_columns; tl.columns_stride = (u64)n_leaves * 1024u;
    for (u32 i = 0; i < n; i++)
    {
        const v3 origin = tgb_v3(p_origins[3 * i], p_origins[3 * i + 1], p_origins[3 * i + 2]);
        const v3 d = tgb_v3(p_dirs[3 * i], p_dirs[3 * i + 1], p_dirs[3 * i + 2]);
        f32 e0, e1;
        if (p_steps) p_steps[i] = 0;
        if (!tgb_ray_aabb(tgb_sub(origin, fr.center), d, fr.bmin, fr.bmax, &e0, &e1)) { p_result[i] = 0; continue; }
        tgb_fast_ray r;
        memset(&r, 0, sizeof r);
        u32 n_cells = 0, n_voxels = 0;
        u32 kind = tgb_fast_start(&fr, origin, d, e0, delta, &r, true);
        while (kind == TGB_FAST_WALK) kind = cube ? tgb_fast_walk_tiled<true>(&fr, &tl, &r, steps, &n_cells, &n_voxels, 4096u, 4096u) : tgb_fast_walk_tiled<false>(&fr, &tl, &r, steps, &n_cells, &n_voxels);
        if (kind == TGB_FAST_UNOCCLUDED && (r.flags & TGB_FAST_UNCERTAIN)) kind = TGB_FAST_EXACT;
        p_result[i] = kind == TGB_FAST_OCCLUDED ? 1 : (kind == TGB_FAST_UNOCCLUDED ? 0 : 2);
        if (p_work) { p_work[0] += n_cells; p_work[1] += n_voxels; p_work[2] += kind == TGB_FAST_EXACT ? 1 : 0; }
        if (p_steps) p_steps[i] = n_cells + n_voxels;
    }
}

}

/*
 * tgb_svo_traverse.cuh -- tg_svo_traverse of assets/shaders/raytracer/svo_functions.inc:1-329 as ONE plain loop per ray: the
 * shader's stack machine, statement for statement, returning what the shader returns (result, node index, voxel index).
 * Used by the BLOCKS debug view's primary-ray pass (tgb_debug_svo.cu: debug_visibility_svo.frag:27-71), where the node index
 * of the leaf is part of the written word -- the flattened tree of the GI kernels does not keep it. Host-compilable
 * (tests/cpu_sim) so that it can be held against the oracle's transcription without a GPU.
 */
#ifndef TGB_SVO_TRAVERSE_CUH
#define TGB_SVO_TRAVERSE_CUH

#include "tgb_gi_walk.cuh"

TGB_HD f32 tgb_svo_traverse_stack(const u32* p_nodes, const u32* p_leaf_data, const u32* p_voxels, v3 bmin, v3 bmax, f32 far_plane,
                                  v3 ray_origin_ws, v3 d, u32* p_node_idx, u32* p_voxel_idx)
{
    const v3 extent = tgb_sub(bmax, bmin);
    const v3 center = tgb_add(tgb_scale(extent, 0.5f), bmin); /* :3-8 */
    const v3 o = tgb_sub(ray_origin_ws, center);

    u32 idx_stack[TG_SVO_TRAVERSE_STACK_CAPACITY];
    v3  min_stack[TG_SVO_TRAVERSE_STACK_CAPACITY], max_stack[TG_SVO_TRAVERSE_STACK_CAPACITY];
    f32 result = 1.0f;
    *p_node_idx = TG_U32_MAX;
    *p_voxel_idx = TG_U32_MAX;

    f32 enter, exit;
    if (!tgb_ray_aabb(o, d, bmin, bmax, &enter, &exit)) return result; /* :21-22 */
    v3 position = o;
    if (enter > 0.0f) position = tgb_add(position, tgb_scale(d, enter));

    u32 stack_size = 1;
    idx_stack[0] = 0; min_stack[0] = bmin; max_stack[0] = bmax;
    u32 iterations = 0;
    while (stack_size > 0)
    {
        if (++iterations > TGB_TRAVERSE_MAX_ITERS) return 1.0f; /* Q9 */
        const u32 parent_idx = idx_stack[stack_size - 1];
        const v3 parent_min = min_stack[stack_size - 1], parent_max = max_stack[stack_size - 1];
        const u32 node_data = TGB_LDG(&p_nodes[parent_idx]);
        const u32 child_pointer = node_data & 0xFFFFu, valid_mask = (node_data >> 16) & 0xFFu, leaf_mask = (node_data >> 24) & 0xFFu;

        /* :57-80 */
        const v3 child_extent = tgb_scale(tgb_sub(parent_max, parent_min), 0.5f);
        u32 relative_child_idx = 0;
        v3 child_min = parent_min;
        v3 child_max = tgb_add(child_min, child_extent);
        if (child_max.x < position.x || (position.x == child_max.x && d.x > 0.0f)) { relative_child_idx += 1; child_min.x += child_extent.x; child_max.x += child_extent.x; }
        if (child_max.y < position.y || (position.y == child_max.y && d.y > 0.0f)) { relative_child_idx += 2; child_min.y += child_extent.y; child_max.y += child_extent.y; }
        if (child_max.z < position.z || (position.z == child_max.z && d.z > 0.0f)) { relative_child_idx += 4; child_min.z += child_extent.z; child_max.z += child_extent.z; }

        bool advance_to_border = true;
        if ((valid_mask & (1u << relative_child_idx)) != 0)
        {
            /* :86-91 */
            u32 relative_child_offset = 0;
            for (u32 i = 0; i < relative_child_idx; i++) relative_child_offset += (valid_mask >> i) & 1u;
            const u32 child_idx = parent_idx + child_pointer + relative_child_offset;
            if ((leaf_mask & (1u << relative_child_idx)) != 0)
            {
                const u32 data_pointer = TGB_LDG(&p_nodes[child_idx]);
                if (TGB_LDG(&p_leaf_data[(u64)data_pointer * 65u]) != 0)
                {
                    /* :111-176 */
                    const u32 first_voxel_idx = data_pointer * TG_SVO_BLOCK_VOXEL_COUNT;
                    v3 hit = position;
                    v3 xyz = tgb_v3(tgb_clamp(floorf(hit.x), child_min.x, child_max.x - 1.0f),
                                    tgb_clamp(floorf(hit.y), child_min.y, child_max.y - 1.0f),
                                    tgb_clamp(floorf(hit.z), child_min.z, child_max.z - 1.0f));
                    hit = tgb_sub(hit, child_min);
                    xyz = tgb_sub(xyz, child_min);
                    i32 x = (i32)xyz.x, y = (i32)xyz.y, z = (i32)xyz.z;
                    i32 step_x = 0, step_y = 0, step_z = 0;
                    f32 t_max_x = TG_F32_MAX, t_max_y = TG_F32_MAX, t_max_z = TG_F32_MAX;
                    f32 t_delta_x = TG_F32_MAX, t_delta_y = TG_F32_MAX, t_delta_z = TG_F32_MAX;
                    if (d.x > 0.0f)      { step_x = 1;  t_max_x = ((f32)(x + 1) - hit.x) / d.x; t_delta_x = 1.0f / d.x; }
                    else if (d.x < 0.0f) { step_x = -1; t_max_x = (hit.x - (f32)x) / -d.x;      t_delta_x = 1.0f / -d.x; }
                    if (d.y > 0.0f)      { step_y = 1;  t_max_y = ((f32)(y + 1) - hit.y) / d.y; t_delta_y = 1.0f / d.y; }
                    else if (d.y < 0.0f) { step_y = -1; t_max_y = (hit.y - (f32)y) / -d.y;      t_delta_y = 1.0f / -d.y; }
                    if (d.z > 0.0f)      { step_z = 1;  t_max_z = ((f32)(z + 1) - hit.z) / d.z; t_delta_z = 1.0f / d.z; }
                    else if (d.z < 0.0f) { step_z = -1; t_max_z = (hit.z - (f32)z) / -d.z;      t_delta_z = 1.0f / -d.z; }

                    /* :178-257 */
                    const u32 ex = (u32)child_extent.x, ey = (u32)child_extent.y;
                    for (;;)
                    {
                        const u32 relative_voxel_idx = ex * ey * (u32)z + ex * (u32)y + (u32)x;
                        const u32 voxel_idx = first_voxel_idx + relative_voxel_idx;
                        const u32 bits = TGB_LDG(&p_voxels[voxel_idx / 32u]);
                        if ((bits & (1u << (voxel_idx % 32u))) != 0)
                        {
                            const v3 voxel_min = tgb_add(child_min, tgb_v3((f32)x, (f32)y, (f32)z));
                            const v3 voxel_max = tgb_add(child_min, tgb_v3((f32)(x + 1), (f32)(y + 1), (f32)(z + 1)));
                            tgb_ray_aabb(o, d, voxel_min, voxel_max, &enter, &exit);
                            result = enter / far_plane; /* :219-256; hit position / normal are not used by any caller here */
                            *p_node_idx = child_idx;
                            *p_voxel_idx = relative_voxel_idx;
                            break;
                        }
                        if (t_max_x < t_max_y)
                        {
                            if (t_max_x < t_max_z) { t_max_x += t_delta_x; x += step_x; if (x < 0 || (f32)x >= child_extent.x) break; }
                            else                   { t_max_z += t_delta_z; z += step_z; if (z < 0 || (f32)z >= child_extent.z) break; }
                        }
                        else
                        {
                            if (t_max_y < t_max_z) { t_max_y += t_delta_y; y += step_y; if (y < 0 || (f32)y >= child_extent.y) break; }
                            else                   { t_max_z += t_delta_z; z += step_z; if (z < 0 || (f32)z >= child_extent.z) break; }
                        }
                    }
                    if (result < 1.0f) break;
                }
            }
            else
            {
                /* :262-270 */
                advance_to_border = false;
                if (stack_size >= TG_SVO_TRAVERSE_STACK_CAPACITY) return result; /* malformed tree: deeper than the shader's stack */
                idx_stack[stack_size] = child_idx;
                min_stack[stack_size] = child_min;
                max_stack[stack_size] = child_max;
                stack_size++;
            }
        }

        if (advance_to_border)
        {
            /* :279-324 */
            exit = tgb_exit_distance(child_min, child_max, position, d);
            position = tgb_add(position, tgb_scale(d, exit + TG_F32_EPSILON));
            while (stack_size > 0)
            {
                if (tgb_exit_distance(min_stack[stack_size - 1], max_stack[stack_size - 1], position, d) > TG_F32_EPSILON) break;
                stack_size--;
            }
        }
    }
    return result;
}

/* debug_visibility_svo.frag:52-71: the word of one pixel, or TG_VIS_CLEAR when nothing is written (d > 1) */
TGB_HD u64 tgb_svo_visibility_word(f32 d, u32 node_idx, u32 voxel_idx)
{
    if (!(d <= 1.0f)) return TG_VIS_CLEAR;
    const f32 dq = d * TG_VIS_DEPTH_SCALE;
    const u64 depth_24b = dq > 0.0f ? (u64)dq : 0ull; /* negative / NaN -> 0 (cvt.rzi.u64.f32); GLSL leaves it undefined */
    return (depth_24b << TG_VIS_DEPTH_SHIFT) | ((u64)(node_idx & 2147483647u) << TG_VIS_POINTER_SHIFT) | (u64)(voxel_idx % 512u);
}

#endif

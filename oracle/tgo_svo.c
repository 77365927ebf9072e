/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY. Nothing under tg_b200/ may include, link or call this.
 *
 * tgo_svo.c: the reference's 1-bit sparse voxel octree, scalar C.
 *   build            graphics/tg_sparse_voxel_octree.c:23-542  (CPU-only in the reference)
 *   AABB/OBB SAT     physics/tg_physics.c:226-392
 *   traversal (GLSL) assets/shaders/raytracer/svo_functions.inc:1-329   <- what the GI kernel follows
 *   traversal (C)    graphics/tg_sparse_voxel_octree.c:558-740 + util/tg_amanatides_woo.c:3-114
 *
 * Reproduced quirks: Q4 leaf AABB uses corners [0,1,2,2,4,6,5,7]; Q5 SAT
 * over face axes only with normalise-then-scale plane distances; trilinear voxel-centre mapping.
 * ONE defined deviation: the reference appends to p_cluster_idcs[64] without a bound
 * (tg_sparse_voxel_octree.c:307-311: TG_ASSERT only, release builds write past the array); here
 * a leaf keeps its first 64 cluster indices and drops the rest. Bits and `n != 0` are unaffected.
 * Also Q9: the traversals carry an iteration cap (4096) that valid input never reaches.
 */
#include <stdlib.h>
#include <string.h>

#include "tgo.h"
#include "tgo_math.h"

#define TGO_TRAVERSE_MAX_ITERATIONS 4096

/* tg_sparse_voxel_octree.c:23-60 */
static m4 tgo__ws2ms_c(const tg_voxel_object* p_object, u32 relative_cluster_pointer)
{
    const m4 rotation = tgo_m4_angle_axis(p_object->angle_in_radians, p_object->axis);
    const u32 rx = relative_cluster_pointer % p_object->n_cluster_pointers_per_dim.x;
    const u32 ry = (relative_cluster_pointer / p_object->n_cluster_pointers_per_dim.x) % p_object->n_cluster_pointers_per_dim.y;
    const u32 rz = relative_cluster_pointer / (p_object->n_cluster_pointers_per_dim.x * p_object->n_cluster_pointers_per_dim.y);
    const v3 relative_cluster_offset = tgo_v3((f32)(rx * 8u), (f32)(ry * 8u), (f32)(rz * 8u));
    const f32 h = (f32)(8 / 2);
    const v3 cluster_half_extent = tgo_v3(h, h, h);
    const v3 dims = tgo_v3((f32)p_object->n_cluster_pointers_per_dim.x, (f32)p_object->n_cluster_pointers_per_dim.y, (f32)p_object->n_cluster_pointers_per_dim.z);
    const v3 object_half_extent = tgo_v3_mul(dims, cluster_half_extent);

    const m4 ws2ms1 = tgo_m4_translate(tgo_v3_neg(p_object->translation));
    const m4 ws2ms2 = tgo_m4_inverse(rotation);
    const m4 ws2ms3 = tgo_m4_translate(object_half_extent);
    const m4 ws2ms4 = tgo_m4_translate(tgo_v3_neg(relative_cluster_offset));
    return tgo_m4_mul(tgo_m4_mul(tgo_m4_mul(ws2ms4, ws2ms3), ws2ms2), ws2ms1);
}

/* tg_sparse_voxel_octree.c:62-100 */
static m4 tgo__cs2ws_c(const tg_voxel_object* p_object, u32 relative_cluster_pointer)
{
    const u32 rx = relative_cluster_pointer % p_object->n_cluster_pointers_per_dim.x;
    const u32 ry = (relative_cluster_pointer / p_object->n_cluster_pointers_per_dim.x) % p_object->n_cluster_pointers_per_dim.y;
    const u32 rz = relative_cluster_pointer / (p_object->n_cluster_pointers_per_dim.x * p_object->n_cluster_pointers_per_dim.y);
    const v3 relative_cluster_offset = tgo_v3((f32)(rx * 8u), (f32)(ry * 8u), (f32)(rz * 8u));
    const f32 h = (f32)(8 / 2);
    const v3 cluster_half_extent = tgo_v3(h, h, h);
    const v3 dims = tgo_v3((f32)p_object->n_cluster_pointers_per_dim.x, (f32)p_object->n_cluster_pointers_per_dim.y, (f32)p_object->n_cluster_pointers_per_dim.z);
    const v3 object_half_extent = tgo_v3_mul(dims, cluster_half_extent);
    const m4 rotation = tgo_m4_angle_axis(p_object->angle_in_radians, p_object->axis);

    const m4 ms2ws1 = tgo_m4_translate(relative_cluster_offset);
    const m4 ms2ws2 = tgo_m4_translate(tgo_v3_neg(object_half_extent));
    const m4 ms2ws3 = rotation;
    const m4 ms2ws4 = tgo_m4_translate(p_object->translation);
    return tgo_m4_mul(tgo_m4_mul(tgo_m4_mul(ms2ws4, ms2ws3), ms2ws2), ms2ws1);
}

/* physics/tg_physics.c:226-392 */
b32 tgo_intersect_aabb_obb_ignore_contact(v3 bmin, v3 bmax, const v3* c)
{
    b32 separated;
    u32 i;

    separated = TG_TRUE; for (i = 0; i < 8 && separated; i++) separated &= c[i].x <= bmin.x; if (separated) return TG_FALSE;
    separated = TG_TRUE; for (i = 0; i < 8 && separated; i++) separated &= c[i].x >= bmax.x; if (separated) return TG_FALSE;
    separated = TG_TRUE; for (i = 0; i < 8 && separated; i++) separated &= c[i].y <= bmin.y; if (separated) return TG_FALSE;
    separated = TG_TRUE; for (i = 0; i < 8 && separated; i++) separated &= c[i].y >= bmax.y; if (separated) return TG_FALSE;
    separated = TG_TRUE; for (i = 0; i < 8 && separated; i++) separated &= c[i].z <= bmin.z; if (separated) return TG_FALSE;
    separated = TG_TRUE; for (i = 0; i < 8 && separated; i++) separated &= c[i].z >= bmax.z; if (separated) return TG_FALSE;

    /* OBB faces: tg_physics.c:332-391 */
    const v3 nxp = tgo_v3_normalized(tgo_v3_sub(c[1], c[0]));
    const v3 nyp = tgo_v3_normalized(tgo_v3_sub(c[2], c[0]));
    const v3 nzp = tgo_v3_normalized(tgo_v3_sub(c[4], c[0]));

    v3 normals[6];
    normals[0] = tgo_v3_neg(nxp); normals[1] = nxp;
    normals[2] = tgo_v3_neg(nyp); normals[3] = nyp;
    normals[4] = tgo_v3_neg(nzp); normals[5] = nzp;

    f32 distances[6];
    const f32 l0 = tgo_v3_mag(c[0]);
    const f32 l1 = tgo_v3_mag(c[7]);
    const v3 n0 = l0 == 0.0f ? tgo_v3(0, 0, 0) : tgo_v3_divf(c[0], l0);
    const v3 n1 = l1 == 0.0f ? tgo_v3(0, 0, 0) : tgo_v3_divf(c[7], l1);
    distances[0] = l0 * tgo_v3_dot(n0, normals[0]);
    distances[1] = l1 * tgo_v3_dot(n1, normals[1]);
    distances[2] = l0 * tgo_v3_dot(n0, normals[2]);
    distances[3] = l1 * tgo_v3_dot(n1, normals[3]);
    distances[4] = l0 * tgo_v3_dot(n0, normals[4]);
    distances[5] = l1 * tgo_v3_dot(n1, normals[5]);

    v3 pts[8];
    pts[0] = tgo_v3(bmin.x, bmin.y, bmin.z);
    pts[1] = tgo_v3(bmax.x, bmin.y, bmin.z);
    pts[2] = tgo_v3(bmin.x, bmax.y, bmin.z);
    pts[3] = tgo_v3(bmax.x, bmax.y, bmin.z);
    pts[4] = tgo_v3(bmin.x, bmin.y, bmax.z);
    pts[5] = tgo_v3(bmax.x, bmin.y, bmax.z);
    pts[6] = tgo_v3(bmin.x, bmax.y, bmax.z);
    pts[7] = tgo_v3(bmax.x, bmax.y, bmax.z);

    for (u32 plane = 0; plane < 6; plane++)
    {
        separated = TG_TRUE;
        for (u32 k = 0; k < 8; k++)
        {
            const f32 dist = tgo_v3_dot(normals[plane], pts[k]) - distances[plane];
            separated &= dist >= 0.0f;
            if (!separated) break;
        }
        if (separated) return TG_FALSE;
    }
    return TG_TRUE;
}

typedef struct tgo__svo_build
{
    tg_svo*                p_svo;
    const tgo_scene_view*  p_scene;
    const tg_voxel_object* p_objects;
} tgo__svo_build;

static void tgo__block_corners(v3 bmin, v3 bmax, v4* p)
{
    p[0] = (v4){ bmin.x, bmin.y, bmin.z, 1.0f };
    p[1] = (v4){ bmax.x, bmin.y, bmin.z, 1.0f };
    p[2] = (v4){ bmin.x, bmax.y, bmin.z, 1.0f };
    p[3] = (v4){ bmax.x, bmax.y, bmin.z, 1.0f };
    p[4] = (v4){ bmin.x, bmin.y, bmax.z, 1.0f };
    p[5] = (v4){ bmax.x, bmin.y, bmax.z, 1.0f };
    p[6] = (v4){ bmin.x, bmax.y, bmax.z, 1.0f };
    p[7] = (v4){ bmax.x, bmax.y, bmax.z, 1.0f };
}

static v3 tgo__xyz(v4 v) { return tgo_v3(v.x, v.y, v.z); }

/* tg_sparse_voxel_octree.c:102-317 */
static void tgo__construct_leaf_node(tgo__svo_build* b, tg_svo_leaf_node* p_parent, v3 parent_min, v3 parent_max, u32 n_cluster_pointers, const u32* p_cluster_pointers)
{
    tg_svo* p_svo = b->p_svo;
    if (p_svo->leaf_node_data_buffer_count >= p_svo->leaf_node_data_buffer_capacity ||
        p_svo->voxel_buffer_count_in_u32 + TG_SVO_BLOCK_WORDS > p_svo->voxel_buffer_capacity_in_u32)
    {
        abort(); /* TG_ASSERT at :116,:124; callers size the capacities */
    }
    tg_svo_leaf_node_data* p_data = &p_svo->p_leaf_node_data_buffer[p_svo->leaf_node_data_buffer_count];
    p_parent->data_pointer = p_svo->leaf_node_data_buffer_count++;

    const v3 parent_extent = tgo_v3_sub(parent_max, parent_min);
    const u32 voxel_count = (u32)parent_extent.x * (u32)parent_extent.y * (u32)parent_extent.z;
    const u32 voxels_in_u32 = (voxel_count + 31) / 32;
    u32* p_block_voxels = &p_svo->p_voxels_buffer[p_svo->voxel_buffer_count_in_u32];
    p_svo->voxel_buffer_count_in_u32 += voxels_in_u32;

    const f32 ext_c = 8.0f;
    v4 corners_cs[8];
    corners_cs[0] = (v4){  0.0f,  0.0f,  0.0f, 1.0f };
    corners_cs[1] = (v4){ ext_c,  0.0f,  0.0f, 1.0f };
    corners_cs[2] = (v4){  0.0f, ext_c,  0.0f, 1.0f };
    corners_cs[3] = (v4){ ext_c, ext_c,  0.0f, 1.0f };
    corners_cs[4] = (v4){  0.0f,  0.0f, ext_c, 1.0f };
    corners_cs[5] = (v4){ ext_c,  0.0f, ext_c, 1.0f };
    corners_cs[6] = (v4){  0.0f, ext_c, ext_c, 1.0f };
    corners_cs[7] = (v4){ ext_c, ext_c, ext_c, 1.0f };

    v4 block_ws[8];
    tgo__block_corners(parent_min, parent_max, block_ws);

    for (u32 k = 0; k < n_cluster_pointers; k++)
    {
        const u32 cluster_pointer = p_cluster_pointers[k];
        const u32 cluster_idx = b->p_scene->p_cluster_pointers[cluster_pointer];
        const u32 object_idx = b->p_scene->p_cluster_idx_to_object_idx[cluster_idx];
        const tg_voxel_object* p_object = &b->p_objects[object_idx];
        const u32 relative_cluster_pointer = cluster_pointer - p_object->first_cluster_pointer;

        const m4 cs2ws = tgo__cs2ws_c(p_object, relative_cluster_pointer);
        v3 cw[8];
        for (u32 i = 0; i < 8; i++) cw[i] = tgo__xyz(tgo_m4_mulv4(cs2ws, corners_cs[i]));

        /* :180-194 -- corner 3 is NOT used, corner 2 twice (Q4) */
        const v3 min_c = tgo_v3_min(
            tgo_v3_min(tgo_v3_min(cw[0], cw[1]), tgo_v3_min(cw[2], cw[2])),
            tgo_v3_min(tgo_v3_min(cw[4], cw[6]), tgo_v3_min(cw[5], cw[7])));
        const v3 max_c = tgo_v3_max(
            tgo_v3_max(tgo_v3_max(cw[0], cw[1]), tgo_v3_max(cw[2], cw[2])),
            tgo_v3_max(tgo_v3_max(cw[4], cw[6]), tgo_v3_max(cw[5], cw[7])));
        const v3 floor_min_c = tgo_v3_floor(min_c);
        const v3 ceil_max_c = tgo_v3_ceil(max_c);

        const m4 ws2cs = tgo_m4_inverse(cs2ws);
        v3 bc[8];
        for (u32 i = 0; i < 8; i++) bc[i] = tgo__xyz(tgo_m4_mulv4(ws2cs, block_ws[i]));

        const v3 trimmed_min_b = tgo_v3_max(parent_min, floor_min_c);
        const v3 trimmed_max_b = tgo_v3_min(parent_max, ceil_max_c);

        /*
         * :228-233. The reference asserts (:221-226) that the differences lie in [0, extent]; a
         * cluster whose AABB misses the block on some axis violates that (the SAT is conservative)
         * and the release build then converts a NEGATIVE float to u32 (undefined in C; on the
         * x64/MSVC target cvttss2si gives a huge value and the `<` loops below do not run).
         * Pinned here: a negative difference yields an empty range.
         */
        const f32 dminx = trimmed_min_b.x - parent_min.x, dminy = trimmed_min_b.y - parent_min.y, dminz = trimmed_min_b.z - parent_min.z;
        const f32 dmaxx = trimmed_max_b.x - parent_min.x, dmaxy = trimmed_max_b.y - parent_min.y, dmaxz = trimmed_max_b.z - parent_min.z;
        if (dminx < 0.0f || dminy < 0.0f || dminz < 0.0f || dmaxx < 0.0f || dmaxy < 0.0f || dmaxz < 0.0f) continue;
        const u32 min_x = (u32)dminx, min_y = (u32)dminy, min_z = (u32)dminz;
        const u32 max_x = (u32)dmaxx, max_y = (u32)dmaxy, max_z = (u32)dmaxz;

        const u32* p_mask = &b->p_scene->p_voxel_cluster_data[(size_t)cluster_idx * TG_CLUSTER_MASK_WORDS];

        for (u32 bz = min_z; bz < max_z; bz++)
        {
            const f32 tz = ((f32)bz + 0.5f) / parent_extent.z;
            const f32 omtz = 1.0f - tz;
            const f32 pz0_x = omtz * bc[0].x + tz * bc[4].x, pz0_y = omtz * bc[0].y + tz * bc[4].y, pz0_z = omtz * bc[0].z + tz * bc[4].z;
            const f32 pz1_x = omtz * bc[1].x + tz * bc[5].x, pz1_y = omtz * bc[1].y + tz * bc[5].y, pz1_z = omtz * bc[1].z + tz * bc[5].z;
            const f32 pz2_x = omtz * bc[2].x + tz * bc[6].x, pz2_y = omtz * bc[2].y + tz * bc[6].y, pz2_z = omtz * bc[2].z + tz * bc[6].z;
            const f32 pz3_x = omtz * bc[3].x + tz * bc[7].x, pz3_y = omtz * bc[3].y + tz * bc[7].y, pz3_z = omtz * bc[3].z + tz * bc[7].z;
            for (u32 by = min_y; by < max_y; by++)
            {
                const f32 ty = ((f32)by + 0.5f) / parent_extent.y;
                const f32 omty = 1.0f - ty;
                const f32 py0_x = omty * pz0_x + ty * pz2_x, py0_y = omty * pz0_y + ty * pz2_y, py0_z = omty * pz0_z + ty * pz2_z;
                const f32 py1_x = omty * pz1_x + ty * pz3_x, py1_y = omty * pz1_y + ty * pz3_y, py1_z = omty * pz1_z + ty * pz3_z;
                for (u32 bx = min_x; bx < max_x; bx++)
                {
                    const f32 tx = ((f32)bx + 0.5f) / parent_extent.x;
                    const f32 omtx = 1.0f - tx;
                    const f32 cx = omtx * py0_x + tx * py1_x;
                    if (cx < 0.0f || cx >= 8.0f) continue;
                    const f32 cy = omtx * py0_y + tx * py1_y;
                    if (cy < 0.0f || cy >= 8.0f) continue;
                    const f32 cz = omtx * py0_z + tx * py1_z;
                    if (cz < 0.0f || cz >= 8.0f) continue;

                    const u32 rel_voxel = 64u * (u32)cz + 8u * (u32)cy + (u32)cx;
                    if ((p_mask[rel_voxel / 32] & (1u << (rel_voxel % 32))) != 0)
                    {
                        const u32 block_voxel_idx = (u32)(parent_extent.x * parent_extent.y * bz + parent_extent.x * by + bx);
                        p_block_voxels[block_voxel_idx / 32] |= 1u << (block_voxel_idx % 32);
                        if (p_data->n == 0 || p_data->p_cluster_idcs[p_data->n - 1] != cluster_idx)
                        {
                            if (p_data->n < TG_SVO_LEAF_MAX_CLUSTERS) /* defined deviation, see header */
                            {
                                p_data->p_cluster_idcs[p_data->n++] = cluster_idx;
                            }
                        }
                    }
                }
            }
        }
    }
}

/* tg_sparse_voxel_octree.c:319-464 */
static void tgo__construct_inner_node(tgo__svo_build* b, u32 parent_node_idx, v3 parent_min, v3 parent_max, u32 n_cluster_pointers, const u32* p_cluster_pointers)
{
    tg_svo* p_svo = b->p_svo;
    u8 valid_mask = 0;
    u32 n_per_child[8] = { 0 };
    u32* p_per_child = (u32*)malloc((size_t)8 * (n_cluster_pointers ? n_cluster_pointers : 1) * sizeof(u32));

    const v3 child_extent = tgo_v3_mulf(tgo_v3_sub(parent_max, parent_min), 0.5f);

    for (u32 k = 0; k < n_cluster_pointers; k++)
    {
        const u32 cluster_pointer = p_cluster_pointers[k];
        const u32 cluster_idx = b->p_scene->p_cluster_pointers[cluster_pointer];
        const u32 object_idx = b->p_scene->p_cluster_idx_to_object_idx[cluster_idx];
        const tg_voxel_object* p_object = &b->p_objects[object_idx];
        const u32 relative_cluster_pointer = cluster_pointer - p_object->first_cluster_pointer;
        const m4 ws2ms = tgo__ws2ms_c(p_object, relative_cluster_pointer);

        for (u32 child_idx = 0; child_idx < 8; child_idx++)
        {
            const f32 d_x = (f32)( child_idx      % 2) * child_extent.x;
            const f32 d_y = (f32)((child_idx / 2) % 2) * child_extent.y;
            const f32 d_z = (f32)((child_idx / 4) % 2) * child_extent.z;
            const v3 child_min = tgo_v3_add(parent_min, tgo_v3(d_x, d_y, d_z));
            const v3 child_max = tgo_v3_add(child_min, child_extent);

            v4 block_ws[8];
            tgo__block_corners(child_min, child_max, block_ws);
            v3 corners[8];
            for (u32 i = 0; i < 8; i++) corners[i] = tgo__xyz(tgo_m4_mulv4(ws2ms, block_ws[i]));

            if (tgo_intersect_aabb_obb_ignore_contact(tgo_v3(0, 0, 0), tgo_v3(8, 8, 8), corners))
            {
                valid_mask |= (u8)(1u << child_idx);
                if (n_per_child[child_idx] == 0 || p_per_child[(size_t)child_idx * n_cluster_pointers + n_per_child[child_idx] - 1] != cluster_pointer)
                {
                    p_per_child[(size_t)child_idx * n_cluster_pointers + n_per_child[child_idx]++] = cluster_pointer;
                }
            }
        }
    }

    if (valid_mask)
    {
        const u32 first_child_idx = p_svo->node_buffer_count;
        const u32 child_count = (u32)__builtin_popcount(valid_mask);
        if (first_child_idx + child_count > p_svo->node_buffer_capacity || first_child_idx - parent_node_idx >= 0xFFFFu) abort(); /* :408,:413 */
        p_svo->p_node_buffer[parent_node_idx].inner.valid_mask = valid_mask;
        p_svo->p_node_buffer[parent_node_idx].inner.child_pointer = (u16)(first_child_idx - parent_node_idx);
        p_svo->node_buffer_count += child_count;

        const u32 child_voxel_count = (u32)child_extent.x * (u32)child_extent.y * (u32)child_extent.z;
        const b32 children_are_leaves = child_voxel_count == TG_SVO_BLOCK_VOXEL_COUNT;

        u32 child_node_offset = 0;
        for (u32 child_idx = 0; child_idx < 8; child_idx++)
        {
            if ((valid_mask & (1u << child_idx)) == 0) continue;
            const f32 dx = (f32)( child_idx      % 2) * child_extent.x;
            const f32 dy = (f32)((child_idx / 2) % 2) * child_extent.y;
            const f32 dz = (f32)((child_idx / 4) % 2) * child_extent.z;
            const v3 child_min = tgo_v3_add(parent_min, tgo_v3(dx, dy, dz));
            const v3 child_max = tgo_v3_add(child_min, child_extent);
            if (!children_are_leaves)
            {
                tgo__construct_inner_node(b, first_child_idx + child_node_offset, child_min, child_max,
                                          n_per_child[child_idx], &p_per_child[(size_t)child_idx * n_cluster_pointers]);
            }
            else
            {
                p_svo->p_node_buffer[parent_node_idx].inner.leaf_mask |= (u8)(1u << child_idx);
                tgo__construct_leaf_node(b, &p_svo->p_node_buffer[first_child_idx + child_node_offset].leaf, child_min, child_max,
                                         n_per_child[child_idx], &p_per_child[(size_t)child_idx * n_cluster_pointers]);
            }
            child_node_offset++;
        }
    }
    free(p_per_child);
}

/* tg_sparse_voxel_octree.c:466-542. Capacities: caller may preset non-zero capacities in *p_svo, else the reference's (:479-484). */
void tgo_svo_create(v3 extent_min, v3 extent_max, const tgo_scene_view* p_scene, const tg_voxel_object* p_objects, tg_svo* p_svo)
{
    const u32 cap_voxels = p_svo->voxel_buffer_capacity_in_u32 ? p_svo->voxel_buffer_capacity_in_u32 : (1u << 21);
    const u32 cap_leaves = p_svo->leaf_node_data_buffer_capacity ? p_svo->leaf_node_data_buffer_capacity : (1u << 13);
    const u32 cap_nodes  = p_svo->node_buffer_capacity ? p_svo->node_buffer_capacity : (1u << 14);
    memset(p_svo, 0, sizeof(*p_svo));
    p_svo->min = extent_min;
    p_svo->max = extent_max;
    p_svo->voxel_buffer_capacity_in_u32 = cap_voxels;
    p_svo->leaf_node_data_buffer_capacity = cap_leaves;
    p_svo->node_buffer_capacity = cap_nodes;
    /* tgp_malloc zero-fills (platform/tg_platform_win32.c:133); the builder relies on it */
    p_svo->p_voxels_buffer = (u32*)calloc(cap_voxels, sizeof(u32));
    p_svo->p_leaf_node_data_buffer = (tg_svo_leaf_node_data*)calloc(cap_leaves, sizeof(tg_svo_leaf_node_data));
    p_svo->p_node_buffer = (tg_svo_node*)calloc(cap_nodes, sizeof(tg_svo_node));
    p_svo->node_buffer_count = 1; /* root */

    /* :496-537: all pointers in object order whose mask is non-zero */
    u32 n = 0;
    u32* p_list = (u32*)malloc((size_t)(p_scene->n_cluster_pointers ? p_scene->n_cluster_pointers : 1) * sizeof(u32));
    u32 curr = 0;
    while (curr < p_scene->n_cluster_pointers)
    {
        const u32 object_idx = p_scene->p_cluster_idx_to_object_idx[p_scene->p_cluster_pointers[curr]];
        const tg_voxel_object* p_object = &p_objects[object_idx];
        const u32 n_of_object = p_object->n_cluster_pointers_per_dim.x * p_object->n_cluster_pointers_per_dim.y * p_object->n_cluster_pointers_per_dim.z;
        curr += n_of_object;
        for (u32 rel = 0; rel < n_of_object; rel++)
        {
            const u32 cluster_pointer = p_object->first_cluster_pointer + rel;
            const u32 cluster_idx = p_scene->p_cluster_pointers[cluster_pointer];
            const u32* p_cluster = &p_scene->p_voxel_cluster_data[(size_t)cluster_idx * TG_CLUSTER_MASK_WORDS];
            b32 contains_voxels = TG_FALSE;
            for (u32 v = 0; v < TG_N_PRIMITIVES_PER_CLUSTER; v++)
            {
                if (p_cluster[v / 32] & (1u << (v % 32))) { contains_voxels = TG_TRUE; break; }
            }
            if (!contains_voxels) continue;
            p_list[n++] = cluster_pointer;
        }
    }

    tgo__svo_build b = { p_svo, p_scene, p_objects };
    tgo__construct_inner_node(&b, 0, extent_min, extent_max, n, p_list);
    free(p_list);
}

void tgo_svo_destroy(tg_svo* p_svo)
{
    free(p_svo->p_voxels_buffer);
    free(p_svo->p_leaf_node_data_buffer);
    free(p_svo->p_node_buffer);
    memset(p_svo, 0, sizeof(*p_svo));
}

/* util/tg_amanatides_woo.c:3-114 (no lower clamp on the start cell, no `enter` term) */
b32 tgo_amanatides_woo(v3 ray_hit_on_grid, v3 ray_direction, v3 extent, const u32* p_voxel_grid, v3i* p_voxel_id)
{
    p_voxel_id->x = -1; p_voxel_id->y = -1; p_voxel_id->z = -1;
    const v3 extent_minus_one = tgo_v3_sub(extent, tgo_v3(1.0f, 1.0f, 1.0f));
    const v3 xyz = tgo_v3_min(tgo_v3_floor(ray_hit_on_grid), extent_minus_one);
    i32 x = (i32)xyz.x, y = (i32)xyz.y, z = (i32)xyz.z;
    i32 step_x = 0, step_y = 0, step_z = 0;
    f32 t_max_x = TG_F32_MAX, t_max_y = TG_F32_MAX, t_max_z = TG_F32_MAX;
    f32 t_delta_x = TG_F32_MAX, t_delta_y = TG_F32_MAX, t_delta_z = TG_F32_MAX;

    if (ray_direction.x > 0.0f)      { step_x = 1;  t_max_x = ((f32)(x + 1) - ray_hit_on_grid.x) / ray_direction.x;  t_delta_x = 1.0f / ray_direction.x; }
    else if (ray_direction.x < 0.0f) { step_x = -1; t_max_x = (ray_hit_on_grid.x - (f32)x) / -ray_direction.x;       t_delta_x = 1.0f / -ray_direction.x; }
    if (ray_direction.y > 0.0f)      { step_y = 1;  t_max_y = ((f32)(y + 1) - ray_hit_on_grid.y) / ray_direction.y;  t_delta_y = 1.0f / ray_direction.y; }
    else if (ray_direction.y < 0.0f) { step_y = -1; t_max_y = (ray_hit_on_grid.y - (f32)y) / -ray_direction.y;       t_delta_y = 1.0f / -ray_direction.y; }
    if (ray_direction.z > 0.0f)      { step_z = 1;  t_max_z = ((f32)(z + 1) - ray_hit_on_grid.z) / ray_direction.z;  t_delta_z = 1.0f / ray_direction.z; }
    else if (ray_direction.z < 0.0f) { step_z = -1; t_max_z = (ray_hit_on_grid.z - (f32)z) / -ray_direction.z;       t_delta_z = 1.0f / -ray_direction.z; }

    for (;;)
    {
        const u32 vox_idx = (u32)extent.x * (u32)extent.y * (u32)z + (u32)extent.x * (u32)y + (u32)x;
        if ((p_voxel_grid[vox_idx / 32] & (1u << (vox_idx % 32))) != 0)
        {
            p_voxel_id->x = x; p_voxel_id->y = y; p_voxel_id->z = z;
            return TG_TRUE;
        }
        if (t_max_x < t_max_y)
        {
            if (t_max_x < t_max_z) { t_max_x += t_delta_x; x += step_x; if (x < 0 || (f32)x >= extent.x) break; }
            else                   { t_max_z += t_delta_z; z += step_z; if (z < 0 || (f32)z >= extent.z) break; }
        }
        else
        {
            if (t_max_y < t_max_z) { t_max_y += t_delta_y; y += step_y; if (y < 0 || (f32)y >= extent.y) break; }
            else                   { t_max_z += t_delta_z; z += step_z; if (z < 0 || (f32)z >= extent.z) break; }
        }
    }
    return TG_FALSE;
}

static inline f32 tgo__exit_c(v3 bmin, v3 bmax, v3 position, v3 d)
{
    /* tg_sparse_voxel_octree.c:704-707: div_zero_check + C max + tgm_f32_min */
    v3 a, b2;
    a.x  = (d.x == 0.0f) ? TG_F32_MIN : ((bmin.x - position.x) / d.x);
    a.y  = (d.y == 0.0f) ? TG_F32_MIN : ((bmin.y - position.y) / d.y);
    a.z  = (d.z == 0.0f) ? TG_F32_MIN : ((bmin.z - position.z) / d.z);
    b2.x = (d.x == 0.0f) ? TG_F32_MAX : ((bmax.x - position.x) / d.x);
    b2.y = (d.y == 0.0f) ? TG_F32_MAX : ((bmax.y - position.y) / d.y);
    b2.z = (d.z == 0.0f) ? TG_F32_MAX : ((bmax.z - position.z) / d.z);
    const v3 f = tgo_v3_max(a, b2);
    const f32 m = f.x < f.y ? f.x : f.y;
    return m < f.z ? m : f.z;
}

/* tg_sparse_voxel_octree.c:558-740 */
b32 tgo_svo_traverse_c(const tg_svo* p_svo, v3 ray_origin, v3 ray_direction, f32* p_distance, u32* p_node_idx, u32* p_voxel_idx)
{
    *p_distance = TG_F32_MAX;
    *p_node_idx = TG_U32_MAX;
    *p_voxel_idx = TG_U32_MAX;

    const v3 extent = tgo_v3_sub(p_svo->max, p_svo->min);
    const v3 center = tgo_v3_add(p_svo->min, tgo_v3_mulf(extent, 0.5f));
    ray_origin = tgo_v3_sub(ray_origin, center);

    u32 stack_size;
    u32 idx_stack[TG_SVO_TRAVERSE_STACK_CAPACITY] = { 0 };
    v3  min_stack[TG_SVO_TRAVERSE_STACK_CAPACITY];
    v3  max_stack[TG_SVO_TRAVERSE_STACK_CAPACITY];

    f32 enter, exit;
    f32 distance_ray_origin_2_position = 0.0f;
    if (!tgo_intersect_ray_aabb_c(ray_origin, ray_direction, p_svo->min, p_svo->max, &enter, &exit)) return TG_FALSE;

    v3 position = enter > 0.0f ? tgo_v3_add(ray_origin, tgo_v3_mulf(ray_direction, enter)) : ray_origin;
    distance_ray_origin_2_position = enter > 0.0f ? enter : 0.0f;

    stack_size = 1;
    idx_stack[0] = 0;
    min_stack[0] = p_svo->min;
    max_stack[0] = p_svo->max;

    u32 iterations = 0;
    while (stack_size > 0)
    {
        if (++iterations > TGO_TRAVERSE_MAX_ITERATIONS) return TG_FALSE;
        const u32 parent_idx = idx_stack[stack_size - 1];
        const v3 parent_min = min_stack[stack_size - 1];
        const v3 parent_max = max_stack[stack_size - 1];
        const tg_svo_inner_node* p_parent_node = &p_svo->p_node_buffer[parent_idx].inner;
        const u32 child_pointer = p_parent_node->child_pointer;
        const u32 valid_mask = p_parent_node->valid_mask;
        const u32 leaf_mask = p_parent_node->leaf_mask;

        const v3 child_extent = tgo_v3_mulf(tgo_v3_sub(parent_max, parent_min), 0.5f);
        u32 relative_child_idx = 0;
        v3 child_min = parent_min;
        v3 child_max = tgo_v3_add(child_min, child_extent);
        if (child_max.x < position.x || (position.x == child_max.x && ray_direction.x > 0.0f)) { relative_child_idx += 1; child_min.x += child_extent.x; child_max.x += child_extent.x; }
        if (child_max.y < position.y || (position.y == child_max.y && ray_direction.y > 0.0f)) { relative_child_idx += 2; child_min.y += child_extent.y; child_max.y += child_extent.y; }
        if (child_max.z < position.z || (position.z == child_max.z && ray_direction.z > 0.0f)) { relative_child_idx += 4; child_min.z += child_extent.z; child_max.z += child_extent.z; }

        b32 advance_to_border = TG_TRUE;
        if ((valid_mask & (1u << relative_child_idx)) != 0)
        {
            u32 relative_child_offset = 0;
            for (u32 i = 0; i < relative_child_idx; i++) relative_child_offset += (valid_mask >> i) & 1;
            const u32 child_idx = parent_idx + child_pointer + relative_child_offset;

            if ((leaf_mask & (1u << relative_child_idx)) != 0)
            {
                const tg_svo_node* p_child_node = &p_svo->p_node_buffer[child_idx];
                const tg_svo_leaf_node_data* p_data = &p_svo->p_leaf_node_data_buffer[p_child_node->leaf.data_pointer];
                if (p_data->n != 0)
                {
                    const u32 first_voxel_id = p_child_node->leaf.data_pointer * TG_SVO_BLOCK_VOXEL_COUNT;
                    const v3 ray_hit_on_grid = tgo_v3_sub(position, child_min);
                    const u32* p_voxel_grid = p_svo->p_voxels_buffer + first_voxel_id / 32;
                    v3i voxel_id;
                    if (tgo_amanatides_woo(ray_hit_on_grid, ray_direction, child_extent, p_voxel_grid, &voxel_id))
                    {
                        const v3 voxel_min = tgo_v3_add(child_min, tgo_v3((f32)voxel_id.x, (f32)voxel_id.y, (f32)voxel_id.z));
                        const v3 voxel_max = tgo_v3_add(voxel_min, tgo_v3(1.0f, 1.0f, 1.0f));
                        tgo_intersect_ray_aabb_c(position, ray_direction, voxel_min, voxel_max, &enter, &exit);
                        const u32 relative_voxel_idx = (u32)child_extent.x * (u32)child_extent.y * (u32)voxel_id.z + (u32)child_extent.x * (u32)voxel_id.y + (u32)voxel_id.x;
                        *p_distance = distance_ray_origin_2_position + enter;
                        *p_node_idx = child_idx;
                        *p_voxel_idx = first_voxel_id + relative_voxel_idx;
                        return TG_TRUE;
                    }
                }
            }
            else
            {
                advance_to_border = TG_FALSE;
                idx_stack[stack_size] = child_idx;
                min_stack[stack_size] = child_min;
                max_stack[stack_size] = child_max;
                stack_size++;
            }
        }

        if (advance_to_border)
        {
            exit = tgo__exit_c(child_min, child_max, position, ray_direction);
            const f32 t_advance = exit + TG_F32_EPSILON;
            position = tgo_v3_add(position, tgo_v3_mulf(ray_direction, t_advance));
            distance_ray_origin_2_position += t_advance;
            while (stack_size > 0)
            {
                exit = tgo__exit_c(min_stack[stack_size - 1], max_stack[stack_size - 1], position, ray_direction);
                if (exit > TG_F32_EPSILON) break;
                stack_size--;
            }
        }
    }
    return TG_FALSE;
}

static inline f32 tgo__exit_glsl(v3 bmin, v3 bmax, v3 position, v3 d)
{
    /* svo_functions.inc:283-292 */
    const f32 ax = (d.x == 0.0f) ? TG_F32_MIN : ((bmin.x - position.x) / d.x);
    const f32 ay = (d.y == 0.0f) ? TG_F32_MIN : ((bmin.y - position.y) / d.y);
    const f32 az = (d.z == 0.0f) ? TG_F32_MIN : ((bmin.z - position.z) / d.z);
    const f32 bx = (d.x == 0.0f) ? TG_F32_MAX : ((bmax.x - position.x) / d.x);
    const f32 by = (d.y == 0.0f) ? TG_F32_MAX : ((bmax.y - position.y) / d.y);
    const f32 bz = (d.z == 0.0f) ? TG_F32_MAX : ((bmax.z - position.z) / d.z);
    const f32 fx = tgo_max(ax, bx), fy = tgo_max(ay, by), fz = tgo_max(az, bz);
    return tgo_min(tgo_min(fx, fy), fz);
}

/* svo_functions.inc:1-329 */
f32 tgo_svo_traverse_glsl(const tg_svo* p_svo, f32 far_plane, v3 ray_origin_ws, v3 ray_direction_ws, v3* p_hit_position, v3* p_hit_normal, u32* p_node_idx, u32* p_voxel_idx)
{
    const v3 extent = tgo_v3_sub(p_svo->max, p_svo->min);
    const v3 center = tgo_v3_add(tgo_v3_mulf(extent, 0.5f), p_svo->min);
    const v3 o = tgo_v3_sub(ray_origin_ws, center);
    const v3 d = ray_direction_ws;

    u32 stack_size;
    u32 idx_stack[TG_SVO_TRAVERSE_STACK_CAPACITY] = { 0 };
    v3  min_stack[TG_SVO_TRAVERSE_STACK_CAPACITY];
    v3  max_stack[TG_SVO_TRAVERSE_STACK_CAPACITY];

    f32 result = 1.0f;
    *p_node_idx = TG_U32_MAX;
    *p_voxel_idx = TG_U32_MAX;
    *p_hit_position = tgo_v3(0, 0, 0);
    *p_hit_normal = tgo_v3(0, 0, 0);

    f32 enter, exit;
    if (!tgo_intersect_ray_aabb_glsl(o, d, p_svo->min, p_svo->max, &enter, &exit)) return result;

    v3 position = o;
    if (enter > 0.0f) position = tgo_v3_add(position, tgo_v3_mulf(d, enter));

    stack_size = 1;
    idx_stack[0] = 0;
    min_stack[0] = p_svo->min;
    max_stack[0] = p_svo->max;

    u32 iterations = 0;
    while (stack_size > 0)
    {
        if (++iterations > TGO_TRAVERSE_MAX_ITERATIONS) return 1.0f;
        const u32 parent_idx = idx_stack[stack_size - 1];
        const v3 parent_min = min_stack[stack_size - 1];
        const v3 parent_max = max_stack[stack_size - 1];
        u32 node_data;
        memcpy(&node_data, &p_svo->p_node_buffer[parent_idx], 4);
        const u32 child_pointer =  node_data        & 0xFFFFu;
        const u32 valid_mask    = (node_data >> 16) & 0xFFu;
        const u32 leaf_mask     = (node_data >> 24) & 0xFFu;

        const v3 child_extent = tgo_v3_mulf(tgo_v3_sub(parent_max, parent_min), 0.5f);
        u32 relative_child_idx = 0;
        v3 child_min = parent_min;
        v3 child_max = tgo_v3_add(child_min, child_extent);
        if (child_max.x < position.x || (position.x == child_max.x && d.x > 0.0f)) { relative_child_idx += 1; child_min.x += child_extent.x; child_max.x += child_extent.x; }
        if (child_max.y < position.y || (position.y == child_max.y && d.y > 0.0f)) { relative_child_idx += 2; child_min.y += child_extent.y; child_max.y += child_extent.y; }
        if (child_max.z < position.z || (position.z == child_max.z && d.z > 0.0f)) { relative_child_idx += 4; child_min.z += child_extent.z; child_max.z += child_extent.z; }

        b32 advance_to_border = TG_TRUE;
        if ((valid_mask & (1u << relative_child_idx)) != 0)
        {
            u32 relative_child_offset = 0;
            for (u32 i = 0; i < relative_child_idx; i++) relative_child_offset += (valid_mask >> i) & 1;
            const u32 child_idx = parent_idx + child_pointer + relative_child_offset;

            if ((leaf_mask & (1u << relative_child_idx)) != 0)
            {
                const u32 data_pointer = p_svo->p_node_buffer[child_idx].leaf.data_pointer;
                if (p_svo->p_leaf_node_data_buffer[data_pointer].n != 0)
                {
                    const u32 first_voxel_idx = data_pointer * TG_SVO_BLOCK_VOXEL_COUNT;
                    v3 hit = position;
                    const v3 fl = tgo_v3_floor(hit);
                    v3 xyz = tgo_v3(
                        tgo_clamp(fl.x, child_min.x, child_max.x - 1.0f),
                        tgo_clamp(fl.y, child_min.y, child_max.y - 1.0f),
                        tgo_clamp(fl.z, child_min.z, child_max.z - 1.0f));
                    hit = tgo_v3_sub(hit, child_min);
                    xyz = tgo_v3_sub(xyz, child_min);
                    i32 x = (i32)xyz.x, y = (i32)xyz.y, z = (i32)xyz.z;
                    i32 step_x = 0, step_y = 0, step_z = 0;
                    f32 t_max_x = TG_F32_MAX, t_max_y = TG_F32_MAX, t_max_z = TG_F32_MAX;
                    f32 t_delta_x = TG_F32_MAX, t_delta_y = TG_F32_MAX, t_delta_z = TG_F32_MAX;
                    if (d.x > 0.0f)      { step_x = 1;  t_max_x = ((f32)(x + 1) - hit.x) / d.x;  t_delta_x = 1.0f / d.x; }
                    else if (d.x < 0.0f) { step_x = -1; t_max_x = (hit.x - (f32)x) / -d.x;       t_delta_x = 1.0f / -d.x; }
                    if (d.y > 0.0f)      { step_y = 1;  t_max_y = ((f32)(y + 1) - hit.y) / d.y;  t_delta_y = 1.0f / d.y; }
                    else if (d.y < 0.0f) { step_y = -1; t_max_y = (hit.y - (f32)y) / -d.y;       t_delta_y = 1.0f / -d.y; }
                    if (d.z > 0.0f)      { step_z = 1;  t_max_z = ((f32)(z + 1) - hit.z) / d.z;  t_delta_z = 1.0f / d.z; }
                    else if (d.z < 0.0f) { step_z = -1; t_max_z = (hit.z - (f32)z) / -d.z;       t_delta_z = 1.0f / -d.z; }

                    const u32 ex = (u32)child_extent.x, ey = (u32)child_extent.y;
                    for (;;)
                    {
                        const u32 relative_voxel_idx = ex * ey * (u32)z + ex * (u32)y + (u32)x;
                        const u32 voxel_idx = first_voxel_idx + relative_voxel_idx;
                        const u32 bits = p_svo->p_voxels_buffer[voxel_idx / 32];
                        if ((bits & (1u << (voxel_idx % 32))) != 0)
                        {
                            const v3 voxel_min = tgo_v3_add(child_min, tgo_v3((f32)x, (f32)y, (f32)z));
                            const v3 voxel_max = tgo_v3_add(child_min, tgo_v3((f32)(x + 1), (f32)(y + 1), (f32)(z + 1)));
                            tgo_intersect_ray_aabb_glsl(o, d, voxel_min, voxel_max, &enter, &exit);

                            const v3 hit_position = tgo_v3_add(position, tgo_v3_mulf(d, enter)); /* :187, mixes frames (Q7) */
                            const v3 voxel_center = tgo_v3_add(voxel_min, tgo_v3(0.5f, 0.5f, 0.5f));
                            v3 n = tgo_v3_sub(hit_position, voxel_center);
                            if (fabsf(n.x) > fabsf(n.y))
                            {
                                n.y = 0.0f;
                                if (fabsf(n.x) > fabsf(n.z)) { n.x = tgo_sign(n.x); n.z = 0.0f; }
                                else                         { n.z = tgo_sign(n.z); n.x = 0.0f; }
                            }
                            else
                            {
                                n.x = 0.0f;
                                if (fabsf(n.y) > fabsf(n.z)) { n.y = tgo_sign(n.y); n.z = 0.0f; }
                                else                         { n.z = tgo_sign(n.z); n.y = 0.0f; }
                            }
                            result = enter / far_plane;
                            *p_hit_position = hit_position;
                            *p_hit_normal = n;
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
                advance_to_border = TG_FALSE;
                idx_stack[stack_size] = child_idx;
                min_stack[stack_size] = child_min;
                max_stack[stack_size] = child_max;
                stack_size++;
            }
        }

        if (advance_to_border)
        {
            exit = tgo__exit_glsl(child_min, child_max, position, d);
            position = tgo_v3_add(position, tgo_v3_mulf(d, exit + TG_F32_EPSILON));
            while (stack_size > 0)
            {
                exit = tgo__exit_glsl(min_stack[stack_size - 1], max_stack[stack_size - 1], position, d);
                if (exit > TG_F32_EPSILON) break;
                stack_size--;
            }
        }
    }
    return result;
}

/*
 * debug_visibility_svo.frag:27-71, the primary-ray pass the reference dispatches INSTEAD of the cluster pass while the BLOCKS
 * debug view is on (tgvk_raytracer.c:1226-1272): one full-screen fragment per pixel traverses the SVO from the camera with the
 * UN-normalised pixel direction and writes  depth24(d) << 40 | (node_idx & 0x7FFFFFFF) << 9 | voxel_idx % 512  with atomicMin
 * whenever d <= 1 (:52-71). A miss returns d = 1, node = voxel = 0xFFFFFFFF: the word is then all ones, the clear value.
 * Defined deviation: `u64(d * 16777215.0)` of a negative d (camera inside a solid voxel) is undefined in GLSL; both sides
 * convert like CUDA's cvt.rzi.u64.f32 (negative and NaN -> 0).
 * Rows y0, y0 + ystep, ... < y1; the other rows keep the clear value.
 */
void tgo_visibility_svo(const tg_svo* p_svo, const tg_camera_rays* p_cam, u32 w, u32 h, u32 y0, u32 y1, u32 ystep, u64* p_out)
{
    if (y1 > h) y1 = h;
    if (ystep == 0) ystep = 1;
    for (size_t i = 0; i < (size_t)w * h; i++) p_out[i] = TG_VIS_CLEAR; /* clear.comp:19 */
    const v3 camera = tgo_v3(p_cam->camera.x, p_cam->camera.y, p_cam->camera.z);
#pragma omp parallel for schedule(dynamic, 4)
    for (i64 py = (i64)y0; py < (i64)y1; py += ystep)
    {
        for (u32 px = 0; px < w; px++)
        {
            const v3 dir = tgo_pixel_ray_direction_nn(p_cam, w, h, px, (u32)py);
            v3 hp, hn; u32 node_idx, voxel_idx;
            const f32 d = tgo_svo_traverse_glsl(p_svo, p_cam->far_plane, camera, dir, &hp, &hn, &node_idx, &voxel_idx);
            if (d <= 1.0f)
            {
                const f32 dq = d * TG_VIS_DEPTH_SCALE;
                const u64 depth_24b = dq > 0.0f ? (u64)dq : 0; /* NaN compares false -> 0 */
                const u64 word = (depth_24b << TG_VIS_DEPTH_SHIFT) | ((u64)(node_idx & 2147483647u) << TG_VIS_POINTER_SHIFT) | (u64)(voxel_idx % 512u);
                u64* p = &p_out[(size_t)py * w + px];
                if (word < *p) *p = word; /* atomicMin on the cleared word */
            }
        }
    }
}

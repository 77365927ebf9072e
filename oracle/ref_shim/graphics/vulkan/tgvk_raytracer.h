/* ORACLE -- TEST INFRASTRUCTURE. Stand-in for /root/reference/tg/src/graphics/vulkan/tgvk_raytracer.h when
 * graphics/tg_sparse_voxel_octree.c (which includes it "// TODO: this is awful!", tg_sparse_voxel_octree.c:7) is compiled
 * without Vulkan: the one type the SVO builder reads, tg_scene, with the member order of tgvk_raytracer.h:90-108, and the
 * one function it calls, tg_object_is_initialized (tgvk_raytracer.h:246, body restated in oracle/ref_shim.c). */
#ifndef TGVK_RAYTRACER_H
#define TGVK_RAYTRACER_H

#include "graphics/tg_sparse_voxel_octree.h"

typedef struct tg_scene
{
    u32                 object_capacity;
    u32                 n_objects;
    tg_voxel_object*    p_objects;
    u32                 n_available_object_indices;
    u32*                p_available_object_indices;

    u32                 cluster_pointer_capacity;
    u32                 n_cluster_pointers;
    u32*                p_cluster_pointers;
    u32                 n_available_cluster_indices;
    u32*                p_available_cluster_indices;

    u32*                p_voxel_cluster_data;
    u32*                p_cluster_idx_to_object_idx;

    tg_svo              svo;
} tg_scene;

b32 tg_object_is_initialized(const tg_scene* p_scene, u32 object_idx);

#endif

/* tgb_svo.cu -- K2 placeholder (filled in next milestone). */
#include "tgb_device.cuh"

extern "C" b32 tgbd_svo_build(struct tgb_device* d, v3 extent_min, v3 extent_max, u32 n_cluster_pointers, u32 object_capacity)
{
    tgb_set_error("tgbd_svo_build: not built yet");
    return TG_FALSE;
}
extern "C" b32 tgbd_svo_update_objects(struct tgb_device* d, u32 n_moved, const u32* p_object_indices, const tg_object_data* p_old_records)
{
    tgb_set_error("tgbd_svo_update_objects: not built yet");
    return TG_FALSE;
}
extern "C" b32 tgbd_svo_counts(struct tgb_device* d, u32* p_n_nodes, u32* p_n_leaves, u32* p_n_voxel_words, v3* p_min, v3* p_max)
{
    tgb_set_error("tgbd_svo_counts: not built yet");
    return TG_FALSE;
}
extern "C" b32 tgbd_svo_set(struct tgb_device* d, v3 bmin, v3 bmax, u32 n_nodes, const void* p_nodes, u32 n_leaves, const void* p_leaf_data, u32 n_voxel_words, const void* p_voxels)
{
    tgb_set_error("tgbd_svo_set: not built yet");
    return TG_FALSE;
}

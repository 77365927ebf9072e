/* tgb_host_extra.c -- placeholders (filled in a later milestone). */
#include "tgb_internal.h"
#include "tgb_math.h"
#include "tgb_hoist.h"

void tgb200_debug_cluster_ray(const tg_object_data* p_object, const tg_camera_rays* p_cam, u32 width, u32 height, u32 px, u32 py,
                              u32 cluster_pointer, v3* p_origin_ms, v3* p_direction_ms)
{
    tgb_object_frame f;
    tgb_hoist_object(p_object, tgb_v3(p_cam->camera.x, p_cam->camera.y, p_cam->camera.z), &f);
    /* visibility.frag:35-38 */
    const u32 rel = cluster_pointer - p_object->first_cluster_pointer;
    const u32 cx = rel % f.nx, cy = (rel / f.nx) % f.ny, cz = rel / (f.nx * f.ny);
    *p_origin_ms = tgb_hoist_cluster_origin(&f, cx, cy, cz);
    *p_direction_ms = tgb_hoist_direction(&f, tgb_pixel_direction(p_cam, width, height, px, py));
}

void tgb200_procedural_solid_bits(u32 object_idx, v3u n_cluster_pointers_per_dim, u32* p_out)
{
    (void)object_idx; (void)n_cluster_pointers_per_dim; (void)p_out;
    tgb_set_error("tgb200_procedural_solid_bits: not built yet");
}

b32 tg_svo_traverse(const tg_svo* p_svo, v3 ray_origin, v3 ray_direction, f32* p_distance, u32* p_node_idx, u32* p_voxel_idx)
{
    (void)p_svo; (void)ray_origin; (void)ray_direction; (void)p_distance; (void)p_node_idx; (void)p_voxel_idx;
    tgb_set_error("tg_svo_traverse: not built yet");
    return TG_FALSE;
}

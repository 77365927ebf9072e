/*
 * TEST INFRASTRUCTURE. The resumable pieces of K1's per-ray walk (tg_b200/csrc/tgb_k1_walk.cuh: set-up, candidate iterator,
 * cluster march) compiled by a plain C++ compiler and driven pixel by pixel on the host: every ray visits every object (no cull,
 * no front-to-back order: the minimum does not depend on either), suspending and resuming the walk exactly where the kernel's
 * scheduler may. tests/test_k1_walk_cpu.py compares the words with the oracle's visibility buffer.
 */
#include <stdint.h>
#include <string.h>
#include "../../tg_b200/csrc/tgb_k1_walk.cuh"

extern "C" {

void tgbsim_visibility(const tg_object_data* p_objects, u32 n_objects, const u32* p_cluster_pointers, const u32* p_masks, const tg_camera_rays* p_cam,
                       u32 w, u32 h, u32 global_pointer_base, u32 y0, u32 y1, u32 ystep, u32 defer, u64* p_out,
                       u64* p_work /* [3]: candidates marched, set-ups that met the box, second voxels found while one was pending */)
{
    for (size_t i = 0; i < (size_t)w * h; i++) p_out[i] = TG_VIS_CLEAR;
    const v3 camera = tgb_v3(p_cam->camera.x, p_cam->camera.y, p_cam->camera.z);
    for (u32 oi = 0; oi < n_objects; oi++)
    {
        const tg_object_data* o = &p_objects[oi];
        if (o->n_cluster_pointers_per_dim.x == 0 || o->n_cluster_pointers_per_dim.y == 0 || o->n_cluster_pointers_per_dim.z == 0) continue;
        tgb_object_frame f;
        memset(&f, 0, sizeof(f));
        tgb_hoist_object(o, camera, &f);
        f.object_idx = oi;
        tgb_frame_conservative(&f, o, camera);
        for (u32 py = y0; py < y1 && py < h; py += (ystep ? ystep : 1))
        {
            for (u32 px = 0; px < w; px++)
            {
                u64 best = p_out[(size_t)py * w + px];
                /* t_skip equivalent of the best word so far (from other objects) */
                f32 t_skip = TG_F32_MAX;
                if (best != TG_VIS_CLEAR) t_skip = ((f32)(u32)(best >> TG_VIS_DEPTH_SHIFT) + 1.0f) * (p_cam->far_plane * (1.00001f / TG_VIS_DEPTH_SCALE));
                const v3 dir_ws = tgb_pixel_direction(p_cam, w, h, px, py);
                tgb_k1_walk walk;
                if (!tgb_k1_setup(f, dir_ws, &walk)) continue;
                if (p_work) p_work[1]++;
                u32 cx, cy, cz; f32 enter;
                /* defer: k_visibility's deferred word -- the walk goes on with an upper bound of t_skip, the exact word of the voxel found is
                 * computed when the object is finished (or when a second voxel turns up) */
                bool pending = false; u32 pcx = 0, pcy = 0, pcz = 0; i32 pvoxel = -1;
                while (tgb_k1_next_candidate(f, &walk, t_skip, &cx, &cy, &cz, &enter))
                {
                    /* the kernel re-derives the cluster from the stored walk */
                    u32 rx, ry, rz;
                    tgb_k1_current_cluster(&walk, &rx, &ry, &rz);
                    if (rx != cx || ry != cy || rz != cz) { p_out[0] = 0xBADBADBADull; return; }
                    tgb_ray_in_object r;
                    tgb_ray_in_object_restore(&r, walk.d, walk.t_delta_x, walk.t_delta_y, walk.t_delta_z, walk.exotic != 0);
                    if (!defer) tgb_cluster_march(f, r, cx, cy, cz, enter, p_cam->far_plane, p_cluster_pointers, p_masks, global_pointer_base, best, t_skip);
                    else
                    {
                        const i32 voxel = tgb_cluster_find(f, r, cx, cy, cz, enter, p_cluster_pointers, p_masks);
                        if (voxel >= 0)
                        {
                            if (pending) { tgb_cluster_word(f, r, pcx, pcy, pcz, pvoxel, p_cam->far_plane, global_pointer_base, best, t_skip); if (p_work) p_work[2]++; }
                            pending = true; pcx = cx; pcy = cy; pcz = cz; pvoxel = voxel;
                            t_skip = fminf(t_skip, tgb_cluster_t_skip_bound(f, r, cx, cy, cz, voxel, p_cam->far_plane));
                        }
                    }
                    if (p_work) p_work[0]++;
                }
                if (pending)
                {
                    tgb_ray_in_object r;
                    tgb_ray_in_object_restore(&r, walk.d, walk.t_delta_x, walk.t_delta_y, walk.t_delta_z, walk.exotic != 0);
                    tgb_cluster_word(f, r, pcx, pcy, pcz, pvoxel, p_cam->far_plane, global_pointer_base, best, t_skip);
                }
                p_out[(size_t)py * w + px] = best;
            }
        }
    }
}

}

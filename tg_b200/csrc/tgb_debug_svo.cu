/*
 * tgb_debug_svo.cu -- the primary-ray pass of the BLOCKS debug view.
 *
 *   fragment     assets/shaders/raytracer/debug_visibility_svo.frag:27-71
 *   dispatch     tgvk_raytracer.c:1226-1272 (one full-screen quad per SVO INSTEAD of the cluster pass while
 *                TG_DEBUG_SHOW_BLOCKS is selected)
 * One thread per pixel shoots the UN-normalised pixel direction from the camera through the SVO with the shader's own stack
 * machine (tgb_svo_traverse.cuh) and resolves  depth24 | leaf node index | voxel % 512  with the same 64-bit atomicMin as K1,
 * so the shading pass (k_shade, view 5: the pointer field goes through the cluster-pointer table and is hashed,
 * shading.frag:122-126,247-256) and a sharded frame's min-merge work unchanged. A debug view: primary rays are coherent,
 * nothing here is tuned.
 */
#include "tgb_device.cuh"
#include "tgb_svo_traverse.cuh"

__global__ void __launch_bounds__(256) k_visibility_svo(const u32* __restrict__ p_nodes, const u32* __restrict__ p_leaf_data, const u32* __restrict__ p_voxels, v3 bmin, v3 bmax,
                                                        tg_camera_rays cam, u32 w, u32 h, u64* __restrict__ p_vis, u32 n_ranks, u32 tile_rows)
{
    /* 8x4 pixel blocks per warp like K1 */
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u32 px = blockIdx.x * 16u + (warp & 1u) * 8u + (lane & 7u);
    const u32 py = blockIdx.y * 16u + (warp >> 1) * 4u + (lane >> 3);
    if (px >= w || py >= h) return;
    const v3 dir = tgb_pixel_direction(&cam, w, h, px, py);
    u32 node_idx, voxel_idx;
    const f32 d = tgb_svo_traverse_stack(p_nodes, p_leaf_data, p_voxels, bmin, bmax, cam.far_plane, tgb_v3(cam.camera.x, cam.camera.y, cam.camera.z), dir, &node_idx, &voxel_idx);
    const u64 word = tgb_svo_visibility_word(d, node_idx, voxel_idx);
    if (word != TG_VIS_CLEAR) atomicMin((unsigned long long*)&p_vis[(u64)tgb_row_to_virtual(py, n_ranks, tile_rows) * w + px], (unsigned long long)word);
}

extern "C" b32 tgbd_render_visibility_svo(struct tgb_device* d, const tg_camera_rays* p_cam)
{
    TGB_CUDA(cudaSetDevice(d->device));
    if (!d->svo.valid) { tgb_set_error("render_visibility_svo: the BLOCKS view needs an SVO (tgb200_svo_update)"); return TG_FALSE; }
    TGB_CUDA(cudaEventRecord(d->ev[2], d->stream));
    TGB_CUDA(cudaEventRecord(d->ev[3], d->stream)); /* no cull stage */
    const dim3 grid((d->width + 15) / 16, (d->height + 15) / 16);
    k_visibility_svo<<<grid, 256, 0, d->stream>>>(d->svo.d_nodes, d->svo.d_leaf_data, d->svo.d_voxels, d->svo.bmin, d->svo.bmax, *p_cam, d->width, d->height, d->d_vis,
                                                  d->n_ranks, d->tile_rows);
    TGB_LAUNCH_CHECK(d);
    TGB_CUDA(cudaEventRecord(d->ev[4], d->stream));
    d->ev_vis = TG_TRUE;
    d->n_visible_objects = 0;
    return TG_TRUE;
}

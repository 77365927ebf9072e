/*
 * tgb_shade.cu -- K3, the shading pass with the composed 1-bounce GI term.
 *
 * Replaces the reference's full-screen fragment pass
 *   shading            assets/shaders/raytracer/shading.frag:114-337 (BRDF helpers :53-110)
 *   SVO traversal      assets/shaders/raytracer/svo_functions.inc:1-329 (bound by shading.frag:40-43 but never
 *                      called there; the only GPU caller is debug_visibility_svo.frag:52)
 *   secondary rays     tgvk_raytracer.c:1405-1431 (CPU sketch under `if (0)`), TODO.h:33-43, util.inc:1-56 (RNG, hash)
 * One thread per pixel: unpack the visibility word, fetch material + object, re-derive the primary ray in the winning
 * cluster's space with the per-object factorisation (tgb_hoist.h, once per object per frame in k_object_frames
 * instead of once per fragment), face normal, Cook-Torrance term, then ONE secondary ray per hit pixel through the
 * replicated 1-bit SVO; a miss keeps the ambient term, a hit removes it (GI spec pinned in DESIGN.md, oracle twin:
 * oracle/tgo_shade.c). Floating point: GI radiance tolerance is 1e-3 relative (BASELINE.json), the arithmetic below
 * nevertheless follows the shader's operation order (this TU is built with -fmad=false like the others).
 */
#include "tgb_device.cuh"

#define TGB_PI_F               3.14159265358979323846f
#define TGB_TRAVERSE_MAX_ITERS 4096u /* Q9: cap that valid input never reaches */

/* ---- per-object frames for ALL objects, indexed by object idx -------------------------------- */
__global__ void k_object_frames(const tg_object_data* __restrict__ p_objects, u32 object_capacity, v3 camera, tgb_object_frame* __restrict__ p_frames)
{
    const u32 object_idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (object_idx >= object_capacity) return;
    const tg_object_data o = p_objects[object_idx];
    if (o.n_cluster_pointers_per_dim.x == 0 || o.n_cluster_pointers_per_dim.y == 0 || o.n_cluster_pointers_per_dim.z == 0) return;
    tgb_object_frame f;
    tgb_hoist_object(&o, camera, &f);
    f.object_idx = object_idx;
    f.og[0] = f.og[1] = f.og[2] = 0.0f; f.eps = 0.0f; f.x0 = f.y0 = f.x1 = f.y1 = 0; f.min_depth24 = 0;
    p_frames[object_idx] = f;
}

/* ---- BRDF: shading.frag:53-110 ------------------------------------------------------------------ */
__device__ __forceinline__ f32 tgb_distribution_ggx(f32 clamped_n_dot_h, f32 roughness)
{
    const f32 a = roughness * roughness;
    const f32 a_sqr = a * a;
    f32 denom = (clamped_n_dot_h * clamped_n_dot_h * (a_sqr - 1.0f) + 1.0f);
    denom = TGB_PI_F * denom * denom;
    return a_sqr / denom;
}
__device__ __forceinline__ f32 tgb_geometry_schlick_ggx(f32 n_dot_v, f32 roughness)
{
    const f32 r = (roughness + 1.0f);
    const f32 k = (r * r) / 8.0f;
    const f32 denom = n_dot_v * (1.0f - k) + k;
    return n_dot_v / denom;
}
__device__ __forceinline__ f32 tgb_pow5(f32 x) { return ((x * x) * (x * x)) * x; } /* pow(x, 5.0) pinned, oracle/tgo_shade.c */
__device__ __forceinline__ f32 tgb_fresnel_schlick(f32 u, f32 roughness)
{
    const f32 f_lambda = 1.0f - roughness;
    return f_lambda + (1.0f - f_lambda) * tgb_pow5(1.0f - u);
}
__device__ __forceinline__ v3 tgb_shade_brdf(v3 n, v3 v, v3 l, v3 diffuse_albedo, v3 specular_albedo, f32 metallic, f32 roughness, v3 radiance)
{
    const v3 h = tgb_normalize(tgb_add(v, l));
    const f32 clamped_h_dot_n = tgb_clamp(tgb_dot(h, n), 0.0f, 1.0f);
    const f32 clamped_l_dot_n = tgb_clamp(tgb_dot(n, l), 0.0f, 1.0f);
    const f32 clamped_n_dot_v = tgb_clamp(tgb_dot(n, v), 0.0f, 1.0f);
    const f32 d = tgb_distribution_ggx(clamped_h_dot_n, roughness);
    const f32 f = tgb_fresnel_schlick(clamped_n_dot_v, roughness);
    const f32 g = tgb_geometry_schlick_ggx(clamped_l_dot_n, roughness) * tgb_geometry_schlick_ggx(clamped_n_dot_v, roughness);
    const f32 dfg = d * f * g;
    const f32 denominator = 4.0f * clamped_n_dot_v * clamped_l_dot_n;
    const v3 specular = tgb_scale(specular_albedo, dfg / tgb_max(denominator, 0.001f));
    const f32 kd = (1.0f - f) * (1.0f - metallic);
    const v3 diffuse = tgb_divf(tgb_scale(diffuse_albedo, kd), TGB_PI_F);
    return tgb_scale(tgb_mul(tgb_add(diffuse, specular), radiance), clamped_l_dot_n);
}

__device__ __forceinline__ float4 tgb_hash_color(u32 v)
{
    const u32 h0 = tgb_hash_u32(v), h1 = tgb_hash_u32(h0), h2 = tgb_hash_u32(h1);
    return make_float4((f32)h0 / 4294967295.0f, (f32)h1 / 4294967295.0f, (f32)h2 / 4294967295.0f, 1.0f);
}

/* ---- SVO traversal: svo_functions.inc:1-329 ---------------------------------------------------- */
struct tgb_svo_view
{
    const u32* __restrict__ p_nodes;
    const u32* __restrict__ p_leaf_data; /* 65 u32 per leaf: n, cluster_idcs[64] */
    const u32* __restrict__ p_voxels;    /* 1024 u32 per leaf */
    v3 bmin, bmax;
};

/* svo_functions.inc:283-292 */
__device__ __forceinline__ f32 tgb_exit_distance(v3 bmin, v3 bmax, v3 position, v3 d)
{
    const f32 ax = (d.x == 0.0f) ? TG_F32_MIN : ((bmin.x - position.x) / d.x);
    const f32 ay = (d.y == 0.0f) ? TG_F32_MIN : ((bmin.y - position.y) / d.y);
    const f32 az = (d.z == 0.0f) ? TG_F32_MIN : ((bmin.z - position.z) / d.z);
    const f32 bx = (d.x == 0.0f) ? TG_F32_MAX : ((bmax.x - position.x) / d.x);
    const f32 by = (d.y == 0.0f) ? TG_F32_MAX : ((bmax.y - position.y) / d.y);
    const f32 bz = (d.z == 0.0f) ? TG_F32_MAX : ((bmax.z - position.z) / d.z);
    return tgb_min(tgb_min(tgb_max(ax, bx), tgb_max(ay, by)), tgb_max(az, bz));
}

/*
 * Returns the hit depth in [0,1) or 1.0 on a miss, like tg_svo_traverse. The reference keeps a stack of five
 * (node, min, max) triples (svo.inc:4); here the stack holds only the node index and the octant taken at each
 * level -- a node's box is re-derived from the root box by the same chain of additions the shader performs
 * (child_min = parent_min [+ child_extent], child_max = child_min + child_extent [+ child_extent]), so the values
 * are the shader's, and the per-thread state stays in registers instead of 35 words of local memory.
 */
__device__ f32 tgb_svo_traverse(const tgb_svo_view& svo, f32 far_plane, v3 ray_origin_ws, v3 d)
{
    const v3 extent = tgb_sub(svo.bmax, svo.bmin);
    const v3 center = tgb_add(tgb_scale(extent, 0.5f), svo.bmin);
    const v3 o = tgb_sub(ray_origin_ws, center);

    f32 enter, exit;
    if (!tgb_ray_aabb(o, d, svo.bmin, svo.bmax, &enter, &exit)) return 1.0f;
    v3 position = o;
    if (enter > 0.0f) position = tgb_add(position, tgb_scale(d, enter));

    /* stack entry s (0 = root): node index idx[s]; path bits [3(s-1), 3s) = octant of entry s inside entry s-1 */
    u32 idx0 = 0, idx1 = 0, idx2 = 0, idx3 = 0, idx4 = 0;
    u32 path = 0;
    u32 stack_size = 1;
    f32 result = 1.0f;

    for (u32 iterations = 0; stack_size > 0 && iterations < TGB_TRAVERSE_MAX_ITERS; iterations++)
    {
        /* box of the top entry, by the shader's additions from the root box */
        v3 parent_min = svo.bmin, parent_max = svo.bmax;
        for (u32 s = 1; s < stack_size; s++)
        {
            const v3 ce = tgb_scale(tgb_sub(parent_max, parent_min), 0.5f);
            const u32 oct = (path >> (3u * (s - 1u))) & 7u;
            v3 cmin = parent_min;
            v3 cmax = tgb_add(cmin, ce);
            if (oct & 1u) { cmin.x += ce.x; cmax.x += ce.x; }
            if (oct & 2u) { cmin.y += ce.y; cmax.y += ce.y; }
            if (oct & 4u) { cmin.z += ce.z; cmax.z += ce.z; }
            parent_min = cmin; parent_max = cmax;
        }
        const u32 top = stack_size - 1u;
        const u32 parent_idx = top == 0 ? idx0 : (top == 1 ? idx1 : (top == 2 ? idx2 : (top == 3 ? idx3 : idx4)));

        const u32 node_data = __ldg(&svo.p_nodes[parent_idx]);
        const u32 child_pointer =  node_data        & 0xFFFFu;
        const u32 valid_mask    = (node_data >> 16) & 0xFFu;
        const u32 leaf_mask     = (node_data >> 24) & 0xFFu;

        /* svo_functions.inc:57-80 */
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
            const u32 child_idx = parent_idx + child_pointer + (u32)__popc(valid_mask & ((1u << relative_child_idx) - 1u));
            if ((leaf_mask & (1u << relative_child_idx)) != 0)
            {
                const u32 data_pointer = __ldg(&svo.p_nodes[child_idx]);
                if (__ldg(&svo.p_leaf_data[(u64)data_pointer * 65u]) != 0)
                {
                    /* :111-257: DDA through the 32^3 block */
                    const u32* __restrict__ p_block = svo.p_voxels + (u64)data_pointer * TG_SVO_BLOCK_WORDS;
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
                    if (d.x > 0.0f)      { step_x = 1;  t_max_x = ((f32)(x + 1) - hit.x) / d.x;  t_delta_x = 1.0f / d.x; }
                    else if (d.x < 0.0f) { step_x = -1; t_max_x = (hit.x - (f32)x) / -d.x;       t_delta_x = 1.0f / -d.x; }
                    if (d.y > 0.0f)      { step_y = 1;  t_max_y = ((f32)(y + 1) - hit.y) / d.y;  t_delta_y = 1.0f / d.y; }
                    else if (d.y < 0.0f) { step_y = -1; t_max_y = (hit.y - (f32)y) / -d.y;       t_delta_y = 1.0f / -d.y; }
                    if (d.z > 0.0f)      { step_z = 1;  t_max_z = ((f32)(z + 1) - hit.z) / d.z;  t_delta_z = 1.0f / d.z; }
                    else if (d.z < 0.0f) { step_z = -1; t_max_z = (hit.z - (f32)z) / -d.z;       t_delta_z = 1.0f / -d.z; }

                    const i32 ex = (i32)child_extent.x, ey = (i32)child_extent.y, ez = (i32)child_extent.z;
                    for (;;)
                    {
                        /* one x-row of the block is one word when the block is 32 wide (the only size the builder makes) */
                        const u32 relative_voxel_idx = (u32)(ex * ey * z + ex * y + x);
                        const u32 bits = __ldg(&p_block[relative_voxel_idx >> 5]);
                        if ((bits >> (relative_voxel_idx & 31u)) & 1u)
                        {
                            const v3 voxel_min = tgb_add(child_min, tgb_v3((f32)x, (f32)y, (f32)z));
                            const v3 voxel_max = tgb_add(child_min, tgb_v3((f32)(x + 1), (f32)(y + 1), (f32)(z + 1)));
                            tgb_ray_aabb(o, d, voxel_min, voxel_max, &enter, &exit);
                            result = enter / far_plane;
                            break;
                        }
                        if (t_max_x < t_max_y)
                        {
                            if (t_max_x < t_max_z) { t_max_x += t_delta_x; x += step_x; if (x < 0 || x >= ex) break; }
                            else                   { t_max_z += t_delta_z; z += step_z; if (z < 0 || z >= ez) break; }
                        }
                        else
                        {
                            if (t_max_y < t_max_z) { t_max_y += t_delta_y; y += step_y; if (y < 0 || y >= ey) break; }
                            else                   { t_max_z += t_delta_z; z += step_z; if (z < 0 || z >= ez) break; }
                        }
                    }
                    if (result < 1.0f) break;
                }
            }
            else
            {
                /* :262-270: push */
                advance_to_border = false;
                if (stack_size == 1) idx1 = child_idx; else if (stack_size == 2) idx2 = child_idx; else if (stack_size == 3) idx3 = child_idx; else idx4 = child_idx;
                path = (path & ~(7u << (3u * (stack_size - 1u)))) | (relative_child_idx << (3u * (stack_size - 1u)));
                stack_size++;
                if (stack_size > TG_SVO_TRAVERSE_STACK_CAPACITY) return 1.0f; /* malformed tree (deeper than 5 inner levels) */
            }
        }

        if (advance_to_border)
        {
            /* :279-324 */
            exit = tgb_exit_distance(child_min, child_max, position, d);
            position = tgb_add(position, tgb_scale(d, exit + TG_F32_EPSILON));
            while (stack_size > 0)
            {
                v3 smin = svo.bmin, smax = svo.bmax;
                for (u32 s = 1; s < stack_size; s++)
                {
                    const v3 ce = tgb_scale(tgb_sub(smax, smin), 0.5f);
                    const u32 oct = (path >> (3u * (s - 1u))) & 7u;
                    v3 cmin = smin;
                    v3 cmax = tgb_add(cmin, ce);
                    if (oct & 1u) { cmin.x += ce.x; cmax.x += ce.x; }
                    if (oct & 2u) { cmin.y += ce.y; cmax.y += ce.y; }
                    if (oct & 4u) { cmin.z += ce.z; cmax.z += ce.z; }
                    smin = cmin; smax = cmax;
                }
                exit = tgb_exit_distance(smin, smax, position, d);
                if (exit > TG_F32_EPSILON) break;
                stack_size--;
            }
        }
    }
    return result < 1.0f ? result : 1.0f;
}

/* ---- K3 ----------------------------------------------------------------------------------------- */
struct tgb_shade_args
{
    const u64* __restrict__ p_vis;
    float4* __restrict__ p_out;
    const u32* __restrict__ p_cluster_pointers;
    const u32* __restrict__ p_c2o;
    const tg_object_data* __restrict__ p_objects;
    const tgb_object_frame* __restrict__ p_frames; /* by object idx */
    const u8* __restrict__ p_lut_idx;
    const u32* __restrict__ p_color_lut;
    tgb_svo_view svo;
    tg_camera_rays cam;
    u32 w, h;
    u32 global_pointer_base, n_local_pointers;
    u32 gi_enabled, frame_seed, debug_visualization;
    u32 y0, y1; /* rows [y0, y1) are shaded (multi-GPU: this rank's screen tile) */
};

__global__ void __launch_bounds__(256) k_shade(const tgb_shade_args a)
{
    /* 8x4 pixel blocks per warp like K1: neighbouring pixels share clusters, objects and SVO leaves */
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u32 px = blockIdx.x * 16u + (warp & 1u) * 8u + (lane & 7u);
    const u32 py = a.y0 + blockIdx.y * 16u + (warp >> 1) * 4u + (lane >> 3);
    if (px >= a.w || py >= a.y1) return;
    const u64 pixel = (u64)py * a.w + px;

    /* shading.frag:116-120 */
    const u64 packed_data = a.p_vis[pixel];
    const f32 depth_24b           = (f32)(u32)(packed_data >> TG_VIS_DEPTH_SHIFT) / TG_VIS_DEPTH_SCALE;
    const u32 cluster_pointer_31b = (u32)(packed_data >> TG_VIS_POINTER_SHIFT) & 2147483647u;
    const u32 voxel_idx_9b        = (u32)(packed_data) & 511u;

    if (!(depth_24b < 1.0f)) { a.p_out[pixel] = make_float4(1.0f, 0.0f, 1.0f, 1.0f); return; } /* :335 */

    const u32 local_pointer = cluster_pointer_31b - a.global_pointer_base;
    if (local_pointer >= a.n_local_pointers) { a.p_out[pixel] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); return; } /* another shard's cluster */
    const u32 cluster_idx = __ldg(&a.p_cluster_pointers[local_pointer]);
    const u32 object_idx = __ldg(&a.p_c2o[cluster_idx]);
    const tgb_object_frame& f = a.p_frames[object_idx];
    const tg_object_data& obj = a.p_objects[object_idx];

    /* :128-135; per-object LUT (Q2) */
    const u32 color_lut_idx = __ldg(&a.p_lut_idx[(u64)cluster_idx * 512u + voxel_idx_9b]);
    const u32 packed_color = __ldg(&a.p_color_lut[obj.lut_idx * 256u + color_lut_idx]);
    const f32 color_r = (f32)( packed_color >> 24        ) / 255.0f;
    const f32 color_g = (f32)((packed_color >> 16) & 0xffu) / 255.0f;
    const f32 color_b = (f32)((packed_color >>  8) & 0xffu) / 255.0f;

    /* :144-181: the primary ray in the winning cluster's space */
    const u32 rel = local_pointer - f.first_cluster_pointer;
    const u32 cx = rel % f.nx, cy = (rel / f.nx) % f.ny, cz = rel / (f.nx * f.ny);
    const v3 ray_origin_ws = tgb_v3(a.cam.camera.x, a.cam.camera.y, a.cam.camera.z);
    const v3 ray_origin_ms = tgb_hoist_cluster_origin(&f, cx, cy, cz);
    const v3 ray_direction_ws = tgb_pixel_direction(&a.cam, a.w, a.h, px, py);
    const v3 ray_direction_ms = tgb_hoist_direction(&f, ray_direction_ws);

    /* :183-228 */
    const v3 voxel_min = tgb_v3((f32)(voxel_idx_9b % 8u), (f32)((voxel_idx_9b / 8u) % 8u), (f32)(voxel_idx_9b / 64u));
    const v3 voxel_max = tgb_add(voxel_min, tgb_v3(1.0f, 1.0f, 1.0f));
    v3 normal_ws = tgb_v3(0.0f, 0.0f, 0.0f);
    f32 enter, exit;
    if (tgb_ray_aabb(ray_origin_ms, ray_direction_ms, voxel_min, voxel_max, &enter, &exit))
    {
        const v3 hit_position_ms = enter > 0.0f ? tgb_add(ray_origin_ms, tgb_scale(ray_direction_ms, enter)) : ray_origin_ms;
        const v3 voxel_center_ms = tgb_add(voxel_min, tgb_v3(0.5f, 0.5f, 0.5f));
        v3 n = tgb_sub(hit_position_ms, voxel_center_ms);
        if (fabsf(n.x) > fabsf(n.y))
        {
            n.y = 0.0f;
            if (fabsf(n.x) > fabsf(n.z)) { n.x = tgb_sign(n.x); n.z = 0.0f; }
            else                         { n.z = tgb_sign(n.z); n.x = 0.0f; }
        }
        else
        {
            n.x = 0.0f;
            if (fabsf(n.y) > fabsf(n.z)) { n.y = tgb_sign(n.y); n.z = 0.0f; }
            else                         { n.z = tgb_sign(n.z); n.y = 0.0f; }
        }
        normal_ws = tgb_normalize(tgb_m4_transform(obj.rotation, n, 0.0f));
    }

    /* :231 -- un-normalised direction (Q3) */
    const v3 hit_position_ws = tgb_add(ray_origin_ws, tgb_scale(ray_direction_ws, depth_24b * a.cam.far_plane));

    /* :233-300 debug views */
    switch (a.debug_visualization)
    {
    case TG_DEBUG_SHOW_OBJECT_INDEX:    a.p_out[pixel] = tgb_hash_color(object_idx); return;
    case TG_DEBUG_SHOW_DEPTH:           { const f32 g = tgb_min(1.0f, 8.0f * depth_24b); a.p_out[pixel] = make_float4(g, g, g, 1.0f); return; }
    case TG_DEBUG_SHOW_CLUSTER_INDEX:
    case TG_DEBUG_SHOW_BLOCKS:          a.p_out[pixel] = tgb_hash_color(cluster_idx); return;
    case TG_DEBUG_SHOW_VOXEL_INDEX:     a.p_out[pixel] = tgb_hash_color(voxel_idx_9b); return;
    case TG_DEBUG_SHOW_COLOR_LUT_INDEX: a.p_out[pixel] = tgb_hash_color(color_lut_idx); return;
    case TG_DEBUG_SHOW_COLOR:           a.p_out[pixel] = make_float4(color_r, color_g, color_b, 1.0f); return;
    case TG_DEBUG_SHOW_NORMAL:          a.p_out[pixel] = make_float4(normal_ws.x * 0.5f + 0.5f, normal_ws.y * 0.5f + 0.5f, normal_ws.z * 0.5f + 0.5f, 1.0f); return;
    default: break;
    }

    /* :285-316 */
    const f32 metallic = 0.1f;
    const v3 v = tgb_normalize(tgb_sub(ray_origin_ws, hit_position_ws));
    const v3 l = tgb_normalize(tgb_v3(0.0f, 0.8f, 0.3f));
    const v3 albedo = a.debug_visualization == TG_DEBUG_SHOW_SHADING ? tgb_v3(1.0f, 1.0f, 1.0f) : tgb_v3(color_r, color_g, color_b);
    const v3 specular_albedo = tgb_mix3(tgb_v3(0.04f, 0.04f, 0.04f), albedo, metallic);
    const f32 roughness = 0.8f;
    const v3 lo = tgb_shade_brdf(normal_ws, v, l, albedo, specular_albedo, metallic, roughness, tgb_v3(3.0f, 3.0f, 3.0f));
    v3 ambient = tgb_scale(albedo, 0.1f);

    /* composed GI term (DESIGN.md "GI spec"; oracle/tgo_shade.c) */
    if (a.gi_enabled && a.debug_visualization == TG_DEBUG_SHOW_NONE)
    {
        f32 visibility = 1.0f;
        if (normal_ws.x != 0.0f || normal_ws.y != 0.0f || normal_ws.z != 0.0f)
        {
            const u32 pixel_idx = a.w * py + px;
            u32 rng = tgb_hash_u32(pixel_idx ^ tgb_hash_u32(a.frame_seed)) | 1u;
            v3 dir = normal_ws;
            for (u32 attempt = 0; attempt < 32; attempt++)
            {
                v3 c;
                c.x = tgb_xorshift32_range(&rng, -1.0f, 1.0f);
                c.y = tgb_xorshift32_range(&rng, -1.0f, 1.0f);
                c.z = tgb_xorshift32_range(&rng, -1.0f, 1.0f);
                c = tgb_normalize(c);
                if (tgb_dot(c, normal_ws) > 0.0f) { dir = c; break; }
            }
            const v3 origin = tgb_add(hit_position_ws, tgb_scale(dir, 1.73205080757f));
            const f32 depth2 = tgb_svo_traverse(a.svo, a.cam.far_plane, origin, dir);
            visibility = depth2 < 1.0f ? 0.0f : 1.0f;
        }
        ambient = tgb_scale(ambient, visibility);
    }

    a.p_out[pixel] = make_float4(ambient.x + lo.x, ambient.y + lo.y, ambient.z + lo.z, 1.0f);
}

extern "C" b32 tgbd_render_shading(struct tgb_device* d, const tg_camera_rays* p_cam, u32 n_local_pointers, u32 gi_enabled, u32 frame_seed, u32 debug_visualization,
                                   u32 y0, u32 y1)
{
    TGB_CUDA(cudaSetDevice(d->device));
    if (gi_enabled && debug_visualization == TG_DEBUG_SHOW_NONE && !d->svo.valid)
    {
        tgb_set_error("render_shading: GI is enabled but no SVO has been built or uploaded");
        return TG_FALSE;
    }
    if (y1 > d->height) y1 = d->height;
    if (y0 >= y1) return TG_TRUE;
    TGB_CUDA(cudaEventRecord(d->ev[7], d->stream));
    k_object_frames<<<(d->object_capacity + 127) / 128, 128, 0, d->stream>>>(d->d_objects, d->object_capacity,
                                                                            tgb_v3(p_cam->camera.x, p_cam->camera.y, p_cam->camera.z), d->d_frames_all);
    TGB_LAUNCH_CHECK(d);

    tgb_shade_args a;
    a.p_vis = d->d_vis;
    a.p_out = d->d_radiance;
    a.p_cluster_pointers = d->d_cluster_pointers;
    a.p_c2o = d->d_c2o;
    a.p_objects = d->d_objects;
    a.p_frames = d->d_frames_all;
    a.p_lut_idx = d->d_lut_idx;
    a.p_color_lut = d->d_color_lut;
    a.svo.p_nodes = d->svo.d_nodes;
    a.svo.p_leaf_data = d->svo.d_leaf_data;
    a.svo.p_voxels = d->svo.d_voxels;
    a.svo.bmin = d->svo.bmin;
    a.svo.bmax = d->svo.bmax;
    a.cam = *p_cam;
    a.w = d->width; a.h = d->height;
    a.global_pointer_base = d->global_pointer_base;
    a.n_local_pointers = n_local_pointers;
    a.gi_enabled = gi_enabled; a.frame_seed = frame_seed; a.debug_visualization = debug_visualization;
    a.y0 = y0; a.y1 = y1;
    const dim3 grid((d->width + 15) / 16, (y1 - y0 + 15) / 16);
    k_shade<<<grid, 256, 0, d->stream>>>(a);
    TGB_LAUNCH_CHECK(d);
    TGB_CUDA(cudaEventRecord(d->ev[8], d->stream));
    d->ev_shade = TG_TRUE;
    return TG_TRUE;
}

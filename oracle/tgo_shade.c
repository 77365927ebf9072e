/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY. Nothing under tg_b200/ may include, link or call this.
 *
 * tgo_shade.c: assets/shaders/raytracer/shading.frag:53-337 as scalar C, plus the composed
 * 1-bounce GI term. The reference has NO GPU GI pass: shading.frag binds
 * the SVO but never traverses it. The GI term below is composed from reference pieces and is
 * PINNED HERE FIRST (DESIGN.md "GI spec"); the CUDA kernel follows this file:
 *   seed      = hash_u32(pixel_idx ^ hash_u32(frame_seed)) | 1          (util.inc:47-56)
 *   direction = normalize(rand in [-1,1]^3), xorshift32 (util.inc:1-31), rejected until
 *               dot(direction, normal_ws) > 0, at most 32 attempts      (tgvk_raytracer.c:1405-1417, with
 *               the author's TODO applied: the surface normal and `> 0`)
 *               all 32 rejected -> direction = normal_ws; normal_ws == 0 -> unoccluded
 *   origin    = hit_position_ws + direction * 1.73205080757             (tgvk_raytracer.c:1419)
 *   V         = tg_svo_traverse(...) misses ? 1 : 0                     (svo_functions.inc:1-329; TODO.h:38-43
 *               "from hit: raycast random towards sky")
 *   out.rgb   = (0.1 * albedo) * V + lo                                 (shading.frag:313-315 with ambient * V)
 * Pinned where GLSL is implementation-defined: pow(x, 5.0) = ((x*x)*(x*x))*x; TG_PI = 3.14159274f.
 */
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "tgo.h"
#include "tgo_math.h"

#define TGO_PI 3.14159265358979323846f

/* shading.frag:53-61 */
static f32 tgo__distribution_ggx(f32 clamped_n_dot_h, f32 roughness)
{
    const f32 a = roughness * roughness;
    const f32 a_sqr = a * a;
    f32 denom = (clamped_n_dot_h * clamped_n_dot_h * (a_sqr - 1.0f) + 1.0f);
    denom = TGO_PI * denom * denom;
    return a_sqr / denom;
}

/* shading.frag:63-70 */
static f32 tgo__geometry_schlick_ggf(f32 n_dot_v, f32 roughness)
{
    const f32 r = (roughness + 1.0f);
    const f32 k = (r * r) / 8.0f;
    const f32 denom = n_dot_v * (1.0f - k) + k;
    return n_dot_v / denom;
}

/* shading.frag:72-78 */
static f32 tgo__geometry_smith(f32 clamped_n_dot_v, f32 clamped_n_dot_l, f32 roughness)
{
    const f32 ggx2 = tgo__geometry_schlick_ggf(clamped_n_dot_v, roughness);
    const f32 ggx1 = tgo__geometry_schlick_ggf(clamped_n_dot_l, roughness);
    return ggx1 * ggx2;
}

static f32 tgo__pow5(f32 x) { return ((x * x) * (x * x)) * x; }

/* shading.frag:80-84 */
static f32 tgo__fresnel_schlick(f32 u, f32 roughness)
{
    const f32 f_lambda = 1.0f - roughness;
    return f_lambda + (1.0f - f_lambda) * tgo__pow5(1.0f - u);
}

/* shading.frag:86-110 */
static v3 tgo__shade(v3 n, v3 v, v3 l, v3 diffuse_albedo, v3 specular_albedo, f32 metallic, f32 roughness, v3 radiance)
{
    const v3 h = tgo_v3_normalized(tgo_v3_add(v, l));

    const f32 h_dot_n = tgo_v3_dot(h, n);
    const f32 l_dot_n = tgo_v3_dot(n, l);
    const f32 n_dot_v = tgo_v3_dot(n, v);

    const f32 clamped_h_dot_n = tgo_clamp(h_dot_n, 0.0f, 1.0f);
    const f32 clamped_l_dot_n = tgo_clamp(l_dot_n, 0.0f, 1.0f);
    const f32 clamped_n_dot_v = tgo_clamp(n_dot_v, 0.0f, 1.0f);

    const f32 d = tgo__distribution_ggx(clamped_h_dot_n, roughness);
    const f32 f = tgo__fresnel_schlick(clamped_n_dot_v, roughness);
    const f32 g = tgo__geometry_smith(clamped_n_dot_v, clamped_l_dot_n, roughness);
    const f32 dfg = d * f * g;
    const f32 denominator = 4.0f * clamped_n_dot_v * clamped_l_dot_n;
    const v3 specular = tgo_v3_mulf(specular_albedo, dfg / tgo_max(denominator, 0.001f));

    const f32 kd = (1.0f - f) * (1.0f - metallic);
    const v3 diffuse = tgo_v3_divf(tgo_v3_mulf(diffuse_albedo, kd), TGO_PI);

    return tgo_v3_mulf(tgo_v3_mul(tgo_v3_add(diffuse, specular), radiance), clamped_l_dot_n);
}

static void tgo__hash_color(u32 v, f32* p_rgba)
{
    const u32 h0 = tgo_hash_u32(v);
    const u32 h1 = tgo_hash_u32(h0);
    const u32 h2 = tgo_hash_u32(h1);
    p_rgba[0] = (f32)h0 / 4294967295.0f;
    p_rgba[1] = (f32)h1 / 4294967295.0f;
    p_rgba[2] = (f32)h2 / 4294967295.0f;
    p_rgba[3] = 1.0f;
}

static void tgo__shade_pixel(const tgo_scene_view* p_scene, const tg_camera_rays* p_cam, u32 w, u32 h, u64 packed_data, const tg_svo* p_svo,
                             u32 gi_enabled, u32 frame_seed, u32 debug_visualization, u32 px, u32 py, f32* p_rgba, f32* p_gi_ray_or_null)
{
    /* shading.frag:116-120 */
    const f32 depth_24b           = (f32)(packed_data >> TG_VIS_DEPTH_SHIFT) / TG_VIS_DEPTH_SCALE;
    const u32 cluster_pointer_31b = (u32)(packed_data >> TG_VIS_POINTER_SHIFT) & 2147483647u;
    const u32 voxel_idx_9b        = (u32)(packed_data) & 511u;

    if (!(depth_24b < 1.0f))
    {
        p_rgba[0] = 1.0f; p_rgba[1] = 0.0f; p_rgba[2] = 1.0f; p_rgba[3] = 1.0f; /* :335 */
        return;
    }

    if (debug_visualization == TG_DEBUG_SHOW_BLOCKS)
    {
        /* The word was written by debug_visibility_svo.frag (tgo_visibility_svo): its pointer field is an SVO NODE index, which
         * shading.frag:122-126,247-256 nevertheless runs through the cluster-pointer table and hashes. Everything the shader computes
         * in between (material, normal) is dead for this view. Defined deviation: a node index beyond the live pointer range reads
         * whatever the reference's SSBO holds there; here it reads as cluster 0. */
        const u32 cluster_idx_of_node = cluster_pointer_31b < p_scene->n_cluster_pointers ? p_scene->p_cluster_pointers[cluster_pointer_31b] : 0u;
        tgo__hash_color(cluster_idx_of_node, p_rgba);
        return;
    }

    /* multi-GPU: the word carries a GLOBAL pointer; this view holds pointers [base, base + n) */
    const u32 local_pointer = cluster_pointer_31b - p_scene->global_pointer_base;
    const u32 cluster_idx = p_scene->p_cluster_pointers[local_pointer];
    const u32 object_idx = p_scene->p_cluster_idx_to_object_idx[cluster_idx];
    const tg_object_data* p_object = &p_scene->p_objects[object_idx];

    /* :128-135 (byte (abs_voxel % 4) of word abs_voxel / 4 == byte abs_voxel); per-object LUT (Q2) */
    const u64 absolute_voxel_idx = (u64)cluster_idx * 512u + (u64)voxel_idx_9b;
    const u32 color_lut_idx = p_scene->p_color_lut_idx_data[absolute_voxel_idx];
    const u32 packed_color = p_scene->p_color_lut[(size_t)p_object->lut_idx * 256u + color_lut_idx];
    const f32 color_r = (f32)( packed_color >> 24        ) / 255.0f;
    const f32 color_g = (f32)((packed_color >> 16) & 0xffu) / 255.0f;
    const f32 color_b = (f32)((packed_color >>  8) & 0xffu) / 255.0f;

    /* :144-181 */
    const m4 ws2ms = tgo_ws2ms(p_object, local_pointer);
    const v3 ray_origin_ws = tgo_v3(p_cam->camera.x, p_cam->camera.y, p_cam->camera.z);
    const v3 ray_origin_ms = tgo_m4_mulv3w(ws2ms, ray_origin_ws, 1.0f);
    const v3 ray_direction_ws = tgo_pixel_ray_direction_nn(p_cam, w, h, px, py);
    const v3 ray_direction_ms = tgo_v3_normalized(tgo_m4_mulv3w(ws2ms, ray_direction_ws, 0.0f));

    /* :183-188 */
    const v3 voxel_min = tgo_v3((f32)(voxel_idx_9b % 8u), (f32)((voxel_idx_9b / 8u) % 8u), (f32)(voxel_idx_9b / 64u));
    const v3 voxel_max = tgo_v3_add(voxel_min, tgo_v3(1.0f, 1.0f, 1.0f));

    v3 normal_ws = tgo_v3(0.0f, 0.0f, 0.0f);
    f32 enter, exit;
    if (tgo_intersect_ray_aabb_glsl(ray_origin_ms, ray_direction_ms, voxel_min, voxel_max, &enter, &exit))
    {
        /* :195-228 */
        const v3 hit_position_ms = enter > 0.0f ? tgo_v3_add(ray_origin_ms, tgo_v3_mulf(ray_direction_ms, enter)) : ray_origin_ms;
        const v3 voxel_center_ms = tgo_v3_add(voxel_min, tgo_v3(0.5f, 0.5f, 0.5f));
        v3 n = tgo_v3_sub(hit_position_ms, voxel_center_ms);
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
        normal_ws = tgo_v3_normalized(tgo_m4_mulv3w(p_object->rotation, n, 0.0f));
    }

    /* :231 -- un-normalised direction (Q3) */
    const v3 hit_position_ws = tgo_v3_add(ray_origin_ws, tgo_v3_mulf(ray_direction_ws, depth_24b * p_cam->far_plane));

    switch (debug_visualization)
    {
    case TG_DEBUG_SHOW_OBJECT_INDEX:    tgo__hash_color(object_idx, p_rgba); return;
    case TG_DEBUG_SHOW_DEPTH:
    {
        const f32 g = tgo_min(1.0f, 8.0f * depth_24b);
        p_rgba[0] = g; p_rgba[1] = g; p_rgba[2] = g; p_rgba[3] = 1.0f;
        return;
    }
    case TG_DEBUG_SHOW_CLUSTER_INDEX:
    case TG_DEBUG_SHOW_BLOCKS:          tgo__hash_color(cluster_idx, p_rgba); return;
    case TG_DEBUG_SHOW_VOXEL_INDEX:     tgo__hash_color(voxel_idx_9b, p_rgba); return;
    case TG_DEBUG_SHOW_COLOR_LUT_INDEX: tgo__hash_color(color_lut_idx, p_rgba); return;
    case TG_DEBUG_SHOW_COLOR:           p_rgba[0] = color_r; p_rgba[1] = color_g; p_rgba[2] = color_b; p_rgba[3] = 1.0f; return;
    case TG_DEBUG_SHOW_NORMAL:
        p_rgba[0] = normal_ws.x * 0.5f + 0.5f; p_rgba[1] = normal_ws.y * 0.5f + 0.5f; p_rgba[2] = normal_ws.z * 0.5f + 0.5f; p_rgba[3] = 1.0f;
        return;
    default: break;
    }

    /* :285-316 (SHADING uses a white albedo, NONE the LUT colour) */
    const f32 metallic = 0.1f;
    const v3 n = normal_ws;
    const v3 v = tgo_v3_normalized(tgo_v3_sub(ray_origin_ws, hit_position_ws));
    const v3 l = tgo_v3_normalized(tgo_v3(0.0f, 0.8f, 0.3f));
    const v3 albedo = debug_visualization == TG_DEBUG_SHOW_SHADING ? tgo_v3(1.0f, 1.0f, 1.0f) : tgo_v3(color_r, color_g, color_b);
    const v3 specular_albedo = tgo_v3_mix(tgo_v3(0.04f, 0.04f, 0.04f), albedo, metallic);
    const f32 roughness = 0.8f;
    const v3 radiance = tgo_v3(3.0f, 3.0f, 3.0f);

    const v3 lo = tgo__shade(n, v, l, albedo, specular_albedo, metallic, roughness, radiance);
    v3 ambient = tgo_v3_mulf(albedo, 0.1f);

    if (gi_enabled && p_svo && debug_visualization == TG_DEBUG_SHOW_NONE)
    {
        f32 visibility = 1.0f;
        if (normal_ws.x != 0.0f || normal_ws.y != 0.0f || normal_ws.z != 0.0f)
        {
            const u32 pixel_idx = w * py + px;
            u32 rng = tgo_hash_u32(pixel_idx ^ tgo_hash_u32(frame_seed)) | 1u;
            v3 dir = normal_ws;
            for (u32 attempt = 0; attempt < 32; attempt++)
            {
                v3 c;
                c.x = tgo_xorshift32_next_f32_range(&rng, -1.0f, 1.0f);
                c.y = tgo_xorshift32_next_f32_range(&rng, -1.0f, 1.0f);
                c.z = tgo_xorshift32_next_f32_range(&rng, -1.0f, 1.0f);
                c = tgo_v3_normalized(c);
                if (tgo_v3_dot(c, normal_ws) > 0.0f) { dir = c; break; }
            }
            const v3 origin = tgo_v3_add(hit_position_ws, tgo_v3_mulf(dir, 1.73205080757f));
            if (p_gi_ray_or_null)
            {
                p_gi_ray_or_null[0] = origin.x; p_gi_ray_or_null[1] = origin.y; p_gi_ray_or_null[2] = origin.z;
                p_gi_ray_or_null[3] = dir.x; p_gi_ray_or_null[4] = dir.y; p_gi_ray_or_null[5] = dir.z;
            }
            v3 hp, hn; u32 node_idx, voxel_idx;
            const f32 depth2 = tgo_svo_traverse_glsl(p_svo, p_cam->far_plane, origin, dir, &hp, &hn, &node_idx, &voxel_idx);
            visibility = depth2 < 1.0f ? 0.0f : 1.0f;
        }
        ambient = tgo_v3_mulf(ambient, visibility);
    }

    p_rgba[0] = ambient.x + lo.x;
    p_rgba[1] = ambient.y + lo.y;
    p_rgba[2] = ambient.z + lo.z;
    p_rgba[3] = 1.0f;
}

void tgo_shade(const tgo_scene_view* p_scene, const tg_camera_rays* p_cam, u32 w, u32 h, const u64* p_vis, const tg_svo* p_svo_or_null,
               u32 gi_enabled, u32 frame_seed, u32 debug_visualization, u32 y0, u32 y1, u32 ystep, f32* p_out_rgba)
{
    if (y1 > h) y1 = h;
    if (ystep == 0) ystep = 1;
    const i64 n_rows = y1 > y0 ? ((i64)(y1 - y0) + ystep - 1) / ystep : 0;
#pragma omp parallel for schedule(dynamic, 4)
    for (i64 row = 0; row < n_rows; row++)
    {
        const i64 py = (i64)y0 + row * ystep;
        for (u32 px = 0; px < w; px++)
        {
            const size_t i = (size_t)py * w + px;
            tgo__shade_pixel(p_scene, p_cam, w, h, p_vis[i], p_svo_or_null, gi_enabled, frame_seed, debug_visualization, px, (u32)py, &p_out_rgba[i * 4], NULL);
        }
    }
}

/*
 * The secondary rays tgo_shade traces (tgvk_raytracer.c:1405-1419 recipe), for tools that study the GI kernels' workload on the host:
 * p_out_rays[6 * pixel] = origin, direction; a pixel that shoots no ray (sky, zero normal) keeps direction (0, 0, 0).
 */
void tgo_shade_gi_rays(const tgo_scene_view* p_scene, const tg_camera_rays* p_cam, u32 w, u32 h, const u64* p_vis, const tg_svo* p_svo,
                       u32 frame_seed, u32 y0, u32 y1, u32 ystep, f32* p_out_rays)
{
    if (y1 > h) y1 = h;
    if (ystep == 0) ystep = 1;
    const i64 n_rows = y1 > y0 ? ((i64)(y1 - y0) + ystep - 1) / ystep : 0;
#pragma omp parallel for schedule(dynamic, 4)
    for (i64 row = 0; row < n_rows; row++)
    {
        const i64 py = (i64)y0 + row * ystep;
        for (u32 px = 0; px < w; px++)
        {
            const size_t i = (size_t)py * w + px;
            f32 rgba[4];
            for (u32 k = 0; k < 6; k++) p_out_rays[i * 6 + k] = 0.0f;
            tgo__shade_pixel(p_scene, p_cam, w, h, p_vis[i], p_svo, 1, frame_seed, TG_DEBUG_SHOW_NONE, px, (u32)py, rgba, &p_out_rays[i * 6]);
        }
    }
}

/*
 * present pass: present.frag:9-12 (out_color = texture(present_texture, v_uv), a 1:1 copy of the HDR target) followed by the
 * swapchain format's fixed-function conversion, VK_FORMAT_B8G8R8A8_UNORM (tgvk_core.c:4239-4246): NaN -> 0, clamp to [0, 1],
 * c * 255 rounded to nearest, ties to even (rintf under the default rounding mode). One u32 per pixel, bytes B, G, R, A.
 */
static u32 tgo__unorm8(f32 c)
{
    f32 v = c > 0.0f ? c : 0.0f;
    v = v > 1.0f ? 1.0f : v;
    return (u32)rintf(v * 255.0f);
}

void tgo_present_bgra8(const f32* p_rgba, u64 n_pixels, u32* p_out)
{
    for (u64 i = 0; i < n_pixels; i++)
    {
        const f32* c = &p_rgba[i * 4];
        p_out[i] = (tgo__unorm8(c[3]) << 24) | (tgo__unorm8(c[0]) << 16) | (tgo__unorm8(c[1]) << 8) | tgo__unorm8(c[2]);
    }
}

/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY. Nothing under tg_b200/ may include, link or call this.
 *
 * tgo_visibility.c: the reference's visibility pass as scalar C.
 *   camera            graphics/vulkan/tgvk_core.c:382-444, tgvk_raytracer.c:1171-1180
 *   clear             assets/shaders/raytracer/clear.comp:15-21
 *   fragment          assets/shaders/raytracer/visibility.frag:22-208
 *   ray/AABB          assets/shaders/raytracer/collide.inc:3-24
 *   ray interpolation assets/shaders/common.inc:40-46
 * Coverage (visibility.vert + rasteriser) only prunes fragments; the result is the min over all
 * clusters of the fragment function, which BRUTE_FORCE mode evaluates literally.
 */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "tgo.h"
#include "tgo_math.h"

i32 tgo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void tgo_set_threads(i32 n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void tgo_camera_rays(const tg_camera* p_camera, tg_camera_rays* p_out)
{
    /* tgvk_core.c:382-388: r = inverse(euler(pitch, yaw, roll)) */
    const m4 r = tgo_m4_inverse(tgo_m4_euler(p_camera->pitch, p_camera->yaw, p_camera->roll));
    /* tgvk_core.c:409-423 (perspective only on this path) */
    const m4 p = tgo_m4_perspective(p_camera->persp.fov_y_in_radians, p_camera->persp.aspect, p_camera->persp.n, p_camera->persp.f);
    /* tgvk_core.c:435-444 */
    const m4 ivp_no_translation = tgo_m4_inverse(tgo_m4_mul(p, r));
    const v4 c_bl = { -1.0f,  1.0f, 1.0f, 1.0f };
    const v4 c_br = {  1.0f,  1.0f, 1.0f, 1.0f };
    const v4 c_tr = {  1.0f, -1.0f, 1.0f, 1.0f };
    const v4 c_tl = { -1.0f, -1.0f, 1.0f, 1.0f };
    const v4 q_bl = tgo_m4_mulv4(ivp_no_translation, c_bl);
    const v4 q_br = tgo_m4_mulv4(ivp_no_translation, c_br);
    const v4 q_tr = tgo_m4_mulv4(ivp_no_translation, c_tr);
    const v4 q_tl = tgo_m4_mulv4(ivp_no_translation, c_tl);
    const v3 bl = tgo_v3_normalized(tgo_v3(q_bl.x, q_bl.y, q_bl.z));
    const v3 br = tgo_v3_normalized(tgo_v3(q_br.x, q_br.y, q_br.z));
    const v3 tr = tgo_v3_normalized(tgo_v3(q_tr.x, q_tr.y, q_tr.z));
    const v3 tl = tgo_v3_normalized(tgo_v3(q_tl.x, q_tl.y, q_tl.z));

    memset(p_out, 0, sizeof(*p_out));
    /* tgvk_raytracer.c:1174-1180 (w components stay 0) */
    p_out->camera.x = p_camera->position.x; p_out->camera.y = p_camera->position.y; p_out->camera.z = p_camera->position.z;
    p_out->ray_bl.x = bl.x; p_out->ray_bl.y = bl.y; p_out->ray_bl.z = bl.z;
    p_out->ray_br.x = br.x; p_out->ray_br.y = br.y; p_out->ray_br.z = br.z;
    p_out->ray_tr.x = tr.x; p_out->ray_tr.y = tr.y; p_out->ray_tr.z = tr.z;
    p_out->ray_tl.x = tl.x; p_out->ray_tl.y = tl.y; p_out->ray_tl.z = tl.z;
    p_out->near_plane = p_camera->persp.n;
    p_out->far_plane  = p_camera->persp.f;
}

void tgo_object_data(const tg_voxel_object* p_object, u32 lut_idx, tg_object_data* p_out)
{
    p_out->n_cluster_pointers_per_dim = p_object->n_cluster_pointers_per_dim;
    p_out->first_cluster_pointer      = p_object->first_cluster_pointer;
    p_out->translation                = p_object->translation;
    p_out->lut_idx                    = lut_idx;
    p_out->rotation                   = tgo_m4_angle_axis(p_object->angle_in_radians, p_object->axis);
}

u32 tgo_pack_color(f32 r, f32 g, f32 b)
{
    const u32 r_u32 = (u32)(r * 255.0f);
    const u32 g_u32 = (u32)(g * 255.0f);
    const u32 b_u32 = (u32)(b * 255.0f);
    return r_u32 << 24 | g_u32 << 16 | b_u32 << 8 | 255u;
}

m4 tgo_ws2ms(const tg_object_data* p_object, u32 cluster_pointer)
{
    /* visibility.frag:35-38 */
    const u32 rel   = cluster_pointer - p_object->first_cluster_pointer;
    const u32 rel_x = rel % p_object->n_cluster_pointers_per_dim.x;
    const u32 rel_y = (rel / p_object->n_cluster_pointers_per_dim.x) % p_object->n_cluster_pointers_per_dim.y;
    const u32 rel_z = rel / (p_object->n_cluster_pointers_per_dim.x * p_object->n_cluster_pointers_per_dim.y);
    /* visibility.frag:40-48 */
    const v3 relative_cluster_offset = tgo_v3((f32)(rel_x * 8u), (f32)(rel_y * 8u), (f32)(rel_z * 8u));
    const v3 cluster_half_extent = tgo_v3_mul(tgo_v3(8.0f, 8.0f, 8.0f), tgo_v3(0.5f, 0.5f, 0.5f));
    const v3 dims = tgo_v3((f32)p_object->n_cluster_pointers_per_dim.x, (f32)p_object->n_cluster_pointers_per_dim.y, (f32)p_object->n_cluster_pointers_per_dim.z);
    const v3 object_half_extent = tgo_v3_mul(cluster_half_extent, dims);
    /* visibility.frag:57-62: ws2ms3 * ws2ms2 * ws2ms1 * ws2ms0, left-associative */
    const m4 ws2ms0 = tgo_m4_translate(tgo_v3_neg(p_object->translation));
    const m4 ws2ms1 = tgo_m4_inverse(p_object->rotation);
    const m4 ws2ms2 = tgo_m4_translate(object_half_extent);
    const m4 ws2ms3 = tgo_m4_translate(tgo_v3_neg(relative_cluster_offset));
    return tgo_m4_mul(tgo_m4_mul(tgo_m4_mul(ws2ms3, ws2ms2), ws2ms1), ws2ms0);
}

v3 tgo_pixel_ray_direction_nn(const tg_camera_rays* p_cam, u32 w, u32 h, u32 px, u32 py)
{
    /* gl_FragCoord = pixel centre; visibility.frag:32-33 */
    const f32 frag_x = (f32)px + 0.5f;
    const f32 frag_y = (f32)py + 0.5f;
    const f32 fx =        frag_x / (f32)w;
    const f32 fy = 1.0f - frag_y / (f32)h;
    const v3 bl = tgo_v3(p_cam->ray_bl.x, p_cam->ray_bl.y, p_cam->ray_bl.z);
    const v3 br = tgo_v3(p_cam->ray_br.x, p_cam->ray_br.y, p_cam->ray_br.z);
    const v3 tr = tgo_v3(p_cam->ray_tr.x, p_cam->ray_tr.y, p_cam->ray_tr.z);
    const v3 tl = tgo_v3(p_cam->ray_tl.x, p_cam->ray_tl.y, p_cam->ray_tl.z);
    /* common.inc:40-46 */
    return tgo_v3_mix(tgo_v3_mix(bl, tl, fy), tgo_v3_mix(br, tr, fy), fx);
}

/* visibility.frag:83-191; `enter` is the cluster slab entry. Returns the voxel index or -1. */
i32 tgo_cluster_dda(const u32* p_mask16, v3 o_ms, v3 d_ms, f32 enter)
{
    const v3 cluster_min = tgo_v3(0.0f, 0.0f, 0.0f);
    const v3 cluster_max = tgo_v3(8.0f, 8.0f, 8.0f);

    const v3 hit = enter > 0.0f ? tgo_v3_add(o_ms, tgo_v3_mulf(d_ms, enter)) : o_ms;
    const v3 fl = tgo_v3_floor(hit);
    const v3 xyz = tgo_v3(
        tgo_clamp(fl.x, cluster_min.x, cluster_max.x - 1.0f),
        tgo_clamp(fl.y, cluster_min.y, cluster_max.y - 1.0f),
        tgo_clamp(fl.z, cluster_min.z, cluster_max.z - 1.0f));

    i32 x = (i32)xyz.x;
    i32 y = (i32)xyz.y;
    i32 z = (i32)xyz.z;

    i32 step_x = 0, step_y = 0, step_z = 0;
    f32 t_max_x = TG_F32_MAX, t_max_y = TG_F32_MAX, t_max_z = TG_F32_MAX;
    f32 t_delta_x = TG_F32_MAX, t_delta_y = TG_F32_MAX, t_delta_z = TG_F32_MAX;

    if (d_ms.x > 0.0f)
    {
        step_x = 1;
        t_max_x = enter + ((f32)(x + 1) - hit.x) / d_ms.x;
        t_delta_x = 1.0f / d_ms.x;
    }
    else if (d_ms.x < 0.0f)
    {
        step_x = -1;
        t_max_x = enter + (hit.x - (f32)x) / -d_ms.x;
        t_delta_x = 1.0f / -d_ms.x;
    }
    if (d_ms.y > 0.0f)
    {
        step_y = 1;
        t_max_y = enter + ((f32)(y + 1) - hit.y) / d_ms.y;
        t_delta_y = 1.0f / d_ms.y;
    }
    else if (d_ms.y < 0.0f)
    {
        step_y = -1;
        t_max_y = enter + (hit.y - (f32)y) / -d_ms.y;
        t_delta_y = 1.0f / -d_ms.y;
    }
    if (d_ms.z > 0.0f)
    {
        step_z = 1;
        t_max_z = enter + ((f32)(z + 1) - hit.z) / d_ms.z;
        t_delta_z = 1.0f / d_ms.z;
    }
    else if (d_ms.z < 0.0f)
    {
        step_z = -1;
        t_max_z = enter + (hit.z - (f32)z) / -d_ms.z;
        t_delta_z = 1.0f / -d_ms.z;
    }

    for (;;)
    {
        const u32 relative_voxel_idx = (u32)(64 * z + 8 * y + x);
        const u32 slot = p_mask16[relative_voxel_idx / 32];
        if ((slot & (1u << (relative_voxel_idx % 32))) != 0)
        {
            return (i32)relative_voxel_idx;
        }
        if (t_max_x < t_max_y)
        {
            if (t_max_x < t_max_z)
            {
                t_max_x += t_delta_x;
                x += step_x;
                if (x < 0 || x >= 8) break;
            }
            else
            {
                t_max_z += t_delta_z;
                z += step_z;
                if (z < 0 || z >= 8) break;
            }
        }
        else
        {
            if (t_max_y < t_max_z)
            {
                t_max_y += t_delta_y;
                y += step_y;
                if (y < 0 || y >= 8) break;
            }
            else
            {
                t_max_z += t_delta_z;
                z += step_z;
                if (z < 0 || z >= 8) break;
            }
        }
    }
    return -1;
}

/* visibility.frag:71-207 given the per-fragment ray in cluster space. */
static inline u64 tgo__fragment_core(const u32* p_mask16, v3 o_ms, v3 d_ms, f32 far_plane, u32 packed_pointer)
{
    const v3 cluster_min = tgo_v3(0.0f, 0.0f, 0.0f);
    const v3 cluster_max = tgo_v3(8.0f, 8.0f, 8.0f);

    f32 d = TG_F32_MAX;
    u32 voxel_idx = 0;
    f32 enter, exit;
    if (tgo_intersect_ray_aabb_glsl(o_ms, d_ms, cluster_min, cluster_max, &enter, &exit))
    {
        const i32 v = tgo_cluster_dda(p_mask16, o_ms, d_ms, enter);
        if (v >= 0)
        {
            /* visibility.frag:151-157 */
            const i32 x = v % 8, y = (v / 8) % 8, z = v / 64;
            const v3 voxel_min = tgo_v3_add(cluster_min, tgo_v3((f32)x, (f32)y, (f32)z));
            const v3 voxel_max = tgo_v3_add(cluster_min, tgo_v3((f32)(x + 1), (f32)(y + 1), (f32)(z + 1)));
            f32 voxel_enter, voxel_exit;
            tgo_intersect_ray_aabb_glsl(o_ms, d_ms, voxel_min, voxel_max, &voxel_enter, &voxel_exit);
            d = tgo_max(0.0f, voxel_enter / far_plane);
            voxel_idx = (u32)v;
        }
    }
    if (d <= 1.0f)
    {
        /* visibility.frag:198-201 */
        const u64 depth_24b           = (u64)(d * TG_VIS_DEPTH_SCALE) << TG_VIS_DEPTH_SHIFT;
        const u64 cluster_pointer_31b = (u64)packed_pointer << TG_VIS_POINTER_SHIFT;
        const u64 voxel_idx_9b        = (u64)voxel_idx;
        return depth_24b | cluster_pointer_31b | voxel_idx_9b;
    }
    return TG_VIS_CLEAR;
}

u64 tgo_visibility_fragment(const tgo_scene_view* p_scene, const tg_camera_rays* p_cam, u32 w, u32 h, u32 px, u32 py, u32 cluster_pointer)
{
    /* visibility.frag:24-27 */
    const u32 cluster_idx = p_scene->p_cluster_pointers[cluster_pointer];
    const u32 object_idx  = p_scene->p_cluster_idx_to_object_idx[cluster_idx];
    const tg_object_data* p_object = &p_scene->p_objects[object_idx];

    const m4 ws2ms = tgo_ws2ms(p_object, cluster_pointer);
    /* visibility.frag:65-69 */
    const v3 ray_origin_ws = tgo_v3(p_cam->camera.x, p_cam->camera.y, p_cam->camera.z);
    const v3 ray_origin_ms = tgo_m4_mulv3w(ws2ms, ray_origin_ws, 1.0f);
    const v3 ray_direction_ws = tgo_pixel_ray_direction_nn(p_cam, w, h, px, py);
    const v3 ray_direction_ms = tgo_v3_normalized(tgo_m4_mulv3w(ws2ms, ray_direction_ws, 0.0f));

    return tgo__fragment_core(&p_scene->p_voxel_cluster_data[(size_t)cluster_idx * TG_CLUSTER_MASK_WORDS], ray_origin_ms, ray_direction_ms,
                              p_cam->far_plane, cluster_pointer + p_scene->global_pointer_base);
}

/* -------------------------------------------------------------------------------------------
 * Frame driver. Hoists ws2ms / ray origin per cluster (identical arithmetic => identical bits,
 * SURVEY.md appendix B.1) and, in SCREEN_RECT mode, prunes pixels outside a conservative
 * rectangle of the cluster box (double precision pin-hole projection + margins).
 * ------------------------------------------------------------------------------------------- */

typedef struct tgo__cluster_setup
{
    m4  ws2ms;
    v3  o_ms;
    u32 cluster_pointer;
    u32 cluster_idx;
    i32 x0, y0, x1, y1; /* inclusive pixel rectangle */
} tgo__cluster_setup;

typedef struct tgo__pinhole
{
    f64 inv[9]; /* inverse of [u v bl] (row-major) */
    f64 cam[3];
    f64 rx[3], uy[3], fz[3]; /* orthonormal camera frame: right (bl->br), up (bl->tl), forward */
    f64 c0, bx, by, lu, lv;  /* plane distance, bl in the frame, |u|, |v| */
    u32 w, h;
    b32 ok;
} tgo__pinhole;

static void tgo__pinhole_init(const tg_camera_rays* c, u32 w, u32 h, tgo__pinhole* p)
{
    const f64 bl[3] = { c->ray_bl.x, c->ray_bl.y, c->ray_bl.z };
    const f64 u[3]  = { c->ray_br.x - bl[0], c->ray_br.y - bl[1], c->ray_br.z - bl[2] };
    const f64 v[3]  = { c->ray_tl.x - bl[0], c->ray_tl.y - bl[1], c->ray_tl.z - bl[2] };
    /* M = [u v bl] as columns */
    const f64 a = u[0], b = v[0], cc = bl[0];
    const f64 d = u[1], e = v[1], f = bl[1];
    const f64 g = u[2], hh = v[2], i = bl[2];
    const f64 det = a * (e * i - f * hh) - b * (d * i - f * g) + cc * (d * hh - e * g);
    p->ok = det != 0.0;
    const f64 id = p->ok ? 1.0 / det : 0.0;
    p->inv[0] = (e * i - f * hh) * id; p->inv[1] = (cc * hh - b * i) * id; p->inv[2] = (b * f - cc * e) * id;
    p->inv[3] = (f * g - d * i) * id;  p->inv[4] = (a * i - cc * g) * id;  p->inv[5] = (cc * d - a * f) * id;
    p->inv[6] = (d * hh - e * g) * id; p->inv[7] = (b * g - a * hh) * id;  p->inv[8] = (a * e - b * d) * id;
    p->cam[0] = c->camera.x; p->cam[1] = c->camera.y; p->cam[2] = c->camera.z;
    p->w = w; p->h = h;
    p->lu = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    p->lv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (!(p->lu > 0.0) || !(p->lv > 0.0)) { p->ok = TG_FALSE; return; }
    for (int k = 0; k < 3; k++) { p->rx[k] = u[k] / p->lu; p->uy[k] = v[k] / p->lv; }
    p->fz[0] = p->rx[1] * p->uy[2] - p->rx[2] * p->uy[1];
    p->fz[1] = p->rx[2] * p->uy[0] - p->rx[0] * p->uy[2];
    p->fz[2] = p->rx[0] * p->uy[1] - p->rx[1] * p->uy[0];
    f64 lf = sqrt(p->fz[0] * p->fz[0] + p->fz[1] * p->fz[1] + p->fz[2] * p->fz[2]);
    if (!(lf > 0.0)) { p->ok = TG_FALSE; return; }
    p->c0 = (bl[0] * p->fz[0] + bl[1] * p->fz[1] + bl[2] * p->fz[2]) / lf;
    if (p->c0 < 0.0) { lf = -lf; p->c0 = -p->c0; }
    for (int k = 0; k < 3; k++) p->fz[k] /= lf;
    p->bx = bl[0] * p->rx[0] + bl[1] * p->rx[1] + bl[2] * p->rx[2];
    p->by = bl[0] * p->uy[0] + bl[1] * p->uy[1] + bl[2] * p->uy[2];
    if (!(p->c0 > 0.0)) p->ok = TG_FALSE;
}

/* bounds of coordinate/depth over a circle of radius r around (a, z) in a plane through the eye; 0 = unbounded side */
static void tgo__tangent_bounds(f64 a, f64 z, f64 r, f64* p_lo, f64* p_hi, b32* p_lo_ok, b32* p_hi_ok)
{
    const f64 d2 = a * a + z * z;
    *p_lo_ok = TG_FALSE; *p_hi_ok = TG_FALSE; *p_lo = 0.0; *p_hi = 0.0;
    if (d2 <= r * r * 1.0001) return;
    const f64 theta = atan2(a, z), alpha = asin(r / sqrt(d2));
    const f64 lo = theta - alpha, hi = theta + alpha, lim = 1.5607963267948966; /* pi/2 - 0.01 */
    if (lo > -lim && lo < lim) { *p_lo = tan(lo); *p_lo_ok = TG_TRUE; }
    if (hi > -lim && hi < lim) { *p_hi = tan(hi); *p_hi_ok = TG_TRUE; }
    if (lo >= lim) { /* entirely beside/behind on the + side: nothing visible, keep conservative full range */ }
}

/* conservative pixel rectangle of a sphere; returns 0 when it is entirely behind the eye */
static b32 tgo__sphere_rect(const tgo__pinhole* p, const f64 C[3], f64 r, i32* x0, i32* y0, i32* x1, i32* y1)
{
    const f64 X = C[0] * p->rx[0] + C[1] * p->rx[1] + C[2] * p->rx[2];
    const f64 Y = C[0] * p->uy[0] + C[1] * p->uy[1] + C[2] * p->uy[2];
    const f64 Z = C[0] * p->fz[0] + C[1] * p->fz[1] + C[2] * p->fz[2];
    *x0 = 0; *y0 = 0; *x1 = (i32)p->w - 1; *y1 = (i32)p->h - 1;
    if (Z + r < 0.0) return TG_FALSE;
    f64 lo, hi; b32 lo_ok, hi_ok;
    tgo__tangent_bounds(X, Z, r, &lo, &hi, &lo_ok, &hi_ok);
    if (lo_ok) { const f64 px = ((p->c0 * lo - p->bx) / p->lu) * (f64)p->w - 0.5; const f64 q = floor(px) - 2.0; if (q > (f64)*x0) *x0 = q > 1e9 ? 1000000000 : (i32)q; }
    if (hi_ok) { const f64 px = ((p->c0 * hi - p->bx) / p->lu) * (f64)p->w - 0.5; const f64 q = ceil(px) + 2.0;  if (q < (f64)*x1) *x1 = q < -1e9 ? -1000000000 : (i32)q; }
    tgo__tangent_bounds(Y, Z, r, &lo, &hi, &lo_ok, &hi_ok);
    /* fy grows upwards, pixel rows grow downwards */
    if (hi_ok) { const f64 py = (1.0 - (p->c0 * hi - p->by) / p->lv) * (f64)p->h - 0.5; const f64 q = floor(py) - 2.0; if (q > (f64)*y0) *y0 = q > 1e9 ? 1000000000 : (i32)q; }
    if (lo_ok) { const f64 py = (1.0 - (p->c0 * lo - p->by) / p->lv) * (f64)p->h - 0.5; const f64 q = ceil(py) + 2.0;  if (q < (f64)*y1) *y1 = q < -1e9 ? -1000000000 : (i32)q; }
    return TG_TRUE;
}

/* Returns 0 if the point is not safely in front of the camera. */
static b32 tgo__pinhole_project(const tgo__pinhole* p, const f64 X[3], f64* p_px, f64* p_py)
{
    const f64 r[3] = { X[0] - p->cam[0], X[1] - p->cam[1], X[2] - p->cam[2] };
    const f64 a = p->inv[0] * r[0] + p->inv[1] * r[1] + p->inv[2] * r[2];
    const f64 b = p->inv[3] * r[0] + p->inv[4] * r[1] + p->inv[5] * r[2];
    const f64 c = p->inv[6] * r[0] + p->inv[7] * r[1] + p->inv[8] * r[2];
    const f64 len = fabs(r[0]) + fabs(r[1]) + fabs(r[2]);
    if (!(c > 1e-4 * len) || !(c > 1e-9)) return TG_FALSE;
    const f64 fx = a / c, fy = b / c;
    *p_px = fx * (f64)p->w - 0.5;
    *p_py = (1.0 - fy) * (f64)p->h - 0.5;
    return TG_TRUE;
}

static void tgo__cluster_rect(const tgo__pinhole* p, const tg_object_data* o, u32 cluster_pointer, f32 far_plane, tgo__cluster_setup* s)
{
    const u32 rel   = cluster_pointer - o->first_cluster_pointer;
    const u32 rel_x = rel % o->n_cluster_pointers_per_dim.x;
    const u32 rel_y = (rel / o->n_cluster_pointers_per_dim.x) % o->n_cluster_pointers_per_dim.y;
    const u32 rel_z = rel / (o->n_cluster_pointers_per_dim.x * o->n_cluster_pointers_per_dim.y);
    const f64 half[3] = { 4.0 * o->n_cluster_pointers_per_dim.x, 4.0 * o->n_cluster_pointers_per_dim.y, 4.0 * o->n_cluster_pointers_per_dim.z };
    const f64 off[3]  = { 8.0 * rel_x, 8.0 * rel_y, 8.0 * rel_z };
    const f64 R[9] = { o->rotation.m00, o->rotation.m01, o->rotation.m02,
                       o->rotation.m10, o->rotation.m11, o->rotation.m12,
                       o->rotation.m20, o->rotation.m21, o->rotation.m22 };
    const f64 margin = 0.0625;
    f64 minx = 1e300, miny = 1e300, maxx = -1e300, maxy = -1e300;
    b32 full = !p->ok;
    f64 centre_ws[3] = { 0, 0, 0 };
    for (u32 k = 0; k < 8 && !full; k++)
    {
        const f64 c[3] = {
            ((k & 1) ? 8.0 + margin : -margin) + off[0] - half[0],
            ((k & 2) ? 8.0 + margin : -margin) + off[1] - half[1],
            ((k & 4) ? 8.0 + margin : -margin) + off[2] - half[2] };
        const f64 X[3] = {
            R[0] * c[0] + R[1] * c[1] + R[2] * c[2] + o->translation.x,
            R[3] * c[0] + R[4] * c[1] + R[5] * c[2] + o->translation.y,
            R[6] * c[0] + R[7] * c[1] + R[8] * c[2] + o->translation.z };
        centre_ws[0] += X[0] / 8.0; centre_ws[1] += X[1] / 8.0; centre_ws[2] += X[2] / 8.0;
        f64 px, py;
        if (!tgo__pinhole_project(p, X, &px, &py)) { full = TG_TRUE; break; }
        if (px < minx) minx = px;
        if (px > maxx) maxx = px;
        if (py < miny) miny = py;
        if (py > maxy) maxy = py;
    }
    if (full)
    {
        /* some corner is beside/behind the camera: keep every pixel unless the whole box is out of range */
        s->x0 = 0; s->y0 = 0; s->x1 = (i32)p->w - 1; s->y1 = (i32)p->h - 1;
        f64 cx[3] = { 4.0 + off[0] - half[0], 4.0 + off[1] - half[1], 4.0 + off[2] - half[2] };
        f64 C[3] = { R[0] * cx[0] + R[1] * cx[1] + R[2] * cx[2] + o->translation.x - p->cam[0],
                     R[3] * cx[0] + R[4] * cx[1] + R[5] * cx[2] + o->translation.y - p->cam[1],
                     R[6] * cx[0] + R[7] * cx[1] + R[8] * cx[2] + o->translation.z - p->cam[2] };
        const f64 dist = sqrt(C[0] * C[0] + C[1] * C[1] + C[2] * C[2]);
        if (dist - 7.2 > (f64)far_plane * 1.001) { s->x1 = -1; return; }
        /* bounding sphere of the inflated box: radius sqrt(3) * (4 + margin) */
        if (p->ok && !tgo__sphere_rect(p, C, 1.7320508075688772 * (4.0 + margin) + 1e-3, &s->x0, &s->y0, &s->x1, &s->y1)) { s->x1 = -1; }
        return;
    }
    {
        const f64 C[3] = { centre_ws[0] - p->cam[0], centre_ws[1] - p->cam[1], centre_ws[2] - p->cam[2] };
        const f64 dist = sqrt(C[0] * C[0] + C[1] * C[1] + C[2] * C[2]);
        if (dist - 7.2 > (f64)far_plane * 1.001) { s->x0 = 0; s->y0 = 0; s->x1 = -1; s->y1 = -1; return; }
    }
    minx = floor(minx) - 2.0; miny = floor(miny) - 2.0; maxx = ceil(maxx) + 2.0; maxy = ceil(maxy) + 2.0;
    if (minx < 0.0) minx = 0.0;
    if (miny < 0.0) miny = 0.0;
    if (maxx > (f64)p->w - 1.0) maxx = (f64)p->w - 1.0;
    if (maxy > (f64)p->h - 1.0) maxy = (f64)p->h - 1.0;
    if (maxx < minx || maxy < miny) { s->x0 = 0; s->y0 = 0; s->x1 = -1; s->y1 = -1; return; }
    s->x0 = (i32)minx; s->y0 = (i32)miny; s->x1 = (i32)maxx; s->y1 = (i32)maxy;
}

u64 tgo_visibility(const tgo_scene_view* p_scene, const tg_camera_rays* p_cam, u32 w, u32 h, u32 mode, u32 y0, u32 y1, u32 ystep, u64* p_out)
{
    if (y1 > h) y1 = h;
    if (ystep == 0) ystep = 1;
    /* clear.comp:19 */
    for (size_t i = 0; i < (size_t)w * h; i++) p_out[i] = TG_VIS_CLEAR;

    const u32 n = p_scene->n_cluster_pointers;
    tgo__cluster_setup* p_setup = (tgo__cluster_setup*)malloc((size_t)(n ? n : 1) * sizeof(*p_setup));
    u32 n_setup = 0;

    tgo__pinhole pin;
    tgo__pinhole_init(p_cam, w, h, &pin);
    const v3 ray_origin_ws = tgo_v3(p_cam->camera.x, p_cam->camera.y, p_cam->camera.z);

    /* per-cluster setup (cheap; serial compaction keeps the order deterministic) */
    tgo__cluster_setup* p_all = (tgo__cluster_setup*)malloc((size_t)(n ? n : 1) * sizeof(*p_all));
#pragma omp parallel for schedule(static)
    for (i64 cp = 0; cp < (i64)n; cp++)
    {
        tgo__cluster_setup* s = &p_all[cp];
        const u32 cluster_idx = p_scene->p_cluster_pointers[cp];
        const u32 object_idx  = p_scene->p_cluster_idx_to_object_idx[cluster_idx];
        const tg_object_data* p_object = &p_scene->p_objects[object_idx];
        s->cluster_pointer = (u32)cp;
        s->cluster_idx = cluster_idx;
        s->x0 = 0; s->y0 = 0; s->x1 = (i32)w - 1; s->y1 = (i32)h - 1;
        if (mode == TGO_VIS_SCREEN_RECT)
        {
            /* an all-zero mask can never produce a write (appendix B.2) */
            const u32* m = &p_scene->p_voxel_cluster_data[(size_t)cluster_idx * TG_CLUSTER_MASK_WORDS];
            u32 any = 0;
            for (u32 k = 0; k < TG_CLUSTER_MASK_WORDS; k++) any |= m[k];
            if (!any) { s->x1 = -1; continue; }
            tgo__cluster_rect(&pin, p_object, (u32)cp, p_cam->far_plane, s);
            if (s->x1 < s->x0 || s->y1 < s->y0) { s->x1 = -1; continue; }
        }
        s->ws2ms = tgo_ws2ms(p_object, (u32)cp);
        s->o_ms  = tgo_m4_mulv3w(s->ws2ms, ray_origin_ws, 1.0f);
    }
    for (u32 cp = 0; cp < n; cp++)
    {
        if (p_all[cp].x1 >= p_all[cp].x0) p_setup[n_setup++] = p_all[cp];
    }
    free(p_all);

    u64 n_fragments = 0;
    const i64 n_rows = ((i64)y1 - (i64)y0 + ystep - 1) / ystep;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : n_fragments)
    for (i64 row = 0; row < n_rows; row++)
    {
        const u32 py = y0 + (u32)row * ystep;
        v3* p_dir_ws = (v3*)malloc((size_t)w * sizeof(v3));
        for (u32 px = 0; px < w; px++) p_dir_ws[px] = tgo_pixel_ray_direction_nn(p_cam, w, h, px, py);
        u64* p_row = &p_out[(size_t)py * w];
        for (u32 i = 0; i < n_setup; i++)
        {
            const tgo__cluster_setup* s = &p_setup[i];
            if ((i32)py < s->y0 || (i32)py > s->y1) continue;
            const u32* p_mask = &p_scene->p_voxel_cluster_data[(size_t)s->cluster_idx * TG_CLUSTER_MASK_WORDS];
            const u32 packed_pointer = s->cluster_pointer + p_scene->global_pointer_base;
            for (i32 px = s->x0; px <= s->x1; px++)
            {
                const v3 d_ms = tgo_v3_normalized(tgo_m4_mulv3w(s->ws2ms, p_dir_ws[px], 0.0f));
                const u64 word = tgo__fragment_core(p_mask, s->o_ms, d_ms, p_cam->far_plane, packed_pointer);
                if (word < p_row[px]) p_row[px] = word; /* atomicMin, visibility.frag:206 */
                n_fragments++;
            }
        }
        free(p_dir_ws);
    }
    free(p_setup);
    return n_fragments;
}

/*
 * Brute force over a pixel WINDOW [x0, x1) x [y0, y1): every pixel of the window against every cluster pointer of the scene through
 * tgo_visibility_fragment (visibility.frag:22-208, nothing hoisted, nothing pruned). The other pixels keep the clear value. For
 * full-resolution frames, where brute force over whole scanlines is out of reach: the screen-rectangle pruning of TGO_VIS_SCREEN_RECT
 * (the same design as the CUDA path's object cull) is then not the only witness.
 */
u64 tgo_visibility_window(const tgo_scene_view* p_scene, const tg_camera_rays* p_cam, u32 w, u32 h, u32 x0, u32 x1, u32 y0, u32 y1, u64* p_out)
{
    if (x1 > w) x1 = w;
    if (y1 > h) y1 = h;
    for (size_t i = 0; i < (size_t)w * h; i++) p_out[i] = TG_VIS_CLEAR;
    const u32 n = p_scene->n_cluster_pointers;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (i64 py = (i64)y0; py < (i64)y1; py++)
    {
        for (i64 px = (i64)x0; px < (i64)x1; px++)
        {
            u64 best = TG_VIS_CLEAR;
            for (u32 cp = 0; cp < n; cp++)
            {
                const u64 word = tgo_visibility_fragment(p_scene, p_cam, w, h, (u32)px, (u32)py, cp);
                if (word < best) best = word;
            }
            p_out[(size_t)py * w + (size_t)px] = best;
        }
    }
    return (u64)(x1 > x0 ? x1 - x0 : 0) * (u64)(y1 > y0 ? y1 - y0 : 0) * n;
}

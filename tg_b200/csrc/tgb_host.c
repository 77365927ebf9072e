/*
 * tgb_host.c -- the C host side of libtgb200: the reference's raytracer entry points
 * (/root/reference/tg/src/graphics/vulkan/tgvk_raytracer.h:240-251) and its scene bookkeeping
 * (tgvk_raytracer.c:662-712, 805-866, 994-1077, 1122-1142), re-stated over the CUDA seam of
 * tgb_internal.h. No Vulkan, no CUDA types; single-threaded like the reference.
 */
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tgb_internal.h"
#include "tgb_rows.h"
#include "tgb_math.h"

/* ---- error recording ---------------------------------------------------------------------- */

static char tgb__error[1024];
static b32  tgb__has_error = TG_FALSE;

void tgb_set_error(const char* p_fmt, ...)
{
    va_list args;
    va_start(args, p_fmt);
    vsnprintf(tgb__error, sizeof(tgb__error), p_fmt, args);
    va_end(args);
    tgb__has_error = TG_TRUE;
    if (getenv("TGB200_VERBOSE")) fprintf(stderr, "[tgb200] %s\n", tgb__error);
}

const char* tgb200_last_error(void) { return tgb__has_error ? tgb__error : NULL; }
void tgb200_clear_error(void) { tgb__has_error = TG_FALSE; tgb__error[0] = 0; }

#define TGB_REQUIRE(cond, ret, ...) do { if (!(cond)) { tgb_set_error(__VA_ARGS__); return ret; } } while (0)
#define TGB_VOID

static i32 tgb__device = 0;

static u32 tgb__default_w = 1920, tgb__default_h = 1080;

i32  tgb200_device_count(void) { return tgbd_device_count(); }
void tgb200_set_device(i32 device) { tgb__device = device; }
void tgb200_set_default_resolution(u32 width, u32 height) { tgb__default_w = width; tgb__default_h = height; }

/* ---- pure host logic: scene bookkeeping ------------------------------------------------------ */

void tgb200_scene_init(tg_scene* p_scene, u32 max_n_objects, u32 max_n_clusters)
{
    memset(p_scene, 0, sizeof(*p_scene));
    /* tgvk_raytracer.c:687-700; TG_MALLOC zero-fills (platform/tg_platform_win32.c:133) */
    p_scene->object_capacity             = max_n_objects;
    p_scene->p_objects                   = (tg_voxel_object*)calloc(max_n_objects, sizeof(tg_voxel_object));
    p_scene->n_available_object_indices  = max_n_objects;
    p_scene->p_available_object_indices  = (u32*)calloc(max_n_objects, sizeof(u32));
    p_scene->cluster_pointer_capacity    = max_n_clusters;
    p_scene->p_cluster_pointers          = (u32*)calloc(max_n_clusters, sizeof(u32));
    p_scene->n_available_cluster_indices = max_n_clusters;
    p_scene->p_available_cluster_indices = (u32*)calloc(max_n_clusters, sizeof(u32));
    p_scene->p_voxel_cluster_data        = (u32*)calloc((size_t)max_n_clusters * TG_CLUSTER_MASK_WORDS, sizeof(u32));
    p_scene->p_cluster_idx_to_object_idx = (u32*)calloc(max_n_clusters, sizeof(u32));
    /* tgvk_raytracer.c:704-712: stacks filled descending, so pops yield 0, 1, 2, ... */
    for (u32 i = 0; i < max_n_objects; i++) p_scene->p_available_object_indices[i] = max_n_objects - i - 1;
    for (u32 i = 0; i < max_n_clusters; i++) p_scene->p_available_cluster_indices[i] = max_n_clusters - i - 1;
}

void tgb200_scene_free(tg_scene* p_scene)
{
    free(p_scene->p_objects);
    free(p_scene->p_available_object_indices);
    free(p_scene->p_cluster_pointers);
    free(p_scene->p_available_cluster_indices);
    free(p_scene->p_voxel_cluster_data);
    free(p_scene->p_cluster_idx_to_object_idx);
    memset(p_scene, 0, sizeof(*p_scene));
}

b32 tg_object_is_initialized(const tg_scene* p_scene, u32 object_idx)
{
    TGB_REQUIRE(p_scene && object_idx < p_scene->object_capacity, TG_FALSE, "tg_object_is_initialized: object %u out of range", object_idx);
    const tg_voxel_object* p_object = &p_scene->p_objects[object_idx];
    return p_object->n_cluster_pointers_per_dim.x != 0 && p_object->n_cluster_pointers_per_dim.y != 0 && p_object->n_cluster_pointers_per_dim.z != 0;
}

/*
 * tgm_m4_angle_axis (math/tg_math.c:1870-1910) does not normalise its axis: a non-unit axis yields a scaled / sheared matrix. The
 * reference's per-cluster rasteriser would still draw such an object; the object-level culling here (distance and screen-rectangle
 * bounds of the rigid box, k_cull_objects / k_svo_object_flags) assumes an orthonormal rotation, so the axis is checked where it
 * enters (the reference only ever passes (0,1,0), tgvk_raytracer.c:841).
 */
static b32 tgb__axis_is_unit(v3 axis)
{
    const f64 l2 = (f64)axis.x * axis.x + (f64)axis.y * axis.y + (f64)axis.z * axis.z;
    return l2 > 1.0 - 2e-5 && l2 < 1.0 + 2e-5;
}

u32 tgb200_scene_alloc_object(tg_scene* p_scene, v3 center, v3u extent, f32 angle_in_radians, v3 axis)
{
    /* tgvk_raytracer.c:807-812 */
    TGB_REQUIRE(extent.x % 8 == 0 && extent.y % 8 == 0 && extent.z % 8 == 0 && extent.x && extent.y && extent.z, TG_U32_MAX,
                "create_object: extent (%u,%u,%u) must be non-zero multiples of 8", extent.x, extent.y, extent.z);
    TGB_REQUIRE(tgb__axis_is_unit(axis), TG_U32_MAX, "create_object: rotation axis (%g,%g,%g) is not a unit vector", (double)axis.x, (double)axis.y, (double)axis.z);
    TGB_REQUIRE(p_scene->n_objects < p_scene->object_capacity && p_scene->n_available_object_indices > 0, TG_U32_MAX, "create_object: object capacity %u exhausted", p_scene->object_capacity);
    const v3u dims = { extent.x / 8, extent.y / 8, extent.z / 8 };
    const u64 n64 = (u64)dims.x * dims.y * dims.z;
    TGB_REQUIRE(n64 <= p_scene->n_available_cluster_indices && (u64)p_scene->n_cluster_pointers + n64 <= p_scene->cluster_pointer_capacity, TG_U32_MAX,
                "create_object: %llu clusters do not fit (capacity %u, used %u)", (unsigned long long)n64, p_scene->cluster_pointer_capacity, p_scene->n_cluster_pointers);
    const u32 n_cluster_pointers = (u32)n64;

    /* tgvk_raytracer.c:816-832 */
    const u32 object_idx = p_scene->p_available_object_indices[--(p_scene->n_available_object_indices)];
    p_scene->n_objects++;
    tg_voxel_object* p_object = &p_scene->p_objects[object_idx];
    p_object->n_cluster_pointers_per_dim = dims;
    p_object->first_cluster_pointer = p_scene->n_cluster_pointers;
    p_scene->n_cluster_pointers += n_cluster_pointers;
    p_object->translation = center;
    p_object->angle_in_radians = angle_in_radians;
    p_object->axis = axis;

    /* tgvk_raytracer.c:851-866 */
    for (u32 rel = 0; rel < n_cluster_pointers; rel++)
    {
        const u32 cluster_pointer = p_object->first_cluster_pointer + rel;
        const u32 cluster_idx = p_scene->p_available_cluster_indices[--(p_scene->n_available_cluster_indices)];
        p_scene->p_cluster_pointers[cluster_pointer] = cluster_idx;
        p_scene->p_cluster_idx_to_object_idx[cluster_idx] = object_idx;
    }
    return object_idx;
}

void tgb200_scene_free_object(tg_scene* p_scene, u32 object_idx, u32* p_first_shifted_pointer, u32* p_n_shifted)
{
    if (p_first_shifted_pointer) *p_first_shifted_pointer = 0;
    if (p_n_shifted) *p_n_shifted = 0;
    TGB_REQUIRE(object_idx < p_scene->object_capacity && tg_object_is_initialized(p_scene, object_idx), TGB_VOID, "destroy_object: object %u is not initialised", object_idx);
    tg_voxel_object* p_object = &p_scene->p_objects[object_idx];
    const u32 n_cluster_pointers = p_object->n_cluster_pointers_per_dim.x * p_object->n_cluster_pointers_per_dim.y * p_object->n_cluster_pointers_per_dim.z;
    const u32 one_past_last = p_object->first_cluster_pointer + n_cluster_pointers;

    /* tgvk_raytracer.c:1011-1015: return the cluster indices (ascending pointer order) */
    for (u32 cp = p_object->first_cluster_pointer; cp < one_past_last; cp++)
    {
        p_scene->p_available_cluster_indices[p_scene->n_available_cluster_indices++] = p_scene->p_cluster_pointers[cp];
    }

    /* tgvk_raytracer.c:1019-1059: compact the pointer table, fix every later object's first pointer */
    if (one_past_last < p_scene->n_cluster_pointers)
    {
        u32 first = one_past_last;
        while (first < p_scene->n_cluster_pointers)
        {
            const u32 first_cluster_idx = p_scene->p_cluster_pointers[first];
            const u32 object_idx_to_mod = p_scene->p_cluster_idx_to_object_idx[first_cluster_idx];
            tg_voxel_object* p_mod = &p_scene->p_objects[object_idx_to_mod];
            p_mod->first_cluster_pointer -= n_cluster_pointers;
            first += p_mod->n_cluster_pointers_per_dim.x * p_mod->n_cluster_pointers_per_dim.y * p_mod->n_cluster_pointers_per_dim.z;
        }
        memmove(&p_scene->p_cluster_pointers[one_past_last - n_cluster_pointers], &p_scene->p_cluster_pointers[one_past_last],
                (size_t)(p_scene->n_cluster_pointers - one_past_last) * sizeof(u32));
        if (p_first_shifted_pointer) *p_first_shifted_pointer = one_past_last - n_cluster_pointers;
        if (p_n_shifted) *p_n_shifted = p_scene->n_cluster_pointers - one_past_last;
    }
    p_scene->n_cluster_pointers -= n_cluster_pointers;

    /* tgvk_raytracer.c:1061-1066; the record is zeroed (the reference does so in debug builds) so is_initialized turns false */
    p_scene->p_available_object_indices[p_scene->n_available_object_indices++] = object_idx;
    memset(p_object, 0, sizeof(*p_object));
    p_scene->n_objects--;
}

void tgb200_camera_rays(const tg_camera* p_camera, tg_camera_rays* p_out)
{
    /* tgvk_core.c:382-388, 409-423, 435-444 */
    const m4 r = tgb_m4_inverse(tgb_m4_euler(p_camera->pitch, p_camera->yaw, p_camera->roll));
    const m4 p = tgb_m4_perspective(p_camera->persp.fov_y_in_radians, p_camera->persp.aspect, p_camera->persp.n, p_camera->persp.f);
    const m4 ivp_no_translation = tgb_m4_inverse(tgb_m4_mul(p, r));
    const f32 sx[4] = { -1.0f, 1.0f, 1.0f, -1.0f };
    const f32 sy[4] = { 1.0f, 1.0f, -1.0f, -1.0f }; /* bl, br, tr, tl */
    v3 rays[4];
    for (int i = 0; i < 4; i++)
    {
        /* tgm_m4_mulv4 with w = 1 (math/tg_math.c:2428-2438), then tgm_v3_normalized */
        rays[i] = tgb_normalize(tgb_m4_transform(ivp_no_translation, tgb_v3(sx[i], sy[i], 1.0f), 1.0f));
    }
    memset(p_out, 0, sizeof(*p_out));
    p_out->camera.x = p_camera->position.x; p_out->camera.y = p_camera->position.y; p_out->camera.z = p_camera->position.z;
    p_out->ray_bl.x = rays[0].x; p_out->ray_bl.y = rays[0].y; p_out->ray_bl.z = rays[0].z;
    p_out->ray_br.x = rays[1].x; p_out->ray_br.y = rays[1].y; p_out->ray_br.z = rays[1].z;
    p_out->ray_tr.x = rays[2].x; p_out->ray_tr.y = rays[2].y; p_out->ray_tr.z = rays[2].z;
    p_out->ray_tl.x = rays[3].x; p_out->ray_tl.y = rays[3].y; p_out->ray_tl.z = rays[3].z;
    p_out->near_plane = p_camera->persp.n;
    p_out->far_plane = p_camera->persp.f;
}

void tgb200_object_data(const tg_scene* p_scene, u32 object_idx, u32 lut_idx, tg_object_data* p_out)
{
    const tg_voxel_object* p_object = &p_scene->p_objects[object_idx];
    memset(p_out, 0, sizeof(*p_out));
    p_out->n_cluster_pointers_per_dim = p_object->n_cluster_pointers_per_dim;
    p_out->first_cluster_pointer = p_object->first_cluster_pointer;
    p_out->translation = p_object->translation;
    p_out->lut_idx = lut_idx;
    if (tg_object_is_initialized(p_scene, object_idx)) p_out->rotation = tgb_m4_angle_axis(p_object->angle_in_radians, p_object->axis);
}

u32 tgb200_pack_color(f32 r, f32 g, f32 b)
{
    const u32 r_u32 = (u32)(r * 255.0f);
    const u32 g_u32 = (u32)(g * 255.0f);
    const u32 b_u32 = (u32)(b * 255.0f);
    return r_u32 << 24 | g_u32 << 16 | b_u32 << 8 | 255u;
}

/* ---- raytracer ------------------------------------------------------------------------------- */

#define TGB_MAX_N_CLUSTERS_31B (1u << 31)

static b32 tgb__alive(const tg_raytracer* p_raytracer, const char* p_what)
{
    if (!p_raytracer) { tgb_set_error("%s: NULL raytracer", p_what); return TG_FALSE; }
    if (!p_raytracer->p_device) { tgb_set_error("%s: raytracer has no device state (creation failed; no CPU fallback)", p_what); return TG_FALSE; }
    return TG_TRUE;
}

void tg_raytracer_create(const tg_camera* p_camera, u32 max_n_objects, u32 max_n_clusters, tg_raytracer* p_raytracer)
{
    TGB_REQUIRE(p_raytracer != NULL, TGB_VOID, "tg_raytracer_create: NULL out pointer");
    memset(p_raytracer, 0, sizeof(*p_raytracer));
    TGB_REQUIRE(p_camera != NULL, TGB_VOID, "tg_raytracer_create: NULL camera");
    TGB_REQUIRE(max_n_objects > 0 && max_n_clusters > 0, TGB_VOID, "tg_raytracer_create: capacities must be > 0");
    /* the reference caps both at 2^21 (tgvk_raytracer.c:21,666-668); lifted to the word's 31 pointer bits (Q12) */
    TGB_REQUIRE(max_n_clusters <= TGB_MAX_N_CLUSTERS_31B - 1, TGB_VOID, "tg_raytracer_create: at most 2^31-1 clusters are addressable");

    p_raytracer->n_color_luts = 1;
    struct tgb_device* p_device = tgbd_create(tgb__device, max_n_objects, max_n_clusters, 256, tgb__default_w, tgb__default_h);
    if (!p_device) return; /* error recorded; fail loudly, no fallback */
    p_raytracer->n_color_luts = 256;

    p_raytracer->p_camera = p_camera;
    p_raytracer->p_device = p_device;
    p_raytracer->width = tgb__default_w;
    p_raytracer->height = tgb__default_h;
    p_raytracer->frame_seed = 1;
    tgb200_scene_init(&p_raytracer->scene, max_n_objects, max_n_clusters);
    p_raytracer->p_object_lut_idx = (u32*)calloc(max_n_objects, sizeof(u32));
    p_raytracer->p_moved_objects = (u32*)calloc(max_n_objects, sizeof(u32));
    p_raytracer->svo_dirty = 2;
}

void tg_raytracer_destroy(tg_raytracer* p_raytracer)
{
    if (!p_raytracer) return;
    tgb200_comm_destroy(p_raytracer);
    if (p_raytracer->scene.svo.p_node_buffer) tg_svo_destroy(&p_raytracer->scene.svo);
    if (p_raytracer->p_device) tgbd_destroy(p_raytracer->p_device);
    if (p_raytracer->scene.p_objects) tgb200_scene_free(&p_raytracer->scene);
    free(p_raytracer->p_object_lut_idx);
    free(p_raytracer->p_moved_objects);
    memset(p_raytracer, 0, sizeof(*p_raytracer));
}

void tg_raytracer_set_debug_visualization(tg_raytracer* p_raytracer, tg_debug_show type)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_set_debug_visualization")) return;
    p_raytracer->debug_visualization = (u32)type;
}

void tg_raytracer_set_resolution(tg_raytracer* p_raytracer, u32 width, u32 height)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_set_resolution")) return;
    TGB_REQUIRE(width > 0 && height > 0, TGB_VOID, "tg_raytracer_set_resolution: empty resolution");
    if (tgbd_resize(p_raytracer->p_device, width, height)) { p_raytracer->width = width; p_raytracer->height = height; }
}

void tg_raytracer_set_gi(tg_raytracer* p_raytracer, b32 enabled, u32 frame_seed)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_set_gi")) return;
    p_raytracer->gi_enabled = enabled ? 1 : 0;
    p_raytracer->frame_seed = frame_seed;
}

static b32 tgb__upload_object_record(tg_raytracer* p_raytracer, u32 object_idx)
{
    tg_object_data rec;
    tgb200_object_data(&p_raytracer->scene, object_idx, p_raytracer->p_object_lut_idx[object_idx], &rec);
    return tgbd_upload(p_raytracer->p_device, TGB_BUF_OBJECTS, (u64)object_idx * sizeof(tg_object_data), &rec, sizeof(rec));
}

/* Exact inverse of tgb200_scene_alloc_object for the object allocated LAST (its pointers end the table): both free-lists get their
 * entries back in the order they were popped, so a failed create leaves the scene as it found it. */
static void tgb__scene_undo_alloc(tg_scene* p_scene, u32 object_idx)
{
    tg_voxel_object* p_object = &p_scene->p_objects[object_idx];
    const u32 n = p_object->n_cluster_pointers_per_dim.x * p_object->n_cluster_pointers_per_dim.y * p_object->n_cluster_pointers_per_dim.z;
    for (u32 rel = n; rel-- > 0;)
        p_scene->p_available_cluster_indices[p_scene->n_available_cluster_indices++] = p_scene->p_cluster_pointers[p_object->first_cluster_pointer + rel];
    p_scene->n_cluster_pointers -= n;
    p_scene->p_available_object_indices[p_scene->n_available_object_indices++] = object_idx;
    memset(p_object, 0, sizeof(*p_object));
    p_scene->n_objects--;
}

/* Are the object's cluster indices one ascending run? (always true for scenes without destroys) */
static b32 tgb__contiguous_run(const tg_scene* p_scene, const tg_voxel_object* p_object, u32 n)
{
    const u32* p = &p_scene->p_cluster_pointers[p_object->first_cluster_pointer];
    return n == 0 || (p[n - 1] - p[0] == n - 1 && p[n - 1] >= p[0]);
}

/*
 * Shared tail of the two object constructors: scene bookkeeping (tgvk_raytracer.c:816-866), device mirrors of the object
 * record / pointer range / cluster->object map, then the voxel bits -- given by the caller, or (p_solid_bits == NULL) the
 * reference's procedural terrain generated ON THE DEVICE straight into the resident mask array (tgb_procedural.cu; the
 * reference's own note at tgvk_raytracer.c:868 is "TODO: gen on GPU") and read back into the scene's CPU mirror
 * p_voxel_cluster_data, which the reference keeps too (:934).
 */
static u32 tgb__create_object(tg_raytracer* p_raytracer, v3 center, v3u extent, f32 angle_in_radians, v3 axis, u32 lut_idx,
                              const u32* p_solid_bits, const u8* p_lut_indices, u32 synthetic_k, u32 synthetic_seed)
{
    tg_scene* p_scene = &p_raytracer->scene;
    const u32 object_idx = tgb200_scene_alloc_object(p_scene, center, extent, angle_in_radians, axis);
    if (object_idx == TG_U32_MAX) return TG_U32_MAX;
    p_raytracer->p_object_lut_idx[object_idx] = lut_idx;

    const tg_voxel_object* p_object = &p_scene->p_objects[object_idx];
    const v3u dims = p_object->n_cluster_pointers_per_dim;
    const u32 n = dims.x * dims.y * dims.z;
    const u32 first = p_object->first_cluster_pointer;
    struct tgb_device* d = p_raytracer->p_device;

    /* failures are tracked locally: the library's error flag is sticky and may still hold an earlier, unrelated error */
    b32 ok = tgb__upload_object_record(p_raytracer, object_idx);
    ok = ok && tgbd_upload(d, TGB_BUF_CLUSTER_POINTERS, (u64)first * 4, &p_scene->p_cluster_pointers[first], (u64)n * 4);
    /* generated on the device, after the pointer upload (same stream): seeded random bits, or the reference's terrain */
    if (ok && !p_solid_bits)
    {
        if (synthetic_k) ok = tgbd_synthetic_fill(d, synthetic_seed, synthetic_k, n, first);
        else             ok = tgbd_procedural_fill(d, object_idx, dims.x, dims.y, dims.z, first);
    }

    if (ok && tgb__contiguous_run(p_scene, p_object, n))
    {
        const u32 idx0 = p_scene->p_cluster_pointers[first];
        u32* p_mirror = &p_scene->p_voxel_cluster_data[(size_t)idx0 * TG_CLUSTER_MASK_WORDS];
        ok = tgbd_upload(d, TGB_BUF_C2O, (u64)idx0 * 4, &p_scene->p_cluster_idx_to_object_idx[idx0], (u64)n * 4);
        if (p_solid_bits)
        {
            memcpy(p_mirror, p_solid_bits, (size_t)n * 64);
            ok = ok && tgbd_upload(d, TGB_BUF_MASKS, (u64)idx0 * 64, p_solid_bits, (u64)n * 64);
        }
        else ok = ok && tgbd_download(d, TGB_BUF_MASKS, (u64)idx0 * 64, p_mirror, (u64)n * 64);
        if (p_lut_indices) ok = ok && tgbd_upload(d, TGB_BUF_LUT_IDX, (u64)idx0 * 512, p_lut_indices, (u64)n * 512);
    }
    else
    {
        for (u32 rel = 0; ok && rel < n; rel++)
        {
            const u32 idx = p_scene->p_cluster_pointers[first + rel];
            u32* p_mirror = &p_scene->p_voxel_cluster_data[(size_t)idx * TG_CLUSTER_MASK_WORDS];
            ok = tgbd_upload(d, TGB_BUF_C2O, (u64)idx * 4, &object_idx, 4);
            if (p_solid_bits)
            {
                memcpy(p_mirror, &p_solid_bits[(size_t)rel * TG_CLUSTER_MASK_WORDS], 64);
                ok = ok && tgbd_upload(d, TGB_BUF_MASKS, (u64)idx * 64, &p_solid_bits[(size_t)rel * TG_CLUSTER_MASK_WORDS], 64);
            }
            else ok = ok && tgbd_download(d, TGB_BUF_MASKS, (u64)idx * 64, p_mirror, 64);
            if (p_lut_indices) ok = ok && tgbd_upload(d, TGB_BUF_LUT_IDX, (u64)idx * 512, &p_lut_indices[(size_t)rel * 512], 512);
        }
    }
    if (ok && !p_lut_indices) ok = tgbd_fill_default_lut_idx(d, first, n, dims.x);
    p_raytracer->svo_dirty = 2;
    if (!ok)
    {
        /* roll the bookkeeping back (the object was the last one allocated: nothing shifts) and blank its device record, so that a
         * failed create leaves neither a leaked slot on the host nor a half-initialised object on the device */
        tgb__scene_undo_alloc(p_scene, object_idx);
        tgb__upload_object_record(p_raytracer, object_idx);
        return TG_U32_MAX;
    }
    return object_idx;
}

u32 tg_raytracer_create_object_from_data(tg_raytracer* p_raytracer, v3 center, v3u extent, f32 angle_in_radians, v3 axis, u32 lut_idx,
                                         const u32* p_solid_bits, const u8* p_lut_indices)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_create_object_from_data")) return TG_U32_MAX;
    TGB_REQUIRE(p_solid_bits != NULL, TG_U32_MAX, "create_object_from_data: NULL solid bits");
    TGB_REQUIRE(lut_idx < p_raytracer->n_color_luts, TG_U32_MAX, "create_object_from_data: LUT index %u out of range (%u LUTs)", lut_idx, p_raytracer->n_color_luts);
    return tgb__create_object(p_raytracer, center, extent, angle_in_radians, axis, lut_idx, p_solid_bits, p_lut_indices, 0, 0);
}

/* math/tg_math.c:328-338,809-820: the per-cluster streams of the seeded random fill, on the host (the device twin is k_synthetic_fill) */
void tgb200_synthetic_solid_bits(u32 object_seed, u32 k, u32 n_clusters, u32* p_out)
{
    for (u32 rel = 0; rel < n_clusters; rel++)
    {
        u32 state = tgb_hash_u32(object_seed ^ tgb_hash_u32(rel)) | 1u;
        for (u32 w = 0; w < TG_CLUSTER_MASK_WORDS; w++)
        {
            u32 word = 0xFFFFFFFFu;
            for (u32 j = 0; j < k; j++) word &= tgb_xorshift32(&state);
            p_out[(size_t)rel * TG_CLUSTER_MASK_WORDS + w] = word;
        }
    }
}

u32 tg_raytracer_create_object_synthetic(tg_raytracer* p_raytracer, v3 center, v3u extent, f32 angle_in_radians, v3 axis, u32 lut_idx, u32 object_seed, u32 k)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_create_object_synthetic")) return TG_U32_MAX;
    TGB_REQUIRE(k >= 1 && k <= 32, TG_U32_MAX, "create_object_synthetic: k = %u draws per word out of range [1, 32]", k);
    TGB_REQUIRE(lut_idx < p_raytracer->n_color_luts, TG_U32_MAX, "create_object_synthetic: LUT index %u out of range (%u LUTs)", lut_idx, p_raytracer->n_color_luts);
    return tgb__create_object(p_raytracer, center, extent, angle_in_radians, axis, lut_idx, NULL, NULL, k, object_seed);
}

/* tgvk_raytracer.c:805-992: the unmodified reference entry point -- procedural terrain, material (8 x + vx) % 256, angle from the index */
void tg_raytracer_create_object(tg_raytracer* p_raytracer, v3 center, v3u extent)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_create_object")) return;
    TGB_REQUIRE(extent.x % 8 == 0 && extent.y % 8 == 0 && extent.z % 8 == 0 && extent.x && extent.y && extent.z, TGB_VOID,
                "create_object: extent (%u,%u,%u) must be non-zero multiples of 8", extent.x, extent.y, extent.z);
    TGB_REQUIRE(p_raytracer->scene.n_available_object_indices > 0, TGB_VOID, "create_object: object capacity exhausted");
    /* tgvk_raytracer.c:816,829-832: the index the pop will yield decides the angle */
    const u32 next_object_idx = p_raytracer->scene.p_available_object_indices[p_raytracer->scene.n_available_object_indices - 1];
    f32 angle = tgb_deg2rad((f32)(next_object_idx * 7));
    if (next_object_idx == 0) angle = tgb_deg2rad(15.0f);
    const v3 axis = { 0.0f, 1.0f, 0.0f };
    tgb__create_object(p_raytracer, center, extent, angle, axis, 0, NULL, NULL, 0, 0);
}

void tg_raytracer_destroy_object(tg_raytracer* p_raytracer, u32 object_idx)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_destroy_object")) return;
    tg_scene* p_scene = &p_raytracer->scene;
    TGB_REQUIRE(object_idx < p_scene->object_capacity && tg_object_is_initialized(p_scene, object_idx), TGB_VOID, "destroy_object: object %u is not initialised", object_idx);
    u32 first_shifted = 0, n_shifted = 0;
    const u32 old_n_pointers = p_scene->n_cluster_pointers;
    tgb200_scene_free_object(p_scene, object_idx, &first_shifted, &n_shifted);
    struct tgb_device* d = p_raytracer->p_device;
    /* device mirror: shifted pointer range, every object record (first pointers moved), the freed record */
    if (n_shifted) tgbd_upload(d, TGB_BUF_CLUSTER_POINTERS, (u64)first_shifted * 4, &p_scene->p_cluster_pointers[first_shifted], (u64)n_shifted * 4);
    (void)old_n_pointers;
    for (u32 i = 0; i < p_scene->object_capacity; i++)
    {
        if (i == object_idx || (tg_object_is_initialized(p_scene, i) && p_scene->p_objects[i].first_cluster_pointer >= first_shifted && n_shifted))
        {
            tgb__upload_object_record(p_raytracer, i);
        }
    }
    p_raytracer->svo_dirty = 2;
}

void tg_raytracer_set_object_transform(tg_raytracer* p_raytracer, u32 object_idx, v3 translation, f32 angle_in_radians, v3 axis)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_set_object_transform")) return;
    tg_scene* p_scene = &p_raytracer->scene;
    TGB_REQUIRE(object_idx < p_scene->object_capacity && tg_object_is_initialized(p_scene, object_idx), TGB_VOID, "set_object_transform: object %u is not initialised", object_idx);
    TGB_REQUIRE(tgb__axis_is_unit(axis), TGB_VOID, "set_object_transform: rotation axis (%g,%g,%g) is not a unit vector", (double)axis.x, (double)axis.y, (double)axis.z);
    tg_voxel_object* p_object = &p_scene->p_objects[object_idx];
    p_object->translation = translation;
    p_object->angle_in_radians = angle_in_radians;
    p_object->axis = axis;
    tgb__upload_object_record(p_raytracer, object_idx);
    if (p_raytracer->n_moved_objects < p_scene->object_capacity) p_raytracer->p_moved_objects[p_raytracer->n_moved_objects++] = object_idx;
    else p_raytracer->svo_dirty = 2; /* more edits than objects: rebuild */
    if (p_raytracer->svo_dirty < 1) p_raytracer->svo_dirty = 1;
}

void tg_raytracer_color_lut_set_ex(tg_raytracer* p_raytracer, u32 lut_idx, u8 index, f32 r, f32 g, f32 b)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_color_lut_set")) return;
    TGB_REQUIRE(lut_idx < p_raytracer->n_color_luts, TGB_VOID, "color_lut_set: LUT index %u out of range", lut_idx);
    TGB_REQUIRE(r <= 1.0f && g <= 1.0f && b <= 1.0f, TGB_VOID, "color_lut_set: channels must be <= 1"); /* tgvk_raytracer.c:1127-1129 */
    const u32 packed_color = tgb200_pack_color(r, g, b);
    tgbd_upload(p_raytracer->p_device, TGB_BUF_COLOR_LUT, ((u64)lut_idx * 256 + index) * 4, &packed_color, 4);
}

void tg_raytracer_color_lut_set(tg_raytracer* p_raytracer, u8 index, f32 r, f32 g, f32 b)
{
    tg_raytracer_color_lut_set_ex(p_raytracer, 0, index, r, g, b); /* color_lut_idx = 0, tgvk_raytracer.c:1124 */
}

void tg_raytracer_clear(tg_raytracer* p_raytracer)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_clear")) return;
    tgbd_clear(p_raytracer->p_device);
}

void tgb200_render_visibility(tg_raytracer* p_raytracer)
{
    if (!tgb__alive(p_raytracer, "tgb200_render_visibility")) return;
    TGB_REQUIRE(p_raytracer->scene.n_objects > 0, TGB_VOID, "render: the scene has no objects (tgvk_raytracer.c:1147)");
    tg_camera_rays cam;
    tgb200_camera_rays(p_raytracer->p_camera, &cam); /* re-read every frame, tgvk_raytracer.c:1153 */
    /* sharded: peer memory is mapped on the first frame (collective; a no-op afterwards and when the NCCL exchange was chosen), so that
     * K1 can already resolve its materials and flag its tiles for the merge */
    if (tgbd_n_ranks(p_raytracer->p_device) > 1) tgbd_p2p_prepare(p_raytracer->p_device);
    if (p_raytracer->debug_visualization == TG_DEBUG_SHOW_BLOCKS)
    {
        /* tgvk_raytracer.c:1226-1272: while the BLOCKS view is selected the SVO primary-ray pass runs INSTEAD of the cluster pass */
        tgb200_svo_update(p_raytracer, TG_FALSE);
        tgbd_render_visibility_svo(p_raytracer->p_device, &cam);
        return;
    }
    tgbd_render_visibility(p_raytracer->p_device, &cam, p_raytracer->scene.object_capacity);
}

void tgb200_svo_update(tg_raytracer* p_raytracer, b32 force_full)
{
    if (!tgb__alive(p_raytracer, "tgb200_svo_update")) return;
    if (!force_full && !p_raytracer->svo_dirty) return;
    /* tgvk_raytracer.c:1193-1195: fixed +-512 box */
    const v3 extent_min = { -512.0f, -512.0f, -512.0f };
    const v3 extent_max = {  512.0f,  512.0f,  512.0f };
    b32 ok;
    if (!force_full && p_raytracer->svo_dirty == 1)
    {
        /* only transforms changed: re-sample the leaves the moved objects touch(ed), copy the rest */
        ok = tgbd_svo_update_objects(p_raytracer->p_device, extent_min, extent_max, p_raytracer->scene.n_cluster_pointers, p_raytracer->scene.object_capacity,
                                     p_raytracer->n_moved_objects, p_raytracer->p_moved_objects);
    }
    else
    {
        ok = tgbd_svo_build(p_raytracer->p_device, extent_min, extent_max, p_raytracer->scene.n_cluster_pointers, p_raytracer->scene.object_capacity);
    }
    if (ok)
    {
        p_raytracer->svo_dirty = 0;
        p_raytracer->n_moved_objects = 0;
    }
}

u32 tgb200_svo_leaves_resampled(tg_raytracer* p_raytracer)
{
    if (!tgb__alive(p_raytracer, "tgb200_svo_leaves_resampled")) return 0;
    return tgbd_svo_leaves_resampled(p_raytracer->p_device);
}

void tgb200_render_shading(tg_raytracer* p_raytracer)
{
    if (!tgb__alive(p_raytracer, "tgb200_render_shading")) return;
    tg_camera_rays cam;
    tgb200_camera_rays(p_raytracer->p_camera, &cam);
    if (tgbd_n_ranks(p_raytracer->p_device) > 1)
    {
        /* multi-GPU: this rank shades its screen tile, from the all-reduced buffer (tgb200_merge_visibility was called) or, if not,
         * from a tile merged over peer memory (collective set-up on first use) */
        tgbd_p2p_prepare(p_raytracer->p_device);
        tgbd_render_shading_sharded(p_raytracer->p_device, &cam, p_raytracer->scene.n_cluster_pointers, p_raytracer->gi_enabled, p_raytracer->frame_seed,
                                    p_raytracer->debug_visualization);
        return;
    }
    tgbd_render_shading(p_raytracer->p_device, &cam, p_raytracer->scene.n_cluster_pointers, p_raytracer->gi_enabled, p_raytracer->frame_seed,
                        p_raytracer->debug_visualization, 0, p_raytracer->height);
}

void tgb200_render_shading_rows(tg_raytracer* p_raytracer, u32 first_row, u32 one_past_last_row)
{
    if (!tgb__alive(p_raytracer, "tgb200_render_shading_rows")) return;
    TGB_REQUIRE(tgbd_n_ranks(p_raytracer->p_device) <= 1, TGB_VOID, "render_shading_rows: a sharded raytracer shades its own tile (tgb200_render_shading)");
    TGB_REQUIRE(first_row < one_past_last_row && one_past_last_row <= p_raytracer->height, TGB_VOID, "render_shading_rows: rows [%u, %u) outside the %u-row frame",
                first_row, one_past_last_row, p_raytracer->height);
    tg_camera_rays cam;
    tgb200_camera_rays(p_raytracer->p_camera, &cam);
    tgbd_render_shading(p_raytracer->p_device, &cam, p_raytracer->scene.n_cluster_pointers, p_raytracer->gi_enabled, p_raytracer->frame_seed,
                        p_raytracer->debug_visualization, first_row, one_past_last_row);
}

void tg_raytracer_render(tg_raytracer* p_raytracer)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_render")) return;
    TGB_REQUIRE(p_raytracer->scene.n_objects > 0, TGB_VOID, "render: the scene has no objects (tgvk_raytracer.c:1147)");
    /* tgvk_raytracer.c:1187-1217 builds the SVO on the first frame; here whenever it is stale and GI needs it -- queued AFTER K1 so that,
     * on one GPU, the build (its own stream) runs next to K1 instead of in front of it */
    tgb200_render_visibility(p_raytracer);
    if (p_raytracer->gi_enabled) tgb200_svo_update(p_raytracer, TG_FALSE);
    /* multi-GPU: the shading stage merges this rank's tile straight from the peers' buffers (tgb_peer.cu) when peer memory can be
     * mapped; otherwise ncclAllReduce(u64, min) over the whole frame first */
    if (tgbd_n_ranks(p_raytracer->p_device) > 1 && !tgbd_p2p_prepare(p_raytracer->p_device)) tgb200_merge_visibility(p_raytracer);
    tgb200_render_shading(p_raytracer);
}

void tgb200_synchronize(tg_raytracer* p_raytracer)
{
    if (!tgb__alive(p_raytracer, "tgb200_synchronize")) return;
    tgbd_synchronize(p_raytracer->p_device);
}

/*
 * The frame buffers keep rows in VIRTUAL order (tgb_rows.h: 16-row bands dealt out to the ranks; the identity on one GPU).
 * Callers see physical rows: these two move rows [first_row, one_past_last_row) between a device frame buffer and caller memory.
 */
static b32 tgb__download_rows(tg_raytracer* p_raytracer, u32 buffer, u32 bytes_per_pixel, u32 first_row, u32 one_past_last_row, void* p_out)
{
    struct tgb_device* d = p_raytracer->p_device;
    const u32 n_ranks = tgbd_n_ranks(d), tile_rows = tgbd_tile_rows(d);
    const u64 row_bytes = (u64)p_raytracer->width * bytes_per_pixel;
    if (first_row >= one_past_last_row) return TG_TRUE;
    if (n_ranks == 1) return tgbd_download(d, buffer, (u64)first_row * row_bytes, p_out, (u64)(one_past_last_row - first_row) * row_bytes);
    u8* p_tmp = (u8*)malloc((size_t)(tgbd_padded_pixels(d) * bytes_per_pixel));
    if (!p_tmp) { tgb_set_error("read-back: out of host memory"); return TG_FALSE; }
    const b32 ok = tgbd_download(d, buffer, 0, p_tmp, tgbd_padded_pixels(d) * bytes_per_pixel);
    for (u32 row = first_row; ok && row < one_past_last_row; row++)
        memcpy((u8*)p_out + (u64)(row - first_row) * row_bytes, p_tmp + (u64)tgb_row_to_virtual(row, n_ranks, tile_rows) * row_bytes, (size_t)row_bytes);
    free(p_tmp);
    return ok;
}

static b32 tgb__upload_rows(tg_raytracer* p_raytracer, u32 buffer, u32 bytes_per_pixel, const void* p_in)
{
    struct tgb_device* d = p_raytracer->p_device;
    const u32 n_ranks = tgbd_n_ranks(d), tile_rows = tgbd_tile_rows(d);
    const u64 row_bytes = (u64)p_raytracer->width * bytes_per_pixel;
    if (n_ranks == 1) return tgbd_upload(d, buffer, 0, p_in, (u64)p_raytracer->height * row_bytes);
    b32 ok = TG_TRUE;
    for (u32 row = 0; ok && row < p_raytracer->height; row++)
        ok = tgbd_upload(d, buffer, (u64)tgb_row_to_virtual(row, n_ranks, tile_rows) * row_bytes, (const u8*)p_in + (u64)row * row_bytes, row_bytes);
    return ok;
}

b32 tg_raytracer_get_hovered_voxel(tg_raytracer* p_raytracer, u32 screen_x, u32 screen_y, f32* p_depth, u32* p_cluster_idx, u32* p_voxel_idx)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_get_hovered_voxel")) return TG_FALSE;
    TGB_REQUIRE(screen_x < p_raytracer->width && screen_y < p_raytracer->height, TG_FALSE, "get_hovered_voxel: pixel (%u,%u) outside %ux%u", screen_x, screen_y, p_raytracer->width, p_raytracer->height);
    u64 packed_data = TG_VIS_CLEAR;
    const u64 pixel_idx = (u64)p_raytracer->width * tgb_row_to_virtual(screen_y, tgbd_n_ranks(p_raytracer->p_device), tgbd_tile_rows(p_raytracer->p_device)) + screen_x;
    if (!tgbd_download(p_raytracer->p_device, TGB_BUF_VISIBILITY_MERGED, pixel_idx * 8, &packed_data, 8)) return TG_FALSE;
    /* tgvk_raytracer.c:1639-1654 */
    *p_depth = (f32)(packed_data >> 40) / 16777215.0f;
    if (*p_depth < 1.0f)
    {
        *p_cluster_idx = (u32)(packed_data >> 9) & 2147483647u;
        *p_voxel_idx = (u32)(packed_data) & 511u;
        return TG_TRUE;
    }
    *p_cluster_idx = TG_U32_MAX;
    *p_voxel_idx = TG_U32_MAX;
    return TG_FALSE;
}

void tg_raytracer_read_visibility(tg_raytracer* p_raytracer, u64* p_out)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_read_visibility")) return;
    tgb__download_rows(p_raytracer, TGB_BUF_VISIBILITY_MERGED, 8, 0, p_raytracer->height, p_out);
}

void tg_raytracer_write_visibility(tg_raytracer* p_raytracer, const u64* p_in)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_write_visibility")) return;
    tgb__upload_rows(p_raytracer, TGB_BUF_VISIBILITY, 8, p_in);
}

void tg_raytracer_read_radiance(tg_raytracer* p_raytracer, f32* p_out)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_read_radiance")) return;
    tgb__download_rows(p_raytracer, TGB_BUF_RADIANCE, 16, 0, p_raytracer->height, p_out);
}

void tg_raytracer_read_radiance_rows(tg_raytracer* p_raytracer, u32 first_row, u32 one_past_last_row, f32* p_out)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_read_radiance_rows")) return;
    TGB_REQUIRE(first_row <= one_past_last_row && one_past_last_row <= p_raytracer->height, TGB_VOID, "read_radiance_rows: rows [%u, %u) outside the frame", first_row, one_past_last_row);
    if (first_row == one_past_last_row) return;
    tgb__download_rows(p_raytracer, TGB_BUF_RADIANCE, 16, first_row, one_past_last_row, p_out);
}

void tgb200_get_timings(tg_raytracer* p_raytracer, tgb200_timings* p_out)
{
    memset(p_out, 0, sizeof(*p_out));
    if (!tgb__alive(p_raytracer, "tgb200_get_timings")) return;
    tgbd_get_timings(p_raytracer->p_device, p_out);
}

void tgb200_reset_launch_counter(tg_raytracer* p_raytracer)
{
    if (!tgb__alive(p_raytracer, "tgb200_reset_launch_counter")) return;
    tgbd_reset_launch_counter(p_raytracer->p_device);
}

void* tgb200_device_visibility(tg_raytracer* p_raytracer) { return tgb__alive(p_raytracer, "tgb200_device_visibility") ? tgbd_buffer(p_raytracer->p_device, TGB_BUF_VISIBILITY) : NULL; }
void* tgb200_device_radiance(tg_raytracer* p_raytracer) { return tgb__alive(p_raytracer, "tgb200_device_radiance") ? tgbd_buffer(p_raytracer->p_device, TGB_BUF_RADIANCE) : NULL; }
void* tgb200_stream(tg_raytracer* p_raytracer) { return tgb__alive(p_raytracer, "tgb200_stream") ? tgbd_stream(p_raytracer->p_device) : NULL; }

/* ---- SVO entry points --------------------------------------------------------------------------- */

/* the scene of a live raytracer sits inside it: recover the owner (scene is the 2nd member) */
static tg_raytracer* tgb__owner_of(const tg_scene* p_scene)
{
    return (tg_raytracer*)((char*)p_scene - offsetof(tg_raytracer, scene));
}

void tgb200_svo_download(tg_raytracer* p_raytracer, tg_svo* p_svo)
{
    memset(p_svo, 0, sizeof(*p_svo));
    if (!tgb__alive(p_raytracer, "tgb200_svo_download")) return;
    u32 n_nodes = 0, n_leaves = 0, n_words = 0;
    v3 bmin, bmax;
    if (!tgbd_svo_counts(p_raytracer->p_device, &n_nodes, &n_leaves, &n_words, &bmin, &bmax)) return;
    /* tg_sparse_voxel_octree.c:476-493 */
    p_svo->min = bmin;
    p_svo->max = bmax;
    p_svo->voxel_buffer_capacity_in_u32 = 1u << 21;
    p_svo->leaf_node_data_buffer_capacity = 1u << 13;
    p_svo->node_buffer_capacity = 1u << 14;
    if (n_words > p_svo->voxel_buffer_capacity_in_u32) p_svo->voxel_buffer_capacity_in_u32 = n_words;
    if (n_leaves > p_svo->leaf_node_data_buffer_capacity) p_svo->leaf_node_data_buffer_capacity = n_leaves;
    if (n_nodes > p_svo->node_buffer_capacity) p_svo->node_buffer_capacity = n_nodes;
    p_svo->p_voxels_buffer = (u32*)calloc(p_svo->voxel_buffer_capacity_in_u32, sizeof(u32));
    p_svo->p_leaf_node_data_buffer = (tg_svo_leaf_node_data*)calloc(p_svo->leaf_node_data_buffer_capacity, sizeof(tg_svo_leaf_node_data));
    p_svo->p_node_buffer = (tg_svo_node*)calloc(p_svo->node_buffer_capacity, sizeof(tg_svo_node));
    p_svo->voxel_buffer_count_in_u32 = n_words;
    p_svo->leaf_node_data_buffer_count = n_leaves;
    p_svo->node_buffer_count = n_nodes;
    if (n_nodes)  tgbd_download(p_raytracer->p_device, TGB_BUF_SVO_NODES, 0, p_svo->p_node_buffer, (u64)n_nodes * 4);
    if (n_leaves) tgbd_download(p_raytracer->p_device, TGB_BUF_SVO_LEAF_DATA, 0, p_svo->p_leaf_node_data_buffer, (u64)n_leaves * 260);
    if (n_words)  tgbd_download(p_raytracer->p_device, TGB_BUF_SVO_VOXELS, 0, p_svo->p_voxels_buffer, (u64)n_words * 4);
}

void tgb200_svo_upload(tg_raytracer* p_raytracer, const tg_svo* p_svo)
{
    if (!tgb__alive(p_raytracer, "tgb200_svo_upload")) return;
    tgbd_svo_set(p_raytracer->p_device, p_svo->min, p_svo->max, p_svo->node_buffer_count, p_svo->p_node_buffer,
                 p_svo->leaf_node_data_buffer_count, p_svo->p_leaf_node_data_buffer, p_svo->voxel_buffer_count_in_u32, p_svo->p_voxels_buffer);
    p_raytracer->svo_dirty = 0;
}

void tg_svo_create(v3 extent_min, v3 extent_max, const tg_scene* p_scene, tg_svo* p_svo)
{
    TGB_REQUIRE(p_scene && p_svo, TGB_VOID, "tg_svo_create: NULL argument");
    /* tg_sparse_voxel_octree.c:468-472 */
    TGB_REQUIRE(extent_max.x - extent_min.x == (f32)TG_SVO_SIDE_LENGTH && extent_max.y - extent_min.y == (f32)TG_SVO_SIDE_LENGTH && extent_max.z - extent_min.z == (f32)TG_SVO_SIDE_LENGTH,
                TGB_VOID, "tg_svo_create: the extent must be %d^3", TG_SVO_SIDE_LENGTH);
    TGB_REQUIRE(p_scene->n_cluster_pointers > 0, TGB_VOID, "tg_svo_create: empty scene");
    tg_raytracer* p_raytracer = tgb__owner_of(p_scene);
    if (!tgb__alive(p_raytracer, "tg_svo_create")) return;
    if (!tgbd_svo_build(p_raytracer->p_device, extent_min, extent_max, p_scene->n_cluster_pointers, p_scene->object_capacity)) return;
    p_raytracer->svo_dirty = 0;
    p_raytracer->n_moved_objects = 0;
    tgb200_svo_download(p_raytracer, p_svo);
}

void tg_svo_destroy(tg_svo* p_svo)
{
    if (!p_svo) return;
    free(p_svo->p_voxels_buffer);
    free(p_svo->p_leaf_node_data_buffer);
    free(p_svo->p_node_buffer);
    memset(p_svo, 0, sizeof(*p_svo));
}

/* ---- multi-GPU ---------------------------------------------------------------------------------- */



void tgb200_set_shard(tg_raytracer* p_raytracer, u32 rank, u32 n_ranks, u32 global_pointer_base)
{
    if (!tgb__alive(p_raytracer, "tgb200_set_shard")) return;
    TGB_REQUIRE(rank < n_ranks, TGB_VOID, "tgb200_set_shard: rank %u >= n_ranks %u", rank, n_ranks);
    TGB_REQUIRE((u64)global_pointer_base + p_raytracer->scene.cluster_pointer_capacity <= (u64)TGB_MAX_N_CLUSTERS_31B, TGB_VOID, "tgb200_set_shard: global pointers exceed 31 bits");
    tgbd_set_shard(p_raytracer->p_device, global_pointer_base);
}

void tgb200_comm_unique_id(u8* p_out_128) { tgbn_unique_id(p_out_128); }

/* the communicator belongs to the raytracer that joined it (one process per GPU, one sharded raytracer per process) */
void tgb200_comm_init(tg_raytracer* p_raytracer, const u8* p_unique_id_128, u32 rank, u32 n_ranks)
{
    if (!tgb__alive(p_raytracer, "tgb200_comm_init")) return;
    TGB_REQUIRE(rank < n_ranks, TGB_VOID, "tgb200_comm_init: rank %u >= n_ranks %u", rank, n_ranks);
    tgb200_comm_destroy(p_raytracer);
    void* p_comm = tgbn_init(p_unique_id_128, rank, n_ranks);
    if (p_comm) tgbd_set_comm(p_raytracer->p_device, p_comm, rank, n_ranks);
}

void tgb200_comm_destroy(tg_raytracer* p_raytracer)
{
    if (!p_raytracer || !p_raytracer->p_device) return;
    void* p_comm = tgbd_comm(p_raytracer->p_device);
    if (!p_comm) return;
    tgbd_set_comm(p_raytracer->p_device, NULL, 0, 1);
    tgbn_destroy(p_comm);
}

void tgb200_set_frame_sink(tg_raytracer* p_raytracer, f32* p_host, u32 n_bands)
{
    if (!tgb__alive(p_raytracer, "tgb200_set_frame_sink")) return;
    tgbd_set_frame_sink(p_raytracer->p_device, p_host, n_bands, TGB200_SINK_RGBA32F);
}

void tgb200_set_frame_sink_ex(tg_raytracer* p_raytracer, void* p_host, u32 n_bands, tgb200_sink_format format)
{
    if (!tgb__alive(p_raytracer, "tgb200_set_frame_sink_ex")) return;
    TGB_REQUIRE(format == TGB200_SINK_RGBA32F || format == TGB200_SINK_BGRA8, TGB_VOID, "set_frame_sink_ex: unknown format %u", (u32)format);
    tgbd_set_frame_sink(p_raytracer->p_device, p_host, n_bands, (u32)format);
}

void tg_raytracer_read_present(tg_raytracer* p_raytracer, u32* p_out)
{
    if (!tgb__alive(p_raytracer, "tg_raytracer_read_present")) return;
    struct tgb_device* d = p_raytracer->p_device;
    const u32 n_ranks = tgbd_n_ranks(d), tile_rows = tgbd_tile_rows(d);
    const u64 n_px = (u64)p_raytracer->width * p_raytracer->height;
    u32* p_tmp = (u32*)malloc((size_t)(tgbd_padded_pixels(d) * 4u)); /* the device frame is padded to whole 16-row bands, rows in virtual order */
    if (!p_tmp) { tgb_set_error("read_present: out of host memory"); return; }
    if (tgbd_read_present(d, p_tmp))
    {
        if (n_ranks == 1) memcpy(p_out, p_tmp, (size_t)n_px * 4u);
        else for (u32 row = 0; row < p_raytracer->height; row++)
            memcpy(p_out + (u64)row * p_raytracer->width, p_tmp + (u64)tgb_row_to_virtual(row, n_ranks, tile_rows) * p_raytracer->width, (size_t)p_raytracer->width * 4u);
    }
    free(p_tmp);
}

u64 tgb200_frame_ticket(tg_raytracer* p_raytracer)
{
    if (!tgb__alive(p_raytracer, "tgb200_frame_ticket")) return 0;
    return tgbd_frames_sunk(p_raytracer->p_device);
}

void tgb200_wait_frame(tg_raytracer* p_raytracer, u64 ticket)
{
    if (!tgb__alive(p_raytracer, "tgb200_wait_frame")) return;
    tgbd_wait_frame(p_raytracer->p_device, ticket);
}

void tgb200_set_gi_traversal(tg_raytracer* p_raytracer, u32 kind)
{
    if (!tgb__alive(p_raytracer, "tgb200_set_gi_traversal")) return;
    tgbd_set_gi_traversal(p_raytracer->p_device, kind);
}

void tgb200_mark_svo_dirty(tg_raytracer* p_raytracer)
{
    if (!tgb__alive(p_raytracer, "tgb200_mark_svo_dirty")) return;
    p_raytracer->svo_dirty = 2;
}

void tgb200_gather_radiance(tg_raytracer* p_raytracer)
{
    if (!tgb__alive(p_raytracer, "tgb200_gather_radiance")) return;
    tgbd_gather_radiance(p_raytracer->p_device);
}

void tgb200_tile_rows(tg_raytracer* p_raytracer, u32* p_first_row, u32* p_one_past_last_row)
{
    *p_first_row = 0; *p_one_past_last_row = 0;
    if (!tgb__alive(p_raytracer, "tgb200_tile_rows")) return;
    if (tgbd_n_ranks(p_raytracer->p_device) == 1) { *p_one_past_last_row = p_raytracer->height; return; }
    const u32 rows = tgbd_tile_rows(p_raytracer->p_device);
    *p_first_row = tgbd_rank(p_raytracer->p_device) * rows;
    *p_one_past_last_row = *p_first_row + rows;
}

u32 tgb200_tile_physical_row(tg_raytracer* p_raytracer, u32 tile_row)
{
    if (!tgb__alive(p_raytracer, "tgb200_tile_physical_row")) return TG_U32_MAX;
    const u32 n_ranks = tgbd_n_ranks(p_raytracer->p_device), rows = tgbd_tile_rows(p_raytracer->p_device);
    if (tile_row >= rows) return TG_U32_MAX;
    const u32 physical = tgb_row_to_physical(tgbd_rank(p_raytracer->p_device) * rows + tile_row, n_ranks, rows);
    return physical < p_raytracer->height ? physical : TG_U32_MAX;
}

void tgb200_set_merge_kind(tg_raytracer* p_raytracer, u32 kind)
{
    if (!tgb__alive(p_raytracer, "tgb200_set_merge_kind")) return;
    tgbd_set_merge_kind(p_raytracer->p_device, kind);
}

void tgb200_merge_visibility(tg_raytracer* p_raytracer)
{
    if (!tgb__alive(p_raytracer, "tgb200_merge_visibility")) return;
    void* p_comm = tgbd_comm(p_raytracer->p_device);
    TGB_REQUIRE(p_comm != NULL, TGB_VOID, "tgb200_merge_visibility: no communicator (call tgb200_comm_init)");
    tgbd_merge_begin(p_raytracer->p_device);
    if (tgbn_allreduce_min_u64(p_comm, tgbd_buffer(p_raytracer->p_device, TGB_BUF_VISIBILITY), tgbd_padded_pixels(p_raytracer->p_device), tgbd_stream(p_raytracer->p_device)))
        tgbd_note_merged(p_raytracer->p_device);
    tgbd_merge_end(p_raytracer->p_device);
}

/* tgb_host_extra.c -- host-only entry points of the boundary: the single-ray SVO query, verification hooks, the procedural fill. */
#include "tgb_internal.h"
#include "tgb_math.h"
#include "tgb_hoist.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/*
 * The presented frame on disc. Same container as the reference's writer (graphics/tg_image_io.c:438-520): 14-byte file header,
 * 124-byte BITMAPV5HEADER with BI_BITFIELDS channel masks, 12 unused bytes, pixels from byte 150, negative height = rows top
 * down, 32 bits per pixel. The pixels are the present pass's B8G8R8A8 words, i.e. masks r 0x00FF0000, g 0x0000FF00,
 * b 0x000000FF, a 0xFF000000.
 */
static void tgb__put_u16(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); }
static void tgb__put_u32(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); p[3] = (u8)(v >> 24); }

b32 tgb200_write_bmp_bgra8(const char* p_filename, u32 width, u32 height, const u32* p_pixels)
{
    enum { FILE_HEADER = 14, V5_HEADER = 124, PIXEL_OFFSET = 150 };
    const u64 n_pixel_bytes = (u64)width * height * 4u;
    if (!p_filename || !p_pixels || width == 0 || height == 0 || PIXEL_OFFSET + n_pixel_bytes > 0xFFFFFFFFull)
    {
        tgb_set_error("write_bmp: bad arguments (%ux%u)", width, height);
        return TG_FALSE;
    }
    u8 header[PIXEL_OFFSET];
    memset(header, 0, sizeof(header));
    header[0] = 'B'; header[1] = 'M';
    tgb__put_u32(header + 2, (u32)(PIXEL_OFFSET + n_pixel_bytes));
    tgb__put_u32(header + 10, PIXEL_OFFSET);
    u8* p_info = header + FILE_HEADER;
    tgb__put_u32(p_info + 0, V5_HEADER);
    tgb__put_u32(p_info + 4, width);
    tgb__put_u32(p_info + 8, (u32)(-(i32)height));  /* top-down */
    tgb__put_u16(p_info + 12, 1);                   /* planes */
    tgb__put_u16(p_info + 14, 32);                  /* bits per pixel */
    tgb__put_u32(p_info + 16, 3);                   /* BI_BITFIELDS */
    tgb__put_u32(p_info + 20, (u32)n_pixel_bytes);
    tgb__put_u32(p_info + 40, 0x00FF0000u);         /* red mask */
    tgb__put_u32(p_info + 44, 0x0000FF00u);         /* green */
    tgb__put_u32(p_info + 48, 0x000000FFu);         /* blue */
    tgb__put_u32(p_info + 52, 0xFF000000u);         /* alpha */
    tgb__put_u32(p_info + 56, 0x57696E20u);         /* 'Win ': LCS_WINDOWS_COLOR_SPACE */
    FILE* p_file = fopen(p_filename, "wb");
    if (!p_file) { tgb_set_error("write_bmp: cannot open %s", p_filename); return TG_FALSE; }
    const b32 ok = fwrite(header, 1, sizeof(header), p_file) == sizeof(header) && fwrite(p_pixels, 1, (size_t)n_pixel_bytes, p_file) == (size_t)n_pixel_bytes;
    if (fclose(p_file) != 0 || !ok) { tgb_set_error("write_bmp: short write to %s", p_filename); return TG_FALSE; }
    return TG_TRUE;
}

b32 tgb200_save_frame_bmp(tg_raytracer* p_raytracer, const char* p_filename)
{
    if (!p_raytracer || !p_raytracer->p_device) { tgb_set_error("tgb200_save_frame_bmp: raytracer is not alive"); return TG_FALSE; }
    const u64 n = (u64)p_raytracer->width * p_raytracer->height;
    u32* p_pixels = (u32*)malloc((size_t)n * 4u);
    if (!p_pixels) { tgb_set_error("tgb200_save_frame_bmp: out of memory"); return TG_FALSE; }
    tgb200_clear_error();
    tg_raytracer_read_present(p_raytracer, p_pixels); /* un-permutes the rows of a sharded frame */
    b32 ok = tgb200_last_error() == NULL;
    if (ok) ok = tgb200_write_bmp_bgra8(p_filename, p_raytracer->width, p_raytracer->height, p_pixels);
    free(p_pixels);
    return ok;
}

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
    if (tgbd_device_count() <= 0) { tgb_set_error("tgb200_procedural_solid_bits: no CUDA device (the fill runs on the GPU; there is no CPU path)"); return; }
    tgbd_procedural_bits_to_host(tgbd_current_device(), object_idx, n_cluster_pointers_per_dim.x, n_cluster_pointers_per_dim.y, n_cluster_pointers_per_dim.z, p_out);
}

/* physics/tg_physics.c:394-406 (C twin of collide.inc: TG_MIN / TG_MAX are C ternaries, math/tg_math.h:17-22) */
static b32 tgb__ray_aabb_c(v3 o, v3 d, v3 bmin, v3 bmax, f32* p_enter, f32* p_exit)
{
    const f32 ax = (d.x == 0.0f) ? TG_F32_MIN : ((bmin.x - o.x) / d.x), bx = (d.x == 0.0f) ? TG_F32_MAX : ((bmax.x - o.x) / d.x);
    const f32 ay = (d.y == 0.0f) ? TG_F32_MIN : ((bmin.y - o.y) / d.y), by = (d.y == 0.0f) ? TG_F32_MAX : ((bmax.y - o.y) / d.y);
    const f32 az = (d.z == 0.0f) ? TG_F32_MIN : ((bmin.z - o.z) / d.z), bz = (d.z == 0.0f) ? TG_F32_MAX : ((bmax.z - o.z) / d.z);
    const f32 nx = ax < bx ? ax : bx, ny = ay < by ? ay : by, nz = az < bz ? az : bz;
    const f32 fx = ax > bx ? ax : bx, fy = ay > by ? ay : by, fz = az > bz ? az : bz;
    const f32 nxy = nx > ny ? nx : ny, fxy = fx < fy ? fx : fy;
    *p_enter = nxy > nz ? nxy : nz;
    *p_exit  = fxy < fz ? fxy : fz;
    return *p_exit > 0.0f && *p_enter <= *p_exit;
}

/* tg_sparse_voxel_octree.c:704-707: distance to the far border of a box */
static f32 tgb__exit_c(v3 bmin, v3 bmax, v3 p, v3 d)
{
    const f32 ax = (d.x == 0.0f) ? TG_F32_MIN : ((bmin.x - p.x) / d.x), bx = (d.x == 0.0f) ? TG_F32_MAX : ((bmax.x - p.x) / d.x);
    const f32 ay = (d.y == 0.0f) ? TG_F32_MIN : ((bmin.y - p.y) / d.y), by = (d.y == 0.0f) ? TG_F32_MAX : ((bmax.y - p.y) / d.y);
    const f32 az = (d.z == 0.0f) ? TG_F32_MIN : ((bmin.z - p.z) / d.z), bz = (d.z == 0.0f) ? TG_F32_MAX : ((bmax.z - p.z) / d.z);
    const f32 fx = ax > bx ? ax : bx, fy = ay > by ? ay : by, fz = az > bz ? az : bz;
    const f32 m = fx < fy ? fx : fy;
    return m < fz ? m : fz;
}

/* util/tg_amanatides_woo.c:3-114 on one leaf block: no lower clamp of the start cell, t_max without an `enter` term */
static b32 tgb__block_dda(v3 hit, v3 d, v3 extent, const u32* p_grid, i32* p_x, i32* p_y, i32* p_z)
{
    const f32 cx = floorf(hit.x) < extent.x - 1.0f ? floorf(hit.x) : extent.x - 1.0f;
    const f32 cy = floorf(hit.y) < extent.y - 1.0f ? floorf(hit.y) : extent.y - 1.0f;
    const f32 cz = floorf(hit.z) < extent.z - 1.0f ? floorf(hit.z) : extent.z - 1.0f;
    i32 x = (i32)cx, y = (i32)cy, z = (i32)cz;
    i32 sx = 0, sy = 0, sz = 0;
    f32 tmx = TG_F32_MAX, tmy = TG_F32_MAX, tmz = TG_F32_MAX, tdx = TG_F32_MAX, tdy = TG_F32_MAX, tdz = TG_F32_MAX;
    if (d.x > 0.0f)      { sx = 1;  tmx = ((f32)(x + 1) - hit.x) / d.x; tdx = 1.0f / d.x; }
    else if (d.x < 0.0f) { sx = -1; tmx = (hit.x - (f32)x) / -d.x;      tdx = 1.0f / -d.x; }
    if (d.y > 0.0f)      { sy = 1;  tmy = ((f32)(y + 1) - hit.y) / d.y; tdy = 1.0f / d.y; }
    else if (d.y < 0.0f) { sy = -1; tmy = (hit.y - (f32)y) / -d.y;      tdy = 1.0f / -d.y; }
    if (d.z > 0.0f)      { sz = 1;  tmz = ((f32)(z + 1) - hit.z) / d.z; tdz = 1.0f / d.z; }
    else if (d.z < 0.0f) { sz = -1; tmz = (hit.z - (f32)z) / -d.z;      tdz = 1.0f / -d.z; }
    const u32 ex = (u32)extent.x, ey = (u32)extent.y;
    if (x < 0 || y < 0 || z < 0) return TG_FALSE; /* the reference would index out of bounds here (AW.c:12-13 has no lower clamp) */
    for (;;)
    {
        const u32 v = ex * ey * (u32)z + ex * (u32)y + (u32)x;
        if (p_grid[v / 32] & (1u << (v % 32))) { *p_x = x; *p_y = y; *p_z = z; return TG_TRUE; }
        if (tmx < tmy)
        {
            if (tmx < tmz) { tmx += tdx; x += sx; if (x < 0 || (f32)x >= extent.x) break; }
            else           { tmz += tdz; z += sz; if (z < 0 || (f32)z >= extent.z) break; }
        }
        else
        {
            if (tmy < tmz) { tmy += tdy; y += sy; if (y < 0 || (f32)y >= extent.y) break; }
            else           { tmz += tdz; z += sz; if (z < 0 || (f32)z >= extent.z) break; }
        }
    }
    return TG_FALSE;
}

/*
 * tg_sparse_voxel_octree.c:558-740: the reference's HOST-side single-ray query over a tg_svo (picking / debugging). It is
 * part of the boundary (tg_sparse_voxel_octree.h:53) and runs on the host in the reference too, so it does here: one ray
 * over host arrays is not device work. This is the C variant (distance accumulates t_advance, :709-712; Amanatides-Woo
 * of util/tg_amanatides_woo.c); the GI kernel follows the GLSL variant.
 */
b32 tg_svo_traverse(const tg_svo* p_svo, v3 ray_origin, v3 ray_direction, f32* p_distance, u32* p_node_idx, u32* p_voxel_idx)
{
    *p_distance = TG_F32_MAX;
    *p_node_idx = TG_U32_MAX;
    *p_voxel_idx = TG_U32_MAX;
    if (!p_svo || !p_svo->p_node_buffer) { tgb_set_error("tg_svo_traverse: no SVO"); return TG_FALSE; }

    const v3 extent = tgb_sub(p_svo->max, p_svo->min);
    const v3 center = tgb_add(p_svo->min, tgb_scale(extent, 0.5f));
    const v3 o = tgb_sub(ray_origin, center);
    const v3 d = ray_direction;

    u32 stack_size, idx_stack[TG_SVO_TRAVERSE_STACK_CAPACITY] = { 0 };
    v3 min_stack[TG_SVO_TRAVERSE_STACK_CAPACITY], max_stack[TG_SVO_TRAVERSE_STACK_CAPACITY];
    f32 enter, exit;
    if (!tgb__ray_aabb_c(o, d, p_svo->min, p_svo->max, &enter, &exit)) return TG_FALSE;
    v3 position = enter > 0.0f ? tgb_add(o, tgb_scale(d, enter)) : o;
    f32 travelled = enter > 0.0f ? enter : 0.0f;

    stack_size = 1;
    min_stack[0] = p_svo->min;
    max_stack[0] = p_svo->max;
    for (u32 iterations = 0; stack_size > 0 && iterations < 4096u; iterations++) /* Q9: cap valid input never reaches */
    {
        const u32 parent_idx = idx_stack[stack_size - 1];
        const v3 parent_min = min_stack[stack_size - 1], parent_max = max_stack[stack_size - 1];
        const tg_svo_inner_node node = p_svo->p_node_buffer[parent_idx].inner;
        const v3 child_extent = tgb_scale(tgb_sub(parent_max, parent_min), 0.5f);
        u32 oct = 0;
        v3 child_min = parent_min;
        v3 child_max = tgb_add(child_min, child_extent);
        if (child_max.x < position.x || (position.x == child_max.x && d.x > 0.0f)) { oct += 1; child_min.x += child_extent.x; child_max.x += child_extent.x; }
        if (child_max.y < position.y || (position.y == child_max.y && d.y > 0.0f)) { oct += 2; child_min.y += child_extent.y; child_max.y += child_extent.y; }
        if (child_max.z < position.z || (position.z == child_max.z && d.z > 0.0f)) { oct += 4; child_min.z += child_extent.z; child_max.z += child_extent.z; }

        b32 advance_to_border = TG_TRUE;
        if (node.valid_mask & (1u << oct))
        {
            u32 rank = 0;
            for (u32 i = 0; i < oct; i++) rank += (node.valid_mask >> i) & 1u;
            const u32 child_idx = parent_idx + node.child_pointer + rank;
            if (node.leaf_mask & (1u << oct))
            {
                const u32 data_pointer = p_svo->p_node_buffer[child_idx].leaf.data_pointer;
                if (p_svo->p_leaf_node_data_buffer[data_pointer].n != 0)
                {
                    const u32 first_voxel_id = data_pointer * TG_SVO_BLOCK_VOXEL_COUNT;
                    i32 vx, vy, vz;
                    if (tgb__block_dda(tgb_sub(position, child_min), d, child_extent, p_svo->p_voxels_buffer + first_voxel_id / 32, &vx, &vy, &vz))
                    {
                        const v3 voxel_min = tgb_add(child_min, tgb_v3((f32)vx, (f32)vy, (f32)vz));
                        const v3 voxel_max = tgb_add(voxel_min, tgb_v3(1.0f, 1.0f, 1.0f));
                        tgb__ray_aabb_c(position, d, voxel_min, voxel_max, &enter, &exit);
                        *p_distance = travelled + enter;
                        *p_node_idx = child_idx;
                        *p_voxel_idx = first_voxel_id + (u32)child_extent.x * (u32)child_extent.y * (u32)vz + (u32)child_extent.x * (u32)vy + (u32)vx;
                        return TG_TRUE;
                    }
                }
            }
            else if (stack_size < TG_SVO_TRAVERSE_STACK_CAPACITY)
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
            const f32 t_advance = tgb__exit_c(child_min, child_max, position, d) + TG_F32_EPSILON;
            position = tgb_add(position, tgb_scale(d, t_advance));
            travelled += t_advance;
            while (stack_size > 0 && !(tgb__exit_c(min_stack[stack_size - 1], max_stack[stack_size - 1], position, d) > TG_F32_EPSILON)) stack_size--;
        }
    }
    return TG_FALSE;
}

/* ---- scene dump / load: reproducible scenes for benches and bug reports ------------------------------- */
/*
 * File = "TGB200SC" | u32 version (1) | u32 n_objects | u32 n_color_luts | n_color_luts * 256 packed colours | per initialised object
 * in ascending object index: { v3u dims; v3 translation; f32 angle; v3 axis; u32 lut_idx; dims.x*dims.y*dims.z * 16 u32 solid
 * masks; the same number * 512 u8 material indices }, clusters in pointer order. Little-endian, no padding. The masks come from
 * the CPU mirror (scene.p_voxel_cluster_data, tgvk_raytracer.h:100), materials and colour LUTs from the device (the reference
 * keeps them in SSBOs only, tgvk_raytracer.c:20,44-47).
 */
#define TGB_SCENE_MAGIC "TGB200SC"

b32 tgb200_scene_save(tg_raytracer* p_raytracer, const char* p_filename)
{
    if (!p_raytracer || !p_raytracer->p_device) { tgb_set_error("tgb200_scene_save: raytracer is not alive"); return TG_FALSE; }
    const tg_scene* p_scene = &p_raytracer->scene;
    FILE* p_file = fopen(p_filename, "wb");
    if (!p_file) { tgb_set_error("tgb200_scene_save: cannot open %s", p_filename); return TG_FALSE; }
    b32 ok = TG_TRUE;
    u32 n_objects = 0;
    for (u32 i = 0; i < p_scene->object_capacity; i++) n_objects += tg_object_is_initialized(p_scene, i) ? 1u : 0u;
    const u32 version = 1, n_luts = p_raytracer->n_color_luts;
    ok = ok && fwrite(TGB_SCENE_MAGIC, 1, 8, p_file) == 8 && fwrite(&version, 4, 1, p_file) == 1 && fwrite(&n_objects, 4, 1, p_file) == 1 && fwrite(&n_luts, 4, 1, p_file) == 1;
    u32* p_lut = (u32*)malloc((size_t)n_luts * 256u * 4u);
    ok = ok && p_lut && tgbd_download(p_raytracer->p_device, TGB_BUF_COLOR_LUT, 0, p_lut, (u64)n_luts * 256u * 4u) && fwrite(p_lut, 4, (size_t)n_luts * 256u, p_file) == (size_t)n_luts * 256u;
    free(p_lut);
    u8* p_materials = NULL;
    size_t materials_capacity = 0;
    for (u32 i = 0; ok && i < p_scene->object_capacity; i++)
    {
        if (!tg_object_is_initialized(p_scene, i)) continue;
        const tg_voxel_object* o = &p_scene->p_objects[i];
        const u32 n = o->n_cluster_pointers_per_dim.x * o->n_cluster_pointers_per_dim.y * o->n_cluster_pointers_per_dim.z;
        const u32 lut_idx = p_raytracer->p_object_lut_idx[i];
        ok = fwrite(&o->n_cluster_pointers_per_dim, sizeof(v3u), 1, p_file) == 1 && fwrite(&o->translation, sizeof(v3), 1, p_file) == 1
          && fwrite(&o->angle_in_radians, 4, 1, p_file) == 1 && fwrite(&o->axis, sizeof(v3), 1, p_file) == 1 && fwrite(&lut_idx, 4, 1, p_file) == 1;
        if ((size_t)n * 512u > materials_capacity)
        {
            free(p_materials);
            materials_capacity = (size_t)n * 512u;
            p_materials = (u8*)malloc(materials_capacity);
            ok = ok && p_materials != NULL;
        }
        for (u32 rel = 0; ok && rel < n; rel++)
        {
            const u32 idx = p_scene->p_cluster_pointers[o->first_cluster_pointer + rel];
            ok = fwrite(&p_scene->p_voxel_cluster_data[(size_t)idx * TG_CLUSTER_MASK_WORDS], 4, TG_CLUSTER_MASK_WORDS, p_file) == TG_CLUSTER_MASK_WORDS;
        }
        /* materials: one download when the object's cluster indices are one ascending run (always, unless objects were destroyed) */
        const u32 idx0 = n ? p_scene->p_cluster_pointers[o->first_cluster_pointer] : 0u;
        b32 run = TG_TRUE;
        for (u32 rel = 1; rel < n; rel++) run = run && p_scene->p_cluster_pointers[o->first_cluster_pointer + rel] == idx0 + rel;
        if (ok && run && n) ok = tgbd_download(p_raytracer->p_device, TGB_BUF_LUT_IDX, (u64)idx0 * 512u, p_materials, (u64)n * 512u);
        for (u32 rel = 0; ok && !run && rel < n; rel++)
            ok = tgbd_download(p_raytracer->p_device, TGB_BUF_LUT_IDX, (u64)p_scene->p_cluster_pointers[o->first_cluster_pointer + rel] * 512u, p_materials + (size_t)rel * 512u, 512u);
        ok = ok && fwrite(p_materials, 512, n, p_file) == n;
    }
    free(p_materials);
    if (fclose(p_file) != 0) ok = TG_FALSE;
    if (!ok && !tgb200_last_error()) tgb_set_error("tgb200_scene_save: short write to %s", p_filename);
    return ok;
}

b32 tgb200_scene_load(tg_raytracer* p_raytracer, const char* p_filename)
{
    if (!p_raytracer || !p_raytracer->p_device) { tgb_set_error("tgb200_scene_load: raytracer is not alive"); return TG_FALSE; }
    FILE* p_file = fopen(p_filename, "rb");
    if (!p_file) { tgb_set_error("tgb200_scene_load: cannot open %s", p_filename); return TG_FALSE; }
    char magic[8];
    u32 version = 0, n_objects = 0, n_luts = 0;
    b32 ok = fread(magic, 1, 8, p_file) == 8 && memcmp(magic, TGB_SCENE_MAGIC, 8) == 0 && fread(&version, 4, 1, p_file) == 1 && version == 1
          && fread(&n_objects, 4, 1, p_file) == 1 && fread(&n_luts, 4, 1, p_file) == 1;
    if (!ok) tgb_set_error("tgb200_scene_load: %s is not a version-1 scene file", p_filename);
    if (ok && n_luts > p_raytracer->n_color_luts) { tgb_set_error("tgb200_scene_load: the file holds %u colour LUTs, the raytracer %u", n_luts, p_raytracer->n_color_luts); ok = TG_FALSE; }
    for (u32 l = 0; ok && l < n_luts; l++)
    {
        u32 lut[256];
        ok = fread(lut, 4, 256, p_file) == 256 && tgbd_upload(p_raytracer->p_device, TGB_BUF_COLOR_LUT, (u64)l * 256u * 4u, lut, sizeof(lut));
    }
    for (u32 i = 0; ok && i < n_objects; i++)
    {
        v3u dims; v3 translation, axis; f32 angle = 0.0f; u32 lut_idx = 0;
        ok = fread(&dims, sizeof(v3u), 1, p_file) == 1 && fread(&translation, sizeof(v3), 1, p_file) == 1 && fread(&angle, 4, 1, p_file) == 1
          && fread(&axis, sizeof(v3), 1, p_file) == 1 && fread(&lut_idx, 4, 1, p_file) == 1;
        if (!ok) break;
        const u64 n = (u64)dims.x * dims.y * dims.z;
        if (n == 0 || n > p_raytracer->scene.n_available_cluster_indices) { tgb_set_error("tgb200_scene_load: object %u needs %llu clusters, %u are free", i, (unsigned long long)n, p_raytracer->scene.n_available_cluster_indices); ok = TG_FALSE; break; }
        u32* p_bits = (u32*)malloc((size_t)n * 64u);
        u8* p_materials = (u8*)malloc((size_t)n * 512u);
        ok = p_bits && p_materials && fread(p_bits, 64, (size_t)n, p_file) == (size_t)n && fread(p_materials, 512, (size_t)n, p_file) == (size_t)n;
        if (ok)
        {
            const v3u extent = { dims.x * 8u, dims.y * 8u, dims.z * 8u };
            ok = tg_raytracer_create_object_from_data(p_raytracer, translation, extent, angle, axis, lut_idx, p_bits, p_materials) != TG_U32_MAX;
        }
        free(p_bits); free(p_materials);
    }
    fclose(p_file);
    if (!ok && !tgb200_last_error()) tgb_set_error("tgb200_scene_load: %s is truncated", p_filename);
    return ok;
}

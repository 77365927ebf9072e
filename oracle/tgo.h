/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline /
 * --impl reference). Nothing under tg_b200/ may include, link or call this.
 *
 * Scalar C restatement of the reference's voxel-rendering path (its shader logic + its CPU SVO
 * builder). PARITY PINNING: the reference has no tests, golden vectors or known-answer fixtures of
 * any kind for this path (SURVEY.md sections 4 and 8c), and its application (Win32 + Vulkan + GLSL) cannot
 * run here. What CAN be built is the reference's portable C: math/tg_math.c, physics/tg_physics.c,
 * util/tg_amanatides_woo.c and graphics/tg_sparse_voxel_octree.c compile under gcc from where they lie
 * (oracle/Makefile -> oracle/_ref/libtg_ref.so). tests/test_reference_pins.py runs THE REFERENCE'S OWN
 * tg_svo_create, tg_svo_traverse, matrix / noise / RNG routines, slab and SAT tests on the same inputs as
 * this restatement and demands identical bits: the SVO builder, the CPU traversal and the whole math layer
 * are pinned by reference-run outputs. The shader logic (visibility.frag, shading.frag, svo_functions.inc)
 * has no runnable reference (no Vulkan ICD, no GLSL compiler in the image): it is pinned by literal
 * transcription with file:line provenance on top of that pinned math layer, by its C twins where the
 * reference has one (tg_svo_traverse, tg_intersect_ray_aabb, tg_amanatides_woo), and by hand-derived
 * known answers in tests/. The transcription of svo_functions.inc is additionally run against the
 * reference's C traversal (same hit / miss decision, leaf node and voxel on 40,000 random rays; the distance
 * to 1e-4, the two variants accumulate it differently). For visibility.frag's assembly of those pieces and
 * for shading.frag: "parity unpinned" by reference-run outputs.
 */
#ifndef TGO_H
#define TGO_H

#include "../include/tg_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Flat, borrowed view of a scene (the arrays D1-D6 of SURVEY.md section 8a). */
typedef struct tgo_scene_view
{
    u32                   n_objects_capacity;
    const tg_object_data* p_objects;            /* D5, indexed by object idx */
    u32                   n_cluster_pointers;
    const u32*            p_cluster_pointers;   /* D3 */
    const u32*            p_cluster_idx_to_object_idx; /* D4 */
    const u32*            p_voxel_cluster_data; /* D1, 16 u32 per cluster idx */
    const u8*             p_color_lut_idx_data; /* D2, 512 u8 per cluster idx (may be NULL for visibility) */
    const u32*            p_color_lut;          /* D6, 256 u32 per LUT */
    u32                   global_pointer_base;  /* multi-GPU: added to the pointer packed into the word */
} tgo_scene_view;

enum { TGO_VIS_BRUTE_FORCE = 0, TGO_VIS_SCREEN_RECT = 1 };

/* camera: tgvk_core.c:382-444, tgvk_raytracer.c:1171-1180 */
void tgo_camera_rays(const tg_camera* p_camera, tg_camera_rays* p_out);
/* tgvk_raytracer.c:836-847 */
void tgo_object_data(const tg_voxel_object* p_object, u32 lut_idx, tg_object_data* p_out);
/* tgvk_raytracer.c:1130-1134 */
u32  tgo_pack_color(f32 r, f32 g, f32 b);
/* visibility.frag:35-62 */
m4   tgo_ws2ms(const tg_object_data* p_object, u32 cluster_pointer);
/* sh/common.inc:40-46 + visibility.frag:29-33 */
v3   tgo_pixel_ray_direction_nn(const tg_camera_rays* p_cam, u32 w, u32 h, u32 px, u32 py);
/* visibility.frag:22-208 for ONE fragment, nothing hoisted. Returns the packed word or TG_VIS_CLEAR. */
u64  tgo_visibility_fragment(const tgo_scene_view* p_scene, const tg_camera_rays* p_cam, u32 w, u32 h, u32 px, u32 py, u32 cluster_pointer);
/*
 * clear.comp + visibility pass over rows y0, y0+ystep, ... < y1 (other rows are left TG_VIS_CLEAR).
 * mode BRUTE_FORCE: every pixel x every cluster. mode SCREEN_RECT: per cluster, only the pixels in a
 * conservative screen rectangle of its box (what the rasteriser's coverage prunes).
 * Returns the number of (pixel, cluster) fragments evaluated.
 */
u64  tgo_visibility(const tgo_scene_view* p_scene, const tg_camera_rays* p_cam, u32 w, u32 h, u32 mode, u32 y0, u32 y1, u32 ystep, u64* p_out);
/* every pixel of the window [x0, x1) x [y0, y1) against EVERY cluster pointer, unpruned (full-resolution frames) */
u64  tgo_visibility_window(const tgo_scene_view* p_scene, const tg_camera_rays* p_cam, u32 w, u32 h, u32 x0, u32 x1, u32 y0, u32 y1, u64* p_out);
i32  tgo_max_threads(void);
void tgo_set_threads(i32 n);

/* DDA of visibility.frag:83-191 on an arbitrary 8^3 mask, exposed for cross-checks. Returns voxel idx or -1. */
i32  tgo_cluster_dda(const u32* p_mask16, v3 o_ms, v3 d_ms, f32 enter);

/* SVO: tg_sparse_voxel_octree.c:23-542 */
void tgo_svo_create(v3 extent_min, v3 extent_max, const tgo_scene_view* p_scene, const tg_voxel_object* p_objects, tg_svo* p_svo);
void tgo_svo_destroy(tg_svo* p_svo);
/* svo_functions.inc:1-329 (GLSL). Returns depth in [0,1) on hit, 1.0 on miss. */
f32  tgo_svo_traverse_glsl(const tg_svo* p_svo, f32 far_plane, v3 ray_origin_ws, v3 ray_direction_ws, v3* p_hit_position, v3* p_hit_normal, u32* p_node_idx, u32* p_voxel_idx);
/* debug_visibility_svo.frag:27-71 (the BLOCKS view's primary-ray pass through the SVO, tgvk_raytracer.c:1226-1272); rows y0, y0+ystep, ... < y1 */
void tgo_visibility_svo(const tg_svo* p_svo, const tg_camera_rays* p_cam, u32 w, u32 h, u32 y0, u32 y1, u32 ystep, u64* p_out);
/* tg_sparse_voxel_octree.c:558-740 (C twin, uses the Amanatides-Woo of util/tg_amanatides_woo.c:3-114) */
b32  tgo_svo_traverse_c(const tg_svo* p_svo, v3 ray_origin, v3 ray_direction, f32* p_distance, u32* p_node_idx, u32* p_voxel_idx);
/* util/tg_amanatides_woo.c:3-114 restated */
b32  tgo_amanatides_woo(v3 ray_hit_on_grid, v3 ray_direction, v3 extent, const u32* p_voxel_grid, v3i* p_voxel_id);
/* physics/tg_physics.c:226-392 */
b32  tgo_intersect_aabb_obb_ignore_contact(v3 bmin, v3 bmax, const v3* p_obb_corners);

/* math/tg_math.c:182-302 */
f32  tgo_simplex_noise(f32 x, f32 y, f32 z);
/* tgvk_raytracer.c:871-943: the reference's procedural terrain bits of one object (16 u32 per cluster, pointer order) */
b32  tgo_procedural_voxel_is_solid(u32 object_idx, u32 voxel_x, u32 voxel_y, u32 voxel_z);
void tgo_procedural_solid_bits(u32 object_idx, v3u dims, u32* p_out);

/* shading.frag:114-337 (+ the pinned GI term of DESIGN.md) for rows y0, y0+ystep, ... < y1 (other rows untouched). RGBA32F out. */
void tgo_shade(const tgo_scene_view* p_scene, const tg_camera_rays* p_cam, u32 w, u32 h, const u64* p_vis, const tg_svo* p_svo_or_null,
               u32 gi_enabled, u32 frame_seed, u32 debug_visualization, u32 y0, u32 y1, u32 ystep, f32* p_out_rgba);

/* the secondary rays tgo_shade traces for those rows (6 floats per pixel: origin, direction; direction 0 = no ray) -- workload studies */
void tgo_shade_gi_rays(const tgo_scene_view* p_scene, const tg_camera_rays* p_cam, u32 w, u32 h, const u64* p_vis, const tg_svo* p_svo,
                       u32 frame_seed, u32 y0, u32 y1, u32 ystep, f32* p_out_rays);

/* present.frag + B8G8R8A8_UNORM conversion of n_pixels RGBA32F pixels (see tgo_shade.c) */
void tgo_present_bgra8(const f32* p_rgba, u64 n_pixels, u32* p_out);

#ifdef __cplusplus
}
#endif

#endif

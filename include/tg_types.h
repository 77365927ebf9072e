/*
 * tg_types.h -- plain-C data model shared by the boundary, the host code and the CUDA kernels.
 *
 * Layout-compatible restatement of the reference's public types (paths relative to
 * /root/reference/tg/src):
 *   scalar aliases            tg_common.h:16-22,100-105
 *   v3 / v3i / v3u / v4 / m4  math/tg_math.h:72-275   (m4 is COLUMN-major: m00 m10 m20 m30 m01 ...)
 *   tg_camera                 graphics/tg_graphics_core.h:43-47,111-131
 *   tg_voxel_cluster          graphics/tg_graphics_core.h:138-141  (8^3 voxels, 1 bit each, 64 B)
 *   tg_voxel_object           graphics/tg_graphics_core.h:143-150  (44 B CPU record)
 *   tg_svo*                   graphics/tg_sparse_voxel_octree.h:6-47
 *   tg_scene                  graphics/vulkan/tgvk_raytracer.h:90-108
 *   tg_debug_show             graphics/vulkan/tgvk_raytracer.h:75-88
 *
 * No Vulkan, no CUDA and no torch types appear here: this header is what a TG maintainer includes.
 */
#ifndef TG_TYPES_H
#define TG_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t  b32;
typedef float    f32;
typedef double   f64;
typedef int8_t   i8;
typedef int32_t  i32;
typedef int64_t  i64;
typedef uint8_t  u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

#define TG_FALSE 0
#define TG_TRUE  1

#define TG_F32_MAX      3.402823466e+38f
#define TG_F32_MIN      (-TG_F32_MAX)
#define TG_F32_EPSILON  1.192092896e-07f
#define TG_U32_MAX      0xffffffffu

/* cluster constants: graphics/tg_graphics_core.h:18-23 == shaders/raytracer/cluster.inc:1-6 */
#define TG_PRIMITIVE_IDX_N_BITS                9
#define TG_N_PRIMITIVES_PER_CLUSTER            (1 << TG_PRIMITIVE_IDX_N_BITS) /* 512 */
#define TG_N_PRIMITIVES_PER_CLUSTER_CUBE_ROOT  8
#define TG_CLUSTERS_IDX_N_BITS                 31
#define TG_CLUSTER_SIZE(n_bits_per_element)    (TG_N_PRIMITIVES_PER_CLUSTER / 8 * (n_bits_per_element))
#define TG_CLUSTER_MASK_WORDS                  (TG_N_PRIMITIVES_PER_CLUSTER / 32) /* 16 x u32 = 64 B */

/* SVO constants: graphics/tg_sparse_voxel_octree.c:11-16 == shaders/raytracer/svo.inc:1-4 */
#define TG_SVO_SIDE_LENGTH                1024
#define TG_SVO_BLOCK_SIDE_LENGTH          32
#define TG_SVO_BLOCK_VOXEL_COUNT          (TG_SVO_BLOCK_SIDE_LENGTH * TG_SVO_BLOCK_SIDE_LENGTH * TG_SVO_BLOCK_SIDE_LENGTH)
#define TG_SVO_BLOCK_WORDS                (TG_SVO_BLOCK_VOXEL_COUNT / 32) /* 1024 x u32 = 4 KiB */
#define TG_SVO_TRAVERSE_STACK_CAPACITY    5
#define TG_SVO_LEAF_MAX_CLUSTERS          64

/* visibility word: 24 b depth | 31 b cluster POINTER | 9 b voxel  (visibility.frag:198-201) */
#define TG_VIS_CLEAR                      0xFFFFFFFFFFFFFFFFull /* clear.comp:19 (intent: all ones) */
#define TG_VIS_DEPTH_SHIFT                40
#define TG_VIS_POINTER_SHIFT              9
#define TG_VIS_DEPTH_SCALE                16777215.0f

typedef struct v3  { f32 x, y, z; }    v3;
typedef struct v3i { i32 x, y, z; }    v3i;
typedef struct v3u { u32 x, y, z; }    v3u;
typedef struct v4  { f32 x, y, z, w; } v4;

/* column-major 4x4; m<row><col> */
typedef struct m4
{
    f32 m00, m10, m20, m30;
    f32 m01, m11, m21, m31;
    f32 m02, m12, m22, m32;
    f32 m03, m13, m23, m33;
} m4;

typedef enum tg_camera_type
{
    TG_CAMERA_TYPE_ORTHOGRAPHIC,
    TG_CAMERA_TYPE_PERSPECTIVE
} tg_camera_type;

typedef struct tg_camera
{
    tg_camera_type type;
    v3             position;
    f32            pitch;
    f32            yaw;
    f32            roll;
    union
    {
        struct { f32 l, r, b, t, n, f; }              ortho;
        struct { f32 fov_y_in_radians, aspect, n, f; } persp;
    };
} tg_camera;

typedef struct tg_voxel_cluster
{
    u32 p_data[TG_CLUSTER_MASK_WORDS]; /* bit 64*z + 8*y + x, LSB first inside word idx/32 */
} tg_voxel_cluster;

typedef struct tg_voxel_object
{
    v3u n_cluster_pointers_per_dim;
    u32 first_cluster_pointer;
    v3  translation;
    f32 angle_in_radians;
    v3  axis;
} tg_voxel_object;

typedef struct tg_svo_inner_node
{
    u16 child_pointer; /* first child, relative to this node, in nodes */
    u8  valid_mask;
    u8  leaf_mask;
} tg_svo_inner_node;

typedef struct tg_svo_leaf_node_data
{
    u32 n;
    u32 p_cluster_idcs[TG_SVO_LEAF_MAX_CLUSTERS];
} tg_svo_leaf_node_data;

typedef struct tg_svo_leaf_node
{
    u32 data_pointer;
} tg_svo_leaf_node;

typedef union tg_svo_node
{
    tg_svo_inner_node inner;
    tg_svo_leaf_node  leaf;
} tg_svo_node;

typedef struct tg_svo
{
    v3                     min;
    v3                     max;
    u32                    voxel_buffer_capacity_in_u32;
    u32                    voxel_buffer_count_in_u32;
    u32                    leaf_node_data_buffer_capacity;
    u32                    leaf_node_data_buffer_count;
    u32                    node_buffer_capacity;
    u32                    node_buffer_count;
    u32*                   p_voxels_buffer;         /* 32^3 bits per leaf, bit 1024*z + 32*y + x */
    tg_svo_leaf_node_data* p_leaf_node_data_buffer;
    tg_svo_node*           p_node_buffer;           /* node 0 is the root inner node */
} tg_svo;

typedef struct tg_scene
{
    u32              object_capacity;
    u32              n_objects;
    tg_voxel_object* p_objects;
    u32              n_available_object_indices;
    u32*             p_available_object_indices;

    u32              cluster_pointer_capacity;
    u32              n_cluster_pointers;
    u32*             p_cluster_pointers;
    u32              n_available_cluster_indices;
    u32*             p_available_cluster_indices;

    u32*             p_voxel_cluster_data;        /* 16 x u32 per cluster, indexed by cluster IDX */
    u32*             p_cluster_idx_to_object_idx;

    tg_svo           svo;
} tg_scene;

typedef enum tg_debug_show
{
    TG_DEBUG_SHOW_NONE            = 0,
    TG_DEBUG_SHOW_OBJECT_INDEX    = 1,
    TG_DEBUG_SHOW_DEPTH           = 2,
    TG_DEBUG_SHOW_CLUSTER_INDEX   = 3,
    TG_DEBUG_SHOW_VOXEL_INDEX     = 4,
    TG_DEBUG_SHOW_BLOCKS          = 5,
    TG_DEBUG_SHOW_COLOR_LUT_INDEX = 6,
    TG_DEBUG_SHOW_COLOR           = 7,
    TG_DEBUG_SHOW_NORMAL          = 8,
    TG_DEBUG_SHOW_SHADING         = 9,
    TG_DEBUG_SHOW_COUNT
} tg_debug_show;

/*
 * GPU-side object record, std430, 96 B (graphics/vulkan/tgvk_raytracer.c:35-42 ==
 * shaders/raytracer/buffers.inc:6-13). `pad` carries the per-object LUT index in this build
 * (README.md:12,18 "LUT index per object"; the reference leaves it 0, tgvk_raytracer.c:1124).
 */
typedef struct tg_object_data
{
    v3u n_cluster_pointers_per_dim;
    u32 first_cluster_pointer;
    v3  translation;
    u32 lut_idx;
    m4  rotation;
} tg_object_data;

/* camera block consumed by the kernels (graphics/vulkan/tgvk_raytracer.c:65-74) */
typedef struct tg_camera_rays
{
    v4  camera;
    v4  ray_bl;
    v4  ray_br;
    v4  ray_tr;
    v4  ray_tl;
    f32 near_plane;
    f32 far_plane;
    f32 pad[2];
} tg_camera_rays;

#ifdef __cplusplus
}
#endif

#endif

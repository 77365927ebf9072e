/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 * Link shim for the reference's portable C files built into oracle/_ref/libtg_ref.so (oracle/Makefile): math/tg_math.c,
 * physics/tg_physics.c, util/tg_amanatides_woo.c and graphics/tg_sparse_voxel_octree.c are compiled from where they lie
 * under /root/reference (never copied; the recipe only rewrites MSVC integer-literal suffixes such as 1ui32 in the
 * preprocessed stream). What they call outside those four files is Win32 / Vulkan code that cannot be built here; the
 * handful of small functions involved are restated below against the reference's own headers:
 *   tgp_malloc / tgp_realloc / tgp_free     platform/tg_platform_win32.c:127-150 (HeapAlloc with HEAP_ZERO_MEMORY: zero-filled)
 *   tg_memory_stack_alloc / _free           memory/tg_memory.c:267-323 (release build: a linear bump allocator, not cleared)
 *   tg_memory_nullify, tg_memcpy            memory/tg_memory.c:177-196
 *   tg_object_is_initialized                graphics/vulkan/tgvk_raytracer.c:1069-1077
 */
#include <stdlib.h>
#include <string.h>

#include "memory/tg_memory.h"
#include "graphics/vulkan/tgvk_raytracer.h"

void* tgp_malloc(tg_size size) { return calloc(1, (size_t)size ? (size_t)size : 1); }
void* tgp_realloc(tg_size size, void* p_memory) { return realloc(p_memory, (size_t)size); }
void  tgp_free(void* p_memory) { free(p_memory); }

/* one linear stack like the reference's per-thread 1 GiB arena; pages are committed lazily by the OS */
#define TGO_REF_STACK_SIZE ((size_t)1 << 32)
static unsigned char* p_stack;
static size_t stack_exhausted;

void* tg_memory_stack_alloc(tg_size size)
{
    if (!p_stack) p_stack = (unsigned char*)malloc(TGO_REF_STACK_SIZE);
    if (!p_stack || stack_exhausted + (size_t)size > TGO_REF_STACK_SIZE) abort();
    void* p = p_stack + stack_exhausted;
    stack_exhausted += (size_t)size;
    return p;
}
void tg_memory_stack_free(tg_size size) { stack_exhausted -= (size_t)size; }
void tg_memory_stack_resize(tg_size old_size, tg_size new_size) { stack_exhausted += (size_t)(new_size - old_size); }

void tg_memory_nullify(tg_size size, void* p_memory) { memset(p_memory, 0, (size_t)size); }
void tg_memcpy(tg_size size, const void* p_source, void* p_destination) { memcpy(p_destination, p_source, (size_t)size); }

b32 tg_object_is_initialized(const tg_scene* p_scene, u32 object_idx)
{
    const tg_voxel_object* p_object = &p_scene->p_objects[object_idx];
    return p_object->n_cluster_pointers_per_dim.x != 0 && p_object->n_cluster_pointers_per_dim.y != 0 && p_object->n_cluster_pointers_per_dim.z != 0;
}

/*
 * tgb_svo.cu -- K2, the parallel build of the 1-bit sparse voxel octree.
 *
 * Replaces the reference's single-threaded recursive CPU builder
 *   tg_svo_create              graphics/tg_sparse_voxel_octree.c:466-542
 *   tg__construct_inner_node   graphics/tg_sparse_voxel_octree.c:319-464
 *   tg__construct_leaf_node    graphics/tg_sparse_voxel_octree.c:102-317
 *   AABB/OBB face-axis SAT     physics/tg_physics.c:226-392
 * and must reproduce its three arrays BIT FOR BIT (oracle twin: oracle/tgo_svo.c), including the DFS allocation
 * order of nodes and leaves and the quirks of SURVEY.md appendix A (Q4 corner 3 skipped, Q5 face axes only,
 * trilinear voxel-centre mapping).
 *
 * The tree is at most 1 + 8 + 64 + 512 + 4096 inner nodes + 32768 leaves (1024^3 world box, 32^3 leaf blocks,
 * tg_sparse_voxel_octree.c:11-16), so every node has a DENSE address (level, path) with path = 3 bits per level from
 * the root. The recursion becomes four data-parallel passes:
 *   1. k_svo_descend<false>  one thread per cluster pointer: skip objects that cannot touch the box and empty
 *                            masks (:515-531), build the cluster's ws->cluster-space matrix ONCE (the reference
 *                            rebuilds it per node, Q6), then walk down the dense tree running the reference's SAT
 *                            against the 8 children of every node the cluster reaches; count arrivals per node.
 *   2. k_svo_layout          one CTA: a node exists iff a cluster reached it; bottom-up subtree sums and a top-down
 *                            sweep give every node the index the reference's DFS allocation would (children of a
 *                            node are consecutive, allocated when the parent is visited; leaves numbered in DFS order),
 *                            write the node array, scan the per-leaf pair counts.
 *   3. k_svo_descend<true>   same walk, now scattering (leaf, cluster pointer) pairs into per-leaf segments.
 *   4. k_svo_fill_leaves     one CTA per leaf, one warp per (leaf, cluster) pair, per z slice the (y, x) samples of the
 *                            clipped range dealt out to the lanes: the reference's trilinear sampling, bits OR-ed into a 4 KiB shared-memory block (order
 *                            independent), then the first 64 contributing cluster indices in ascending pointer order.
 * Everything that decides a bit or an index is the reference's arithmetic, operation for operation (-fmad=false).
 */
#include <stdlib.h>
#include <string.h>

#include "tgb_device.cuh"

#define TGB_SVO_LEVELS        5u      /* inner levels 0..4, leaves at level 5 */
#define TGB_SVO_DENSE_TOTAL   37449u  /* 1 + 8 + 64 + 512 + 4096 + 32768 */
#define TGB_SVO_MAX_LEAVES    32768u

__host__ __device__ __forceinline__ u32 tgb_level_offset(u32 level)
{
    /* (8^level - 1) / 7 */
    return level == 0 ? 0u : (level == 1 ? 1u : (level == 2 ? 9u : (level == 3 ? 73u : (level == 4 ? 585u : 4681u))));
}

/* scratch layout (u32 words) */
enum
{
    TGB_SCR_CNT   = 0,                                   /* arrivals per dense node            */
    TGB_SCR_S     = TGB_SCR_CNT   + TGB_SVO_DENSE_TOTAL, /* sum of child counts in the subtree */
    TGB_SCR_LS    = TGB_SCR_S     + TGB_SVO_DENSE_TOTAL, /* leaves in the subtree              */
    TGB_SCR_BASE  = TGB_SCR_LS    + TGB_SVO_DENSE_TOTAL, /* nodes allocated before the visit   */
    TGB_SCR_LBASE = TGB_SCR_BASE  + TGB_SVO_DENSE_TOTAL, /* leaves numbered before the visit   */
    TGB_SCR_NIDX  = TGB_SCR_LBASE + TGB_SVO_DENSE_TOTAL, /* node index (levels 0..5)           */
    TGB_SCR_POFF  = TGB_SCR_NIDX  + TGB_SVO_DENSE_TOTAL, /* [32768] pair offset per dense leaf  */
    TGB_SCR_CUR   = TGB_SCR_POFF  + TGB_SVO_MAX_LEAVES,  /* [32768] scatter cursor             */
    TGB_SCR_DENSE_OF_LEAF = TGB_SCR_CUR + TGB_SVO_MAX_LEAVES, /* [32768] dense leaf per data_pointer */
    TGB_SCR_DIRTY = TGB_SCR_DENSE_OF_LEAF + TGB_SVO_MAX_LEAVES, /* [32768] leaf must be re-sampled (incremental) */
    TGB_SCR_CNT_G = TGB_SCR_DIRTY + TGB_SVO_MAX_LEAVES,  /* arrivals per dense node summed over all ranks (== CNT on one GPU) */
    TGB_SCR_CLEARED = TGB_SCR_CNT_G + TGB_SVO_DENSE_TOTAL, /* everything above is zeroed at the start of a build */
    TGB_SCR_PREV_DP = TGB_SCR_CLEARED,                    /* [32768] data_pointer + 1 of the dense leaf in the PREVIOUS build, 0 = none */
    TGB_SCR_TOTAL = TGB_SCR_PREV_DP + TGB_SVO_MAX_LEAVES
};

/* counts block d_counts: [0] nodes, [1] leaves, [2] pairs, [3] error flags */

/* ---- physics/tg_physics.c:226-392 against the cluster box [0,8]^3 ------------------------------- */
__device__ bool tgb_intersect_cluster_box_obb(const v3* c)
{
    bool separated;
    separated = true; for (u32 i = 0; i < 8; i++) separated &= c[i].x <= 0.0f; if (separated) return false;
    separated = true; for (u32 i = 0; i < 8; i++) separated &= c[i].x >= 8.0f; if (separated) return false;
    separated = true; for (u32 i = 0; i < 8; i++) separated &= c[i].y <= 0.0f; if (separated) return false;
    separated = true; for (u32 i = 0; i < 8; i++) separated &= c[i].y >= 8.0f; if (separated) return false;
    separated = true; for (u32 i = 0; i < 8; i++) separated &= c[i].z <= 0.0f; if (separated) return false;
    separated = true; for (u32 i = 0; i < 8; i++) separated &= c[i].z >= 8.0f; if (separated) return false;

    /* :332-391 OBB faces */
    const v3 nxp = tgb_normalize(tgb_sub(c[1], c[0]));
    const v3 nyp = tgb_normalize(tgb_sub(c[2], c[0]));
    const v3 nzp = tgb_normalize(tgb_sub(c[4], c[0]));
    const f32 l0 = tgb_length(c[0]);
    const f32 l1 = tgb_length(c[7]);
    const v3 n0 = l0 == 0.0f ? tgb_v3(0.0f, 0.0f, 0.0f) : tgb_divf(c[0], l0);
    const v3 n1 = l1 == 0.0f ? tgb_v3(0.0f, 0.0f, 0.0f) : tgb_divf(c[7], l1);

#pragma unroll
    for (u32 plane = 0; plane < 6; plane++)
    {
        const v3 axis = plane < 2 ? nxp : (plane < 4 ? nyp : nzp);
        const v3 normal = (plane & 1u) ? axis : tgb_neg(axis);
        const f32 distance = (plane & 1u) ? l1 * tgb_dot(n1, normal) : l0 * tgb_dot(n0, normal);
        separated = true;
#pragma unroll
        for (u32 k = 0; k < 8; k++)
        {
            const v3 pt = tgb_v3((k & 1u) ? 8.0f : 0.0f, (k & 2u) ? 8.0f : 0.0f, (k & 4u) ? 8.0f : 0.0f);
            const f32 dist = tgb_dot(normal, pt) - distance;
            separated &= dist >= 0.0f;
        }
        if (separated) return false;
    }
    return true;
}

/* tg_sparse_voxel_octree.c:23-60 from the 96-byte object record (rotation built on the host, tgvk_raytracer.c:841) */
__device__ m4 tgb_svo_ws2ms(const tg_object_data& o, u32 rel)
{
    const u32 rx = rel % o.n_cluster_pointers_per_dim.x;
    const u32 ry = (rel / o.n_cluster_pointers_per_dim.x) % o.n_cluster_pointers_per_dim.y;
    const u32 rz = rel / (o.n_cluster_pointers_per_dim.x * o.n_cluster_pointers_per_dim.y);
    const v3 off = tgb_v3((f32)(rx * 8u), (f32)(ry * 8u), (f32)(rz * 8u));
    const v3 half = tgb_mul(tgb_v3((f32)o.n_cluster_pointers_per_dim.x, (f32)o.n_cluster_pointers_per_dim.y, (f32)o.n_cluster_pointers_per_dim.z), tgb_v3(4.0f, 4.0f, 4.0f));
    const m4 ws2ms1 = tgb_m4_translate(tgb_neg(o.translation));
    const m4 ws2ms2 = tgb_m4_inverse(o.rotation);
    const m4 ws2ms3 = tgb_m4_translate(half);
    const m4 ws2ms4 = tgb_m4_translate(tgb_neg(off));
    return tgb_m4_mul(tgb_m4_mul(tgb_m4_mul(ws2ms4, ws2ms3), ws2ms2), ws2ms1);
}

/* tg_sparse_voxel_octree.c:62-100 */
__device__ m4 tgb_svo_cs2ws(const tg_object_data& o, u32 rel)
{
    const u32 rx = rel % o.n_cluster_pointers_per_dim.x;
    const u32 ry = (rel / o.n_cluster_pointers_per_dim.x) % o.n_cluster_pointers_per_dim.y;
    const u32 rz = rel / (o.n_cluster_pointers_per_dim.x * o.n_cluster_pointers_per_dim.y);
    const v3 off = tgb_v3((f32)(rx * 8u), (f32)(ry * 8u), (f32)(rz * 8u));
    const v3 half = tgb_mul(tgb_v3((f32)o.n_cluster_pointers_per_dim.x, (f32)o.n_cluster_pointers_per_dim.y, (f32)o.n_cluster_pointers_per_dim.z), tgb_v3(4.0f, 4.0f, 4.0f));
    const m4 ms2ws1 = tgb_m4_translate(off);
    const m4 ms2ws2 = tgb_m4_translate(tgb_neg(half));
    const m4 ms2ws3 = o.rotation;
    const m4 ms2ws4 = tgb_m4_translate(o.translation);
    return tgb_m4_mul(tgb_m4_mul(tgb_m4_mul(ms2ws4, ms2ws3), ms2ws2), ms2ws1);
}

/* the child box of tg__construct_inner_node (:359-366), operation for operation */
__device__ __forceinline__ void tgb_child_box(v3 parent_min, v3 parent_max, u32 child_idx, v3* p_min, v3* p_max)
{
    const v3 child_extent = tgb_scale(tgb_sub(parent_max, parent_min), 0.5f);
    const f32 dx = (f32)( child_idx      % 2u) * child_extent.x;
    const f32 dy = (f32)((child_idx / 2u) % 2u) * child_extent.y;
    const f32 dz = (f32)((child_idx / 4u) % 2u) * child_extent.z;
    *p_min = tgb_add(parent_min, tgb_v3(dx, dy, dz));
    *p_max = tgb_add(*p_min, child_extent);
}

/* box of the dense node (level, path) by the same chain of additions from the root box */
__device__ __forceinline__ void tgb_dense_box(v3 bmin, v3 bmax, u32 level, u32 path, v3* p_min, v3* p_max)
{
    v3 lo = bmin, hi = bmax;
    for (u32 l = 1; l <= level; l++)
    {
        const u32 oct = (path >> (3u * (level - l))) & 7u;
        v3 cmin, cmax;
        tgb_child_box(lo, hi, oct, &cmin, &cmax);
        lo = cmin; hi = cmax;
    }
    *p_min = lo; *p_max = hi;
}

/* ---- pass 0: which objects can touch the box at all ------------------------------------------------ */
/*
 * Conservative: an object is skipped only when its bounding sphere lies beyond one face plane of the root box by a
 * margin (4 voxels + 1e-4 of the coordinate magnitude) that dwarfs the float error of the reference's plane test;
 * then the OBB-face test of tg_physics.c:332-391 separates each of its clusters from every root child.
 */
__global__ void k_svo_object_flags(const tg_object_data* __restrict__ p_objects, u32 object_capacity, v3 bmin, v3 bmax, u32* __restrict__ p_flags)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= object_capacity) return;
    const tg_object_data& o = p_objects[i];
    const f32 nx = (f32)o.n_cluster_pointers_per_dim.x, ny = (f32)o.n_cluster_pointers_per_dim.y, nz = (f32)o.n_cluster_pointers_per_dim.z;
    u32 flag = 0;
    if (nx != 0.0f && ny != 0.0f && nz != 0.0f)
    {
        const f32 r = 4.0f * sqrtf(nx * nx + ny * ny + nz * nz) * 1.0001f;
        const f32 mag = fmaxf(fmaxf(fabsf(o.translation.x), fabsf(o.translation.y)), fabsf(o.translation.z)) + r
                      + fmaxf(fmaxf(fabsf(bmin.x), fabsf(bmin.y)), fabsf(bmin.z)) + 1024.0f;
        const f32 m = r + 4.0f + 1e-4f * mag;
        const bool outside = o.translation.x + m < bmin.x || o.translation.x - m > bmax.x
                          || o.translation.y + m < bmin.y || o.translation.y - m > bmax.y
                          || o.translation.z + m < bmin.z || o.translation.z - m > bmax.z;
        flag = outside ? 0u : 1u;
        if (!(mag == mag) || isinf(mag)) flag = 1u;
    }
    p_flags[i] = flag;
}

/* ---- passes 1 and 3: the per-cluster walk ------------------------------------------------------------- */
template <bool SCATTER>
__global__ void __launch_bounds__(128) k_svo_descend(const u32* __restrict__ p_cluster_pointers, const u32* __restrict__ p_c2o, const tg_object_data* __restrict__ p_objects,
                                                     const u32* __restrict__ p_object_flags, const u32* __restrict__ p_masks, u32 n_cluster_pointers, v3 bmin, v3 bmax,
                                                     u32* __restrict__ p_scratch, u32* __restrict__ p_pairs, u32* __restrict__ p_pair_leaf, const u32* __restrict__ p_moved)
{
    const u32 cluster_pointer = blockIdx.x * blockDim.x + threadIdx.x;
    if (cluster_pointer >= n_cluster_pointers) return;
    const u32 cluster_idx = __ldg(&p_cluster_pointers[cluster_pointer]);
    const u32 object_idx = __ldg(&p_c2o[cluster_idx]);
    if (__ldg(&p_object_flags[object_idx]) == 0) return;

    /* :515-531: clusters without voxels are not candidates */
    const uint4* p_mask = reinterpret_cast<const uint4*>(p_masks + (u64)cluster_idx * TG_CLUSTER_MASK_WORDS);
    const uint4 m0 = __ldg(&p_mask[0]), m1 = __ldg(&p_mask[1]), m2 = __ldg(&p_mask[2]), m3 = __ldg(&p_mask[3]);
    if ((m0.x | m0.y | m0.z | m0.w | m1.x | m1.y | m1.z | m1.w | m2.x | m2.y | m2.z | m2.w | m3.x | m3.y | m3.z | m3.w) == 0) return;

    const tg_object_data o = p_objects[object_idx];
    const m4 ws2ms = tgb_svo_ws2ms(o, cluster_pointer - o.first_cluster_pointer);
    const bool moved = !SCATTER && p_moved != NULL && p_moved[object_idx] != 0;

    u32* __restrict__ p_cnt = p_scratch + TGB_SCR_CNT;

    /* iterative DFS: per level the node box and the next child to try */
    v3 lo[TGB_SVO_LEVELS], hi[TGB_SVO_LEVELS];
    u32 next_child[TGB_SVO_LEVELS];
    u32 path = 0; /* dense path of the node at `level` */
    i32 level = 0;
    lo[0] = bmin; hi[0] = bmax; next_child[0] = 0;
    while (level >= 0)
    {
        if (next_child[level] == 8u)
        {
            level--;
            path >>= 3;
            continue;
        }
        const u32 child = next_child[level]++;
        v3 cmin, cmax;
        tgb_child_box(lo[level], hi[level], child, &cmin, &cmax);
        /* :367-385 the 8 corners of the child box in cluster space */
        v3 corners[8];
#pragma unroll
        for (u32 k = 0; k < 8; k++)
        {
            const v3 p = tgb_v3((k & 1u) ? cmax.x : cmin.x, (k & 2u) ? cmax.y : cmin.y, (k & 4u) ? cmax.z : cmin.z);
            corners[k] = tgb_m4_transform(ws2ms, p, 1.0f);
        }
        if (!tgb_intersect_cluster_box_obb(corners)) continue;

        const u32 child_path = (path << 3) | child;
        const u32 child_level = (u32)level + 1u;
        if (!SCATTER)
        {
            atomicAdd(&p_cnt[tgb_level_offset(child_level) + child_path], 1u);
        }
        if (child_level == TGB_SVO_LEVELS)
        {
            if (SCATTER)
            {
                const u32 slot = p_scratch[TGB_SCR_POFF + child_path] + atomicAdd(&p_scratch[TGB_SCR_CUR + child_path], 1u);
                p_pairs[slot] = cluster_pointer;
                p_pair_leaf[slot] = child_path;
            }
            else if (moved) p_scratch[TGB_SCR_DIRTY + child_path] = 1u; /* a moved object reaches this leaf NOW */
        }
        else
        {
            level++;
            path = child_path;
            lo[level] = cmin; hi[level] = cmax; next_child[level] = 0;
        }
    }
}

/* ---- pass 2: DFS layout of the dense tree ---------------------------------------------------------- */
__global__ void __launch_bounds__(1024) k_svo_layout(u32* __restrict__ p_scratch, u32* __restrict__ p_nodes, u32* __restrict__ p_counts, u32 node_capacity, u32 leaf_capacity)
{
    /* existence and DFS indices come from the GLOBAL arrival counts, the pair segments from this rank's own */
    u32* __restrict__ p_cnt = p_scratch + TGB_SCR_CNT_G;
    const u32* __restrict__ p_cnt_local = p_scratch + TGB_SCR_CNT;
    u32* __restrict__ p_s = p_scratch + TGB_SCR_S;
    u32* __restrict__ p_ls = p_scratch + TGB_SCR_LS;
    u32* __restrict__ p_base = p_scratch + TGB_SCR_BASE;
    u32* __restrict__ p_lbase = p_scratch + TGB_SCR_LBASE;
    u32* __restrict__ p_nidx = p_scratch + TGB_SCR_NIDX;
    const u32 tid = threadIdx.x;

    if (tid == 0) p_cnt[0] = 1; /* the root always exists (:490-493) */
    /* bottom-up: S = children counts summed over the inner nodes of the subtree, LS = leaves of the subtree */
    for (u32 i = tid; i < TGB_SVO_MAX_LEAVES; i += blockDim.x) p_ls[tgb_level_offset(5) + i] = p_cnt[tgb_level_offset(5) + i] ? 1u : 0u;
    __syncthreads();
    for (i32 level = 4; level >= 0; level--)
    {
        const u32 n = 1u << (3 * level);
        const u32 off = tgb_level_offset((u32)level), coff = tgb_level_offset((u32)level + 1u);
        for (u32 i = tid; i < n; i += blockDim.x)
        {
            u32 c = 0, s = 0, ls = 0;
            for (u32 k = 0; k < 8; k++)
            {
                const u32 j = coff + 8u * i + k;
                if (p_cnt[j]) { c++; ls += p_ls[j]; if (level < 4) s += p_s[j]; }
            }
            p_s[off + i] = p_cnt[off + i] ? c + s : 0u;
            p_ls[off + i] = p_cnt[off + i] ? ls : 0u;
        }
        __syncthreads();
    }
    /* top-down: indices */
    if (tid == 0) { p_base[0] = 0; p_lbase[0] = 0; p_nidx[0] = 0; }
    __syncthreads();
    for (u32 level = 0; level < 5; level++)
    {
        const u32 n = 1u << (3 * level);
        const u32 off = tgb_level_offset(level), coff = tgb_level_offset(level + 1u);
        for (u32 i = tid; i < n; i += blockDim.x)
        {
            if (!p_cnt[off + i]) continue;
            u32 valid_mask = 0;
            for (u32 k = 0; k < 8; k++) if (p_cnt[coff + 8u * i + k]) valid_mask |= 1u << k;
            const u32 first_child = 1u + p_base[off + i];        /* node_buffer_count when this node is visited (:404) */
            u32 running = p_base[off + i] + (u32)__popc(valid_mask);
            u32 lrun = p_lbase[off + i];
            u32 rank = 0;
            for (u32 k = 0; k < 8; k++)
            {
                if (!(valid_mask & (1u << k))) continue;
                const u32 j = coff + 8u * i + k;
                const u32 child_node = first_child + rank++;
                p_nidx[j] = child_node;
                if (level < 4)
                {
                    p_base[j] = running;  running += p_s[j];
                    p_lbase[j] = lrun;    lrun += p_ls[j];
                }
                else
                {
                    /* leaf: data_pointer = leaf count at visit time (:119) */
                    if (child_node < node_capacity) p_nodes[child_node] = lrun;
                    if (lrun < leaf_capacity) p_scratch[TGB_SCR_DENSE_OF_LEAF + lrun] = 8u * i + k;
                    p_lbase[j] = lrun;
                    lrun++;
                }
            }
            const u32 me = p_nidx[off + i];
            /* :408-410: relative u16 child pointer; leaf_mask set for every valid child of a level-4 node (:447) */
            const u32 word = valid_mask ? (((first_child - me) & 0xFFFFu) | (valid_mask << 16) | ((level == 4 ? valid_mask : 0u) << 24)) : 0u;
            if (me < node_capacity) p_nodes[me] = word;
            if (valid_mask && first_child - me >= 0xFFFFu) atomicOr(&p_counts[3], 1u);
        }
        __syncthreads();
    }
    /* exclusive scan of the pair counts of the dense leaves: 32 consecutive leaves per thread */
    __shared__ u32 s_part[1024];
    {
        const u32 off5 = tgb_level_offset(5);
        u32 sum = 0;
        for (u32 k = 0; k < 32; k++) sum += p_cnt_local[off5 + tid * 32u + k];
        s_part[tid] = sum;
        __syncthreads();
        for (u32 stride = 1; stride < 1024; stride <<= 1)
        {
            const u32 v = tid >= stride ? s_part[tid - stride] : 0u;
            __syncthreads();
            s_part[tid] += v;
            __syncthreads();
        }
        u32 run = s_part[tid] - sum;
        for (u32 k = 0; k < 32; k++)
        {
            p_scratch[TGB_SCR_POFF + tid * 32u + k] = run;
            run += p_cnt_local[off5 + tid * 32u + k];
        }
        if (tid == 1023)
        {
            p_counts[0] = 1u + p_s[0];
            p_counts[1] = p_ls[0];
            p_counts[2] = s_part[1023];
            if (1u + p_s[0] > node_capacity) atomicOr(&p_counts[3], 2u);
            if (p_ls[0] > leaf_capacity) atomicOr(&p_counts[3], 4u);
        }
    }
}

/* ---- pass 4: leaves ---------------------------------------------------------------------------------- */
#define TGB_LEAF_THREADS 256
#define TGB_LEAF_WARPS   (TGB_LEAF_THREADS / 32)

__global__ void __launch_bounds__(TGB_LEAF_THREADS) k_svo_fill_leaves(const u32* __restrict__ p_cluster_pointers, const u32* __restrict__ p_c2o, const tg_object_data* __restrict__ p_objects,
                                                                      const u32* __restrict__ p_masks, v3 bmin, v3 bmax, const u32* __restrict__ p_scratch, const u32* __restrict__ p_pairs,
                                                                      u8* __restrict__ p_pair_flags, u32* __restrict__ p_leaf_data, u32* __restrict__ p_voxels, u32 only_dirty, u32 cluster_idx_base)
{
    __shared__ u32 s_bits[TG_SVO_BLOCK_WORDS];
    __shared__ f32 s_t[3][32]; /* (b + 0.5) / parent_extent per axis, :244-246 */
    __shared__ u32 s_n;

    const u32 data_pointer = blockIdx.x;
    const u32 dense = p_scratch[TGB_SCR_DENSE_OF_LEAF + data_pointer];
    if (only_dirty && p_scratch[TGB_SCR_DIRTY + dense] == 0) return;
    const u32 n_pairs = p_scratch[TGB_SCR_CNT + tgb_level_offset(5) + dense];
    const u32 pair_off = p_scratch[TGB_SCR_POFF + dense];
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    for (u32 i = tid; i < TG_SVO_BLOCK_WORDS; i += TGB_LEAF_THREADS) s_bits[i] = 0;
    if (tid == 0) s_n = 0;
    __syncthreads();

    v3 parent_min, parent_max;
    tgb_dense_box(bmin, bmax, 5, dense, &parent_min, &parent_max);
    const v3 parent_extent = tgb_sub(parent_max, parent_min);
    if (tid < 96u)
    {
        const u32 axis = tid >> 5, b = tid & 31u;
        s_t[axis][b] = ((f32)b + 0.5f) / (axis == 0 ? parent_extent.x : (axis == 1 ? parent_extent.y : parent_extent.z));
    }
    __syncthreads();

    for (u32 k = warp; k < n_pairs; k += TGB_LEAF_WARPS)
    {
        const u32 cluster_pointer = p_pairs[pair_off + k];
        const u32 cluster_idx = __ldg(&p_cluster_pointers[cluster_pointer]);
        const u32 object_idx = __ldg(&p_c2o[cluster_idx]);
        const tg_object_data& o = p_objects[object_idx];
        const u32 rel = cluster_pointer - o.first_cluster_pointer;

        /* :128-194: world AABB of the cluster from corners [0,1,2,2,4,6,5,7] (Q4) */
        const m4 cs2ws = tgb_svo_cs2ws(o, rel);
        const v3 c0 = tgb_m4_transform(cs2ws, tgb_v3(0.0f, 0.0f, 0.0f), 1.0f);
        const v3 c1 = tgb_m4_transform(cs2ws, tgb_v3(8.0f, 0.0f, 0.0f), 1.0f);
        const v3 c2 = tgb_m4_transform(cs2ws, tgb_v3(0.0f, 8.0f, 0.0f), 1.0f);
        const v3 c4 = tgb_m4_transform(cs2ws, tgb_v3(0.0f, 0.0f, 8.0f), 1.0f);
        const v3 c5 = tgb_m4_transform(cs2ws, tgb_v3(8.0f, 0.0f, 8.0f), 1.0f);
        const v3 c6 = tgb_m4_transform(cs2ws, tgb_v3(0.0f, 8.0f, 8.0f), 1.0f);
        const v3 c7 = tgb_m4_transform(cs2ws, tgb_v3(8.0f, 8.0f, 8.0f), 1.0f);
        const v3 min_c = tgb_cmin(tgb_cmin(tgb_cmin(c0, c1), tgb_cmin(c2, c2)), tgb_cmin(tgb_cmin(c4, c6), tgb_cmin(c5, c7)));
        const v3 max_c = tgb_cmax(tgb_cmax(tgb_cmax(c0, c1), tgb_cmax(c2, c2)), tgb_cmax(tgb_cmax(c4, c6), tgb_cmax(c5, c7)));
        const v3 floor_min_c = tgb_v3(floorf(min_c.x), floorf(min_c.y), floorf(min_c.z));
        const v3 ceil_max_c = tgb_v3(ceilf(max_c.x), ceilf(max_c.y), ceilf(max_c.z));

        /* :196-216: the 8 block corners in cluster space */
        const m4 ws2cs = tgb_m4_inverse(cs2ws);
        v3 bc[8];
#pragma unroll
        for (u32 i = 0; i < 8; i++)
        {
            const v3 p = tgb_v3((i & 1u) ? parent_max.x : parent_min.x, (i & 2u) ? parent_max.y : parent_min.y, (i & 4u) ? parent_max.z : parent_min.z);
            bc[i] = tgb_m4_transform(ws2cs, p, 1.0f);
        }

        /* :218-233; a negative difference = the cluster's AABB misses the block: empty range (pinned in oracle/tgo_svo.c) */
        const v3 trimmed_min_b = tgb_cmax(parent_min, floor_min_c);
        const v3 trimmed_max_b = tgb_cmin(parent_max, ceil_max_c);
        const f32 dminx = trimmed_min_b.x - parent_min.x, dminy = trimmed_min_b.y - parent_min.y, dminz = trimmed_min_b.z - parent_min.z;
        const f32 dmaxx = trimmed_max_b.x - parent_min.x, dmaxy = trimmed_max_b.y - parent_min.y, dmaxz = trimmed_max_b.z - parent_min.z;
        bool contributed = false;
        if (!(dminx < 0.0f || dminy < 0.0f || dminz < 0.0f || dmaxx < 0.0f || dmaxy < 0.0f || dmaxz < 0.0f))
        {
            const u32 min_x = (u32)dminx, min_y = (u32)dminy, min_z = (u32)dminz;
            const u32 max_x = (u32)dmaxx, max_y = (u32)dmaxy, max_z = (u32)dmaxz;

            /* the cluster's 64-byte mask: lanes 0..15 hold one word each */
            const u32 my_word = lane < 16u ? __ldg(&p_masks[(u64)cluster_idx * TG_CLUSTER_MASK_WORDS + lane]) : 0u;

            /*
             * :242-315. Per z slice the (by, bx) samples of the clipped range are dealt out to the lanes 32 at a time (the range
             * is about 14 x 14 for a rotated cluster: one lane per bx of a row kept 14 of 32 lanes busy). The interpolation
             * weights t = (b + 0.5) / extent of the three axes come from the CTA's table (the same division, done once).
             * Every lerp keeps the reference's operands and order: z, then y, then x, omt * a + t * b.
             */
            const u32 nxr = max_x > min_x ? max_x - min_x : 0u, nyr = max_y > min_y ? max_y - min_y : 0u, n_xy = nxr * nyr; /* a range may be empty */
            const u32 inv_nxr = nxr ? (65536u + nxr - 1u) / nxr : 0u; /* idx / nxr == (idx * inv_nxr) >> 16 for idx < 1024, nxr <= 32 */
            for (u32 bz = min_z; bz < max_z && n_xy; bz++)
            {
                const f32 tz = s_t[2][bz];
                const f32 omtz = 1.0f - tz;
                const v3 pz0 = tgb_v3(omtz * bc[0].x + tz * bc[4].x, omtz * bc[0].y + tz * bc[4].y, omtz * bc[0].z + tz * bc[4].z);
                const v3 pz1 = tgb_v3(omtz * bc[1].x + tz * bc[5].x, omtz * bc[1].y + tz * bc[5].y, omtz * bc[1].z + tz * bc[5].z);
                const v3 pz2 = tgb_v3(omtz * bc[2].x + tz * bc[6].x, omtz * bc[2].y + tz * bc[6].y, omtz * bc[2].z + tz * bc[6].z);
                const v3 pz3 = tgb_v3(omtz * bc[3].x + tz * bc[7].x, omtz * bc[3].y + tz * bc[7].y, omtz * bc[3].z + tz * bc[7].z);
                for (u32 base = 0; base < n_xy; base += 32u)
                {
                    const u32 idx = base + lane;
                    const bool in_range = idx < n_xy;
                    const u32 iy = in_range ? (idx * inv_nxr) >> 16 : 0u;
                    const u32 by = min_y + iy, bx = min_x + (in_range ? idx - iy * nxr : 0u);
                    const f32 ty = s_t[1][by], tx = s_t[0][bx];
                    const f32 omty = 1.0f - ty, omtx = 1.0f - tx;
                    const v3 py0 = tgb_v3(omty * pz0.x + ty * pz2.x, omty * pz0.y + ty * pz2.y, omty * pz0.z + ty * pz2.z);
                    const v3 py1 = tgb_v3(omty * pz1.x + ty * pz3.x, omty * pz1.y + ty * pz3.y, omty * pz1.z + ty * pz3.z);
                    const f32 cx = omtx * py0.x + tx * py1.x;
                    const f32 cy = omtx * py0.y + tx * py1.y;
                    const f32 cz = omtx * py0.z + tx * py1.z;
                    const bool inside = in_range && !(cx < 0.0f || cx >= 8.0f) && !(cy < 0.0f || cy >= 8.0f) && !(cz < 0.0f || cz >= 8.0f);
                    const u32 rel_voxel = inside ? 64u * (u32)cz + 8u * (u32)cy + (u32)cx : 0u;
                    /* every lane takes part in the shuffle */
                    const u32 word = __shfl_sync(0xFFFFFFFFu, my_word, (int)(rel_voxel >> 5));
                    if (inside && ((word >> (rel_voxel & 31u)) & 1u))
                    {
                        contributed = true;
                        /* block voxel 1024*bz + 32*by + bx: word 32*bz + by, bit bx (parent_extent == 32) */
                        atomicOr(&s_bits[32u * bz + by], 1u << bx);
                    }
                }
            }
        }
        contributed = __any_sync(0xFFFFFFFFu, contributed);
        if (lane == 0) p_pair_flags[pair_off + k] = contributed ? 1 : 0;
    }
    __syncthreads();

    /* coalesced store of the block */
    u32* __restrict__ p_block = p_voxels + (u64)data_pointer * TG_SVO_BLOCK_WORDS;
    for (u32 i = tid; i < TG_SVO_BLOCK_WORDS; i += TGB_LEAF_THREADS) p_block[i] = s_bits[i];

    /* :307-311: cluster indices of the contributing clusters in ascending POINTER order, first 64 kept */
    u32* __restrict__ p_data = p_leaf_data + (u64)data_pointer * 65u;
    for (u32 i = tid; i < 64u; i += TGB_LEAF_THREADS) p_data[1 + i] = 0;
    __syncthreads();
    u32 local_count = 0;
    for (u32 k = tid; k < n_pairs; k += TGB_LEAF_THREADS)
    {
        if (!p_pair_flags[pair_off + k]) continue;
        local_count++;
        const u32 cp = p_pairs[pair_off + k];
        u32 rank = 0;
        for (u32 j = 0; j < n_pairs; j++)
        {
            if (p_pair_flags[pair_off + j] && p_pairs[pair_off + j] < cp) rank++;
        }
        if (rank < TG_SVO_LEAF_MAX_CLUSTERS) p_data[1 + rank] = __ldg(&p_cluster_pointers[cp]) + cluster_idx_base;
    }
    if (local_count) atomicAdd(&s_n, local_count);
    __syncthreads();
    if (tid == 0) p_data[0] = s_n < TG_SVO_LEAF_MAX_CLUSTERS ? s_n : TG_SVO_LEAF_MAX_CLUSTERS;
}

/* ---- incremental update ---------------------------------------------------------------------------------------- */
__global__ void k_svo_set_moved(const u32* __restrict__ p_indices, u32 n, u32 object_capacity, u32* __restrict__ p_moved)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && p_indices[i] < object_capacity) p_moved[p_indices[i]] = 1u;
}

/* a moved object reached this leaf in the PREVIOUS build: its old contribution has to go */
__global__ void k_svo_mark_dirty_prev(const u32* __restrict__ p_pairs_prev, const u32* __restrict__ p_pair_leaf_prev, u32 n_pairs_prev, const u32* __restrict__ p_cluster_pointers,
                                      const u32* __restrict__ p_c2o, const u32* __restrict__ p_moved, u32* __restrict__ p_scratch)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs_prev) return;
    const u32 object_idx = __ldg(&p_c2o[__ldg(&p_cluster_pointers[p_pairs_prev[i]])]);
    if (p_moved[object_idx]) p_scratch[TGB_SCR_DIRTY + p_pair_leaf_prev[i]] = 1u;
}

/* one CTA per leaf of the NEW layout: an untouched leaf that existed before is copied (4 KiB + 260 B), any other is flagged for re-sampling */
__global__ void __launch_bounds__(256) k_svo_copy_clean(u32* __restrict__ p_scratch, const u32* __restrict__ p_leaf_data_old, const u32* __restrict__ p_voxels_old,
                                                        u32* __restrict__ p_leaf_data, u32* __restrict__ p_voxels, u32* __restrict__ p_counts)
{
    const u32 data_pointer = blockIdx.x;
    const u32 dense = p_scratch[TGB_SCR_DENSE_OF_LEAF + data_pointer];
    const u32 prev = p_scratch[TGB_SCR_PREV_DP + dense];
    const bool dirty = p_scratch[TGB_SCR_DIRTY + dense] != 0 || prev == 0;
    __syncthreads();
    if (dirty)
    {
        if (threadIdx.x == 0) { p_scratch[TGB_SCR_DIRTY + dense] = 1u; atomicAdd(&p_counts[4], 1u); }
        return;
    }
    const uint4* p_src = reinterpret_cast<const uint4*>(p_voxels_old + (u64)(prev - 1u) * TG_SVO_BLOCK_WORDS);
    uint4* p_dst = reinterpret_cast<uint4*>(p_voxels + (u64)data_pointer * TG_SVO_BLOCK_WORDS);
    for (u32 i = threadIdx.x; i < TG_SVO_BLOCK_WORDS / 4; i += blockDim.x) p_dst[i] = p_src[i];
    if (threadIdx.x < 65u) p_leaf_data[(u64)data_pointer * 65u + threadIdx.x] = p_leaf_data_old[(u64)(prev - 1u) * 65u + threadIdx.x];
}

/* remember where every dense leaf lives in the arrays just built */
__global__ void k_svo_save_prev(u32* __restrict__ p_scratch, u32 n_leaves)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_leaves) p_scratch[TGB_SCR_PREV_DP + p_scratch[TGB_SCR_DENSE_OF_LEAF + i]] = i + 1u;
}

/* ---- pass 5 (multi-GPU): merge the per-rank leaf contributions ------------------------------------------------ */
/*
 * Each rank sampled only its own clusters. A leaf's voxel bits are an OR over clusters (order-free, S4) and its index
 * list is "ascending pointer order, first 64" -- pointer ranges are contiguous per rank, so that is the ranks' lists
 * concatenated in rank order. NCCL has no bitwise-OR reduction: the partial leaves are all-gathered and combined here.
 * Gathered layout per rank: [n_leaves * 1024 voxel words | n_leaves * 65 leaf-record words].
 */
__global__ void __launch_bounds__(256) k_svo_combine(const u32* __restrict__ p_gather, u32 n_ranks, u32 n_leaves, u32* __restrict__ p_leaf_data, u32* __restrict__ p_voxels)
{
    const u32 leaf = blockIdx.x;
    const u64 per_rank = (u64)n_leaves * (TG_SVO_BLOCK_WORDS + 65u);
    for (u32 i = threadIdx.x; i < TG_SVO_BLOCK_WORDS; i += blockDim.x)
    {
        u32 bits = 0;
        for (u32 r = 0; r < n_ranks; r++) bits |= p_gather[r * per_rank + (u64)leaf * TG_SVO_BLOCK_WORDS + i];
        p_voxels[(u64)leaf * TG_SVO_BLOCK_WORDS + i] = bits;
    }
    if (threadIdx.x < 65u) p_leaf_data[(u64)leaf * 65u + threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        u32 n = 0;
        for (u32 r = 0; r < n_ranks && n < TG_SVO_LEAF_MAX_CLUSTERS; r++)
        {
            const u32* p_rec = p_gather + r * per_rank + (u64)n_leaves * TG_SVO_BLOCK_WORDS + (u64)leaf * 65u;
            const u32 n_r = p_rec[0];
            for (u32 j = 0; j < n_r && n < TG_SVO_LEAF_MAX_CLUSTERS; j++) p_leaf_data[(u64)leaf * 65u + 1u + n++] = p_rec[1 + j];
        }
        p_leaf_data[(u64)leaf * 65u] = n;
    }
}

/*
 * The tree flattened for the stackless secondary-ray kernel (k_gi_trace_flat, tgb_shade.cu). tg_svo_traverse
 * (svo_functions.inc:44-110) finds, from the top of its stack, the TERMINAL node that contains `position`: an octant
 * whose valid bit is clear, or a leaf. Which node that is depends on the position and the comparison rule only, so it
 * can be tabulated: one word per 32^3 cell of the 1024^3 box = depth of the inner node whose child is terminal
 * (child side 512 >> depth), and for a leaf with n != 0 its data pointer. One thread per cell walks the five levels.
 * The last word says whether the table is complete (all leaves at depth 5, no inner node deeper: what
 * tg__construct_inner_node builds, tg_sparse_voxel_octree.c:319-464); otherwise the stack kernel runs instead.
 */
__global__ void k_svo_flatten(const u32* __restrict__ p_nodes, const u32* __restrict__ p_leaf_data, u32 n_nodes, u32 n_leaves, u32* __restrict__ p_grid, unsigned short* __restrict__ p_grid16)
{
    const u32 cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= TGB_TOP_GRID_CELLS) return;
    const u32 cx = cell & 31u, cy = (cell >> 5) & 31u, cz = cell >> 10;
    u32 node = 0, entry = 0;
    bool ok = true, done = false;
    for (u32 level = 0; level < 5u && !done; level++)
    {
        const u32 shift = 4u - level;
        const u32 oct = ((cx >> shift) & 1u) | (((cy >> shift) & 1u) << 1) | (((cz >> shift) & 1u) << 2); /* svo_functions.inc:57-80 */
        const u32 node_data = p_nodes[node];
        const u32 child_pointer = node_data & 0xFFFFu, valid_mask = (node_data >> 16) & 0xFFu, leaf_mask = node_data >> 24;
        entry = level << TGB_TOP_LEVEL_SHIFT;
        if (((valid_mask >> oct) & 1u) == 0) { done = true; break; }
        const u32 child = node + child_pointer + (u32)__popc(valid_mask & ((1u << oct) - 1u)); /* :86-91 */
        if (child >= n_nodes) { ok = false; done = true; break; }
        if ((leaf_mask >> oct) & 1u)
        {
            if (level != 4u) ok = false; /* a leaf larger than 32^3: not a tree the builders make */
            const u32 data_pointer = p_nodes[child];
            if (data_pointer >= n_leaves || data_pointer > TGB_TOP_POINTER_MASK) ok = false;
            else if (p_leaf_data[(u64)data_pointer * 65u] != 0) entry |= TGB_TOP_HAS_DATA | data_pointer;
            done = true;
            break;
        }
        node = child;
    }
    if (!done) ok = false; /* an inner node at depth 5 */
    p_grid[cell] = entry;
    p_grid16[cell] = (unsigned short)tgb_top16_pack(entry); /* the same in 16 bits (a data pointer is a leaf index < 32768) */
    if (!ok) p_grid[TGB_TOP_GRID_CELLS] = 0;
}

static b32 tgbd__svo_flatten(struct tgb_device* d, cudaStream_t st)
{
    tgb_svo_device* s = &d->svo;
    TGB_CUDA(cudaMemsetAsync(s->d_top_grid + TGB_TOP_GRID_CELLS, 1, sizeof(u32), st)); /* non-zero = complete, cleared by the kernel */
    k_svo_flatten<<<TGB_TOP_GRID_CELLS / 256, 256, 0, st>>>(s->d_nodes, s->d_leaf_data, s->n_nodes, s->n_leaves, s->d_top_grid, (unsigned short*)(s->d_top_grid + TGB_TOP_GRID_CELLS + 1));
    TGB_LAUNCH_CHECK(d);
    /* the certified fast walk's coarser tiling of the same tree (tgb_gi_fast.cu): here, on the build's stream, when that kernel is selected;
     * otherwise the first trace that needs it builds it */
    s->fast_tiling_valid = TG_FALSE;
    if (tgbd_env_int("TGB_GI_KERNEL", TGB_GI_KERNEL_DEFAULT) == 4 && !tgbd_gi_fast_tiling_build(d, st)) return TG_FALSE;
    return TG_TRUE;
}

/* ---- host side of the seam ---------------------------------------------------------------------------- */
static b32 tgbd__svo_ensure(struct tgb_device* d)
{
    tgb_svo_device* s = &d->svo;
    if (!s->d_scratch)
    {
        TGB_CUDA(cudaMalloc(&s->d_scratch, (u64)TGB_SCR_TOTAL * sizeof(u32)));
        s->scratch_capacity = TGB_SCR_TOTAL;
        TGB_CUDA(cudaMalloc(&s->d_object_flags, (u64)d->object_capacity * sizeof(u32)));
        TGB_CUDA(cudaMalloc(&s->d_moved_indices, (u64)d->object_capacity * sizeof(u32)));
    }
    return TG_TRUE;
}

/* the pair lists of the build about to run go to half "a"; an incremental update first moves the previous lists to half "b" */
static b32 tgbd__svo_ensure_pairs(struct tgb_device* d, u64 n_pairs)
{
    tgb_svo_device* s = &d->svo;
    if (n_pairs <= s->pair_capacity) return TG_TRUE;
    u64 cap = s->pair_capacity ? s->pair_capacity : (1u << 16);
    while (cap < n_pairs) cap *= 2;
    if (s->d_pairs_a) TGB_CUDA(cudaFree(s->d_pairs_a));
    if (s->d_pair_leaf_a) TGB_CUDA(cudaFree(s->d_pair_leaf_a));
    s->d_pairs_a = NULL; s->d_pair_leaf_a = NULL; s->pair_capacity = 0;
    TGB_CUDA(cudaMalloc(&s->d_pairs_a, cap * sizeof(u32)));
    TGB_CUDA(cudaMalloc(&s->d_pair_leaf_a, cap * sizeof(u32)));
    s->pair_capacity = cap;
    return TG_TRUE;
}

static b32 tgbd__svo_ensure_pair_flags(struct tgb_device* d, u64 n_pairs)
{
    tgb_svo_device* s = &d->svo;
    if (n_pairs <= s->pair_flags_capacity) return TG_TRUE;
    u64 cap = s->pair_flags_capacity ? s->pair_flags_capacity : (1u << 16);
    while (cap < n_pairs) cap *= 2;
    if (s->d_pair_flags) TGB_CUDA(cudaFree(s->d_pair_flags));
    s->d_pair_flags = NULL; s->pair_flags_capacity = 0;
    TGB_CUDA(cudaMalloc(&s->d_pair_flags, cap));
    s->pair_flags_capacity = cap;
    return TG_TRUE;
}

static b32 tgbd__svo_ensure_gather(struct tgb_device* d, u32 n_leaves)
{
    tgb_svo_device* s = &d->svo;
    if (n_leaves <= s->part_capacity_leaves) return TG_TRUE;
    u64 cap = s->part_capacity_leaves ? s->part_capacity_leaves : 1024;
    while (cap < n_leaves) cap *= 2;
    if (s->d_part) TGB_CUDA(cudaFree(s->d_part));
    if (s->d_gather) TGB_CUDA(cudaFree(s->d_gather));
    s->d_part = NULL; s->d_gather = NULL; s->part_capacity_leaves = 0;
    const u64 words = cap * (TG_SVO_BLOCK_WORDS + 65u);
    TGB_CUDA(cudaMalloc(&s->d_part, words * sizeof(u32)));
    TGB_CUDA(cudaMalloc(&s->d_gather, words * sizeof(u32) * d->n_ranks));
    s->part_capacity_leaves = cap;
    return TG_TRUE;
}

/*
 * Full build. On one GPU: passes 1-4. Sharded (communicator set): every rank walks its own clusters, the arrival counts
 * are summed over the ranks (ncclAllReduce, 150 KB) so that every rank lays out the SAME tree, each rank samples its
 * clusters into a partial copy of every leaf, the partial leaves are all-gathered and OR-combined (pass 5). The result
 * is bit-identical on every rank and to a single-GPU build over the union of the shards. Collective: all ranks call it.
 *
 * Incremental update (single GPU, only transforms changed since the previous K2 build): the per-cluster walk and the
 * layout run in full (they are cheap and keep the node array canonical), but only the leaves a moved object reaches now
 * or reached in the previous build are re-sampled; every other leaf is copied from the previous arrays to its new
 * data_pointer. The three arrays equal a full rebuild bit for bit (tests/test_svo_gpu.py).
 */
static b32 tgbd__svo_run(struct tgb_device* d, v3 extent_min, v3 extent_max, u32 n_cluster_pointers, u32 object_capacity, bool incremental, u32 n_moved, const u32* p_moved_indices)
{
    TGB_CUDA(cudaSetDevice(d->device));
    if (!tgbd_flush_objects(d)) return TG_FALSE; /* on the main stream, in front of the event the build waits for */
    if (!tgbd__svo_ensure(d)) return TG_FALSE;
    tgb_svo_device* s = &d->svo;
    const bool sharded = d->p_comm != NULL && d->n_ranks > 1;
    incremental = incremental && !sharded && s->valid && s->incremental_ok
               && s->bmin.x == extent_min.x && s->bmin.y == extent_min.y && s->bmin.z == extent_min.z
               && s->bmax.x == extent_max.x && s->bmax.y == extent_max.y && s->bmax.z == extent_max.z;
    const u32 n_pairs_prev = s->n_pairs;
    s->valid = TG_FALSE;
    s->bmin = extent_min; s->bmax = extent_max;

    /*
     * K3 is the only consumer of the tree, K1 neither reads nor writes it: on one GPU the build runs on its OWN stream, next to the K1
     * the caller has already queued on the main stream (tg_raytracer_render and bench.py call K1 first). It starts when the frame's
     * inputs are complete -- the event tgbd_clear records after the application's uploads; an upload after the clear falls back to "now"
     * -- which also orders it behind the previous frame's shading (the reader of the arrays it rewrites), and the main stream waits
     * for its end. The one-CTA layout pass, the host round trip and the short passes then hide behind K1. (Sharded builds keep the main
     * stream: their collectives are ordered with the frame's.)
     */
    cudaStream_t st = d->stream;
    if (!sharded && d->svo_stream && tgbd_env_int("TGB_SVO_STREAM", 1))
    {
        st = d->svo_stream;
        if (d->inputs_changed_since_clear || !d->ev_inputs_valid) { TGB_CUDA(cudaEventRecord(d->ev_inputs, d->stream)); d->ev_inputs_valid = TG_TRUE; d->inputs_changed_since_clear = TG_FALSE; }
        TGB_CUDA(cudaStreamWaitEvent(st, d->ev_inputs, 0));
    }

    TGB_CUDA(cudaEventRecord(d->ev[5], st));
    if (incremental)
    {
        /* previous pair lists -> half "b" (pointer swap), spare leaf arrays on demand */
        u32* t;
        t = s->d_pairs_a; s->d_pairs_a = s->d_pairs_b; s->d_pairs_b = t;
        t = s->d_pair_leaf_a; s->d_pair_leaf_a = s->d_pair_leaf_b; s->d_pair_leaf_b = t;
        const u64 c = s->pair_capacity; s->pair_capacity = s->pair_capacity_b; s->pair_capacity_b = c;
        if (!s->d_voxels_alt)
        {
            TGB_CUDA(cudaMalloc(&s->d_voxels_alt, (u64)s->voxel_word_capacity * sizeof(u32)));
            TGB_CUDA(cudaMalloc(&s->d_leaf_data_alt, (u64)s->leaf_capacity * 65 * sizeof(u32)));
        }
        TGB_CUDA(cudaMemsetAsync(s->d_object_moved, 0, (u64)object_capacity * sizeof(u32), st));
        if (n_moved > object_capacity) n_moved = object_capacity;
        if (n_moved)
        {
            TGB_CUDA(cudaMemcpyAsync(s->d_moved_indices, p_moved_indices, (u64)n_moved * sizeof(u32), cudaMemcpyHostToDevice, st));
            k_svo_set_moved<<<(n_moved + 127) / 128, 128, 0, st>>>(s->d_moved_indices, n_moved, object_capacity, s->d_object_moved);
            TGB_LAUNCH_CHECK(d);
        }
    }
    TGB_CUDA(cudaMemsetAsync(s->d_scratch, 0, (u64)TGB_SCR_CLEARED * sizeof(u32), st));
    if (!incremental) TGB_CUDA(cudaMemsetAsync(s->d_scratch + TGB_SCR_PREV_DP, 0, (u64)TGB_SVO_MAX_LEAVES * sizeof(u32), st));
    TGB_CUDA(cudaMemsetAsync(s->d_counts, 0, 16 * sizeof(u32), st));
    TGB_CUDA(cudaMemsetAsync(s->d_nodes, 0, (u64)s->node_capacity * sizeof(u32), st));

    k_svo_object_flags<<<(object_capacity + 127) / 128, 128, 0, st>>>(d->d_objects, object_capacity, extent_min, extent_max, s->d_object_flags);
    TGB_LAUNCH_CHECK(d);
    const u32 grid = (n_cluster_pointers + 127) / 128;
    if (grid)
    {
        k_svo_descend<false><<<grid, 128, 0, st>>>(d->d_cluster_pointers, d->d_c2o, d->d_objects, s->d_object_flags, d->d_masks, n_cluster_pointers,
                                                         extent_min, extent_max, s->d_scratch, NULL, NULL, incremental ? s->d_object_moved : NULL);
        TGB_LAUNCH_CHECK(d);
    }
    if (incremental && n_pairs_prev)
    {
        k_svo_mark_dirty_prev<<<(n_pairs_prev + 255) / 256, 256, 0, st>>>(s->d_pairs_b, s->d_pair_leaf_b, n_pairs_prev, d->d_cluster_pointers, d->d_c2o,
                                                                                s->d_object_moved, s->d_scratch);
        TGB_LAUNCH_CHECK(d);
    }
    TGB_CUDA(cudaMemcpyAsync(s->d_scratch + TGB_SCR_CNT_G, s->d_scratch + TGB_SCR_CNT, (u64)TGB_SVO_DENSE_TOTAL * sizeof(u32), cudaMemcpyDeviceToDevice, st));
    if (sharded && !tgbn_allreduce_sum_u32(d->p_comm, s->d_scratch + TGB_SCR_CNT_G, TGB_SVO_DENSE_TOTAL, st)) return TG_FALSE;
    k_svo_layout<<<1, 1024, 0, st>>>(s->d_scratch, s->d_nodes, s->d_counts, s->node_capacity, s->leaf_capacity);
    TGB_LAUNCH_CHECK(d);

    /* the one host round trip of a build: node / leaf / pair counts size the remaining launches */
    u32 counts[4] = { 0, 0, 0, 0 };
    TGB_CUDA(cudaMemcpyAsync(counts, s->d_counts, sizeof(counts), cudaMemcpyDeviceToHost, st));
    TGB_CUDA(cudaStreamSynchronize(st));
    if (counts[3])
    {
        tgb_set_error("svo build: capacity exceeded (flags %u: 1 = child pointer >= 0xFFFF, 2 = nodes %u > %u, 4 = leaves %u > %u; tg_sparse_voxel_octree.c:116,408,413)",
                      counts[3], counts[0], s->node_capacity, counts[1], s->leaf_capacity);
        s->incremental_ok = TG_FALSE;
        return TG_FALSE;
    }
    s->n_nodes = counts[0];
    s->n_leaves = counts[1];
    if (!tgbd__svo_ensure_pairs(d, counts[2] ? counts[2] : 1) || !tgbd__svo_ensure_pair_flags(d, counts[2] ? counts[2] : 1)) return TG_FALSE;
    if (sharded && !tgbd__svo_ensure_gather(d, s->n_leaves)) return TG_FALSE;

    if (counts[2] && grid)
    {
        k_svo_descend<true><<<grid, 128, 0, st>>>(d->d_cluster_pointers, d->d_c2o, d->d_objects, s->d_object_flags, d->d_masks, n_cluster_pointers,
                                                        extent_min, extent_max, s->d_scratch, s->d_pairs_a, s->d_pair_leaf_a, NULL);
        TGB_LAUNCH_CHECK(d);
    }
    s->n_leaves_resampled = s->n_leaves;
    if (s->n_leaves)
    {
        u32* p_voxels = sharded ? s->d_part : (incremental ? s->d_voxels_alt : s->d_voxels);
        u32* p_leaf = sharded ? s->d_part + (u64)s->n_leaves * TG_SVO_BLOCK_WORDS : (incremental ? s->d_leaf_data_alt : s->d_leaf_data);
        if (incremental)
        {
            k_svo_copy_clean<<<s->n_leaves, 256, 0, st>>>(s->d_scratch, s->d_leaf_data, s->d_voxels, p_leaf, p_voxels, s->d_counts);
            TGB_LAUNCH_CHECK(d);
        }
        k_svo_fill_leaves<<<s->n_leaves, TGB_LEAF_THREADS, 0, st>>>(d->d_cluster_pointers, d->d_c2o, d->d_objects, d->d_masks, extent_min, extent_max,
                                                                           s->d_scratch, s->d_pairs_a, s->d_pair_flags, p_leaf, p_voxels, incremental ? 1u : 0u,
                                                                           sharded ? d->global_pointer_base : 0u);
        TGB_LAUNCH_CHECK(d);
        if (sharded)
        {
            const u64 part_bytes = (u64)s->n_leaves * (TG_SVO_BLOCK_WORDS + 65u) * sizeof(u32);
            if (!tgbn_allgather_bytes(d->p_comm, s->d_part, s->d_gather, part_bytes, st)) return TG_FALSE;
            k_svo_combine<<<s->n_leaves, 256, 0, st>>>(s->d_gather, d->n_ranks, s->n_leaves, s->d_leaf_data, s->d_voxels);
            TGB_LAUNCH_CHECK(d);
        }
        if (incremental)
        {
            u32* t;
            t = s->d_voxels; s->d_voxels = s->d_voxels_alt; s->d_voxels_alt = t;
            t = s->d_leaf_data; s->d_leaf_data = s->d_leaf_data_alt; s->d_leaf_data_alt = t;
        }
    }
    /* where every dense leaf lives now (for the next incremental update) */
    TGB_CUDA(cudaMemsetAsync(s->d_scratch + TGB_SCR_PREV_DP, 0, (u64)TGB_SVO_MAX_LEAVES * sizeof(u32), st));
    if (s->n_leaves)
    {
        k_svo_save_prev<<<(s->n_leaves + 255) / 256, 256, 0, st>>>(s->d_scratch, s->n_leaves);
        TGB_LAUNCH_CHECK(d);
    }
    if (incremental) TGB_CUDA(cudaMemcpyAsync(&s->n_leaves_resampled, s->d_counts + 4, sizeof(u32), cudaMemcpyDeviceToHost, st));
    if (!tgbd__svo_flatten(d, st)) return TG_FALSE;
    TGB_CUDA(cudaEventRecord(d->ev[6], st));
    if (st != d->stream) TGB_CUDA(cudaStreamWaitEvent(d->stream, d->ev[6], 0)); /* whatever follows on the main stream (K3, downloads) sees the finished tree */
    d->ev_svo = TG_TRUE;
    s->n_pairs = counts[2];
    s->valid = TG_TRUE;
    s->incremental_ok = !sharded;
    return TG_TRUE;
}

extern "C" b32 tgbd_svo_build(struct tgb_device* d, v3 extent_min, v3 extent_max, u32 n_cluster_pointers, u32 object_capacity)
{
    return tgbd__svo_run(d, extent_min, extent_max, n_cluster_pointers, object_capacity, false, 0, NULL);
}

extern "C" b32 tgbd_svo_update_objects(struct tgb_device* d, v3 extent_min, v3 extent_max, u32 n_cluster_pointers, u32 object_capacity, u32 n_moved, const u32* p_object_indices)
{
    return tgbd__svo_run(d, extent_min, extent_max, n_cluster_pointers, object_capacity, true, n_moved, p_object_indices);
}

extern "C" void tgbd_svo_invalidate_incremental(struct tgb_device* d) { d->svo.incremental_ok = TG_FALSE; }
extern "C" u32 tgbd_svo_leaves_resampled(struct tgb_device* d) { cudaStreamSynchronize(d->stream); return d->svo.n_leaves_resampled; }

extern "C" b32 tgbd_svo_counts(struct tgb_device* d, u32* p_n_nodes, u32* p_n_leaves, u32* p_n_voxel_words, v3* p_min, v3* p_max)
{
    if (!d->svo.valid) { tgb_set_error("no SVO has been built or uploaded"); return TG_FALSE; }
    *p_n_nodes = d->svo.n_nodes;
    *p_n_leaves = d->svo.n_leaves;
    *p_n_voxel_words = d->svo.n_leaves * TG_SVO_BLOCK_WORDS;
    *p_min = d->svo.bmin;
    *p_max = d->svo.bmax;
    return TG_TRUE;
}

extern "C" b32 tgbd_svo_set(struct tgb_device* d, v3 bmin, v3 bmax, u32 n_nodes, const void* p_nodes, u32 n_leaves, const void* p_leaf_data, u32 n_voxel_words, const void* p_voxels)
{
    TGB_CUDA(cudaSetDevice(d->device));
    tgb_svo_device* s = &d->svo;
    if (n_nodes > s->node_capacity || n_leaves > s->leaf_capacity || n_voxel_words > s->voxel_word_capacity)
    {
        tgb_set_error("svo upload: %u nodes / %u leaves / %u voxel words exceed the device capacities %u / %u / %u", n_nodes, n_leaves, n_voxel_words,
                      s->node_capacity, s->leaf_capacity, s->voxel_word_capacity);
        return TG_FALSE;
    }
    if (bmax.x - bmin.x != (f32)TG_SVO_SIDE_LENGTH || bmax.y - bmin.y != (f32)TG_SVO_SIDE_LENGTH || bmax.z - bmin.z != (f32)TG_SVO_SIDE_LENGTH)
    {
        tgb_set_error("svo upload: the box must be %d^3 (32^3 leaf blocks, tg_sparse_voxel_octree.c:468-472)", TG_SVO_SIDE_LENGTH);
        return TG_FALSE;
    }
    s->valid = TG_FALSE;
    if (n_nodes)       TGB_CUDA(cudaMemcpyAsync(s->d_nodes, p_nodes, (u64)n_nodes * 4, cudaMemcpyHostToDevice, d->stream));
    if (n_leaves)      TGB_CUDA(cudaMemcpyAsync(s->d_leaf_data, p_leaf_data, (u64)n_leaves * 260, cudaMemcpyHostToDevice, d->stream));
    if (n_voxel_words) TGB_CUDA(cudaMemcpyAsync(s->d_voxels, p_voxels, (u64)n_voxel_words * 4, cudaMemcpyHostToDevice, d->stream));
    TGB_CUDA(cudaStreamSynchronize(d->stream));
    s->bmin = bmin; s->bmax = bmax;
    s->n_nodes = n_nodes; s->n_leaves = n_leaves;
    if (!n_nodes) TGB_CUDA(cudaMemsetAsync(s->d_nodes, 0, 4, d->stream)); /* an empty upload is an empty root */
    if (!tgbd__svo_flatten(d, d->stream)) return TG_FALSE;
    s->valid = TG_TRUE;
    s->incremental_ok = TG_FALSE; /* uploaded arrays: there are no pair lists to update from */
    return TG_TRUE;
}

/*
 * tgb_gi_fast.cu -- K3b, first pass: the queued secondary rays decided by the certified fast walk (tgb_gi_fast.cuh).
 *
 *   SVO traversal      assets/shaders/raytracer/svo_functions.inc:1-329
 *   secondary rays     tgvk_raytracer.c:1405-1431, TODO.h:33-43 (queued by k_shade, tgb_shade.cu)
 *
 * One ray per lane, everything in registers, ONE kind of step for every cell of the SVO (empty terminal box or voxel of a leaf
 * block): all lanes that hold a ray run the same instructions, where the exact kernels split every warp into a tree camp and a
 * DDA camp. Decisions the walk can certify (tgb_gi_fast.cuh: every ray displaced sideways by less than DELTA decides the same)
 * are final: occluded rays keep the radiance k_shade wrote, unoccluded rays add their ambient term. The others -- a few per
 * cent -- are appended to a list of queue slots, and k_gi_trace_pool (tgb_gi_pool.cu: the shader's own arithmetic) traces exactly
 * those right after this kernel. The frame is therefore the one the exact kernel alone produces (tests: radiance bit-identical
 * between TGB_GI_KERNEL=2 and 3 on every pixel of the 4K frame).
 *
 * Persistent: lanes whose ray is decided wait until `service_lanes` of the warp do, then the warp returns ambient terms, hands
 * uncertain rays over and fetches new rays together.
 */
#include "tgb_device.cuh"
#include "tgb_gi_fast.cuh"

#define TGB_FAST_THREADS 128
#define TGB_FAST_WARPS   (TGB_FAST_THREADS / 32)
#define TGB_FAST_CHUNK   64u    /* queue slots a warp reserves with one atomic */

/* 16 bytes global -> shared without passing through registers (LDGSTS; L2 only: a queue record is read once) */
__device__ __forceinline__ void tgb_cp_async16(void* p_shared, const void* p_global)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((u32)__cvta_generic_to_shared(p_shared)), "l"(p_global) : "memory");
}
__device__ __forceinline__ void tgb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tgb_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

/* words of a ready ray in the warp's pool */
enum { P_OBX = 0, P_OBY, P_OBZ, P_DX, P_DY, P_DZ, P_IX, P_IY, P_IZ, P_W, P_T, P_VOX, P_SLOT, P_PIXEL, P_AX, P_AY, P_AZ, P_WORDS };
#define TGB_FAST_POOL_EXACT 0x80000000u /* P_VOX: the ray is not taken by the fast walk (shallow direction) */
#define TGB_FAST_POOL_FIRST 0x40000000u /* P_VOX: TGB_FAST_FIRST */

/*
 * How a ray reaches a lane. Queue slots are reserved per warp in chunks (one atomic per TGB_FAST_CHUNK rays). The records of the next
 * 32 slots (origin + pixel | direction + root enter | ambient: 3 x 16 B each) are fetched by cp.async, one record per lane, while
 * the warp walks; when the warp's pool of ready rays is empty all 32 lanes set their record up TOGETHER (reciprocals, W, start
 * voxel: tgb_fast_start) and file the ready rays in shared memory, one column per word. A lane whose ray is decided only copies
 * a ready ray into its registers: a service costs neither memory latency nor a set-up run by a quarter of the warp.
 */
/* CUBE: the careful second pass over the slots the first pass handed over (p_in_list, counted in p_q_count[in_count_word], fetched through
 * p_q_count[in_count_word + 1]): the same walk with the cube check of near-edge steps (tgb_fast_cube_free) and generous step caps; what it
 * hands over in turn (p_exact_list, counted in p_q_count[out_count_word]) goes to the shader's own arithmetic (k_gi_trace_list). */
template <bool TILED, bool CUBE>
__global__ void __launch_bounds__(TGB_FAST_THREADS, (TILED && !CUBE) ? 8 : 1) k_gi_trace_fast(const tgb_gi_frame fr, const tgb_fast_tiling tiling, const u32* __restrict__ p_in_list, u32 in_count_word, u32 out_count_word, const float4* __restrict__ p_q0, const float4* __restrict__ p_q1,
                                                                    const float4* __restrict__ p_q2, u32* __restrict__ p_q_count, u32* __restrict__ p_exact_list,
                                                                    float4* __restrict__ p_out, u32 service_lanes, u32 steps, u32 max_steps, u32 max_steps_uncertain, f32 delta)
{
    if (fr.p_grid[TGB_TOP_GRID_CELLS] == 0) return; /* not tabulated: k_gi_trace runs */

    __shared__ float4 s_rec[3][TGB_FAST_THREADS];                 /* staged records, one per lane */
    __shared__ u32 s_pool[TGB_FAST_WARPS][P_WORDS][32];           /* ready rays of the warp */
    __shared__ u32 s_cur[5][TGB_FAST_THREADS];                    /* of the ray a lane walks: queue slot, pixel, ambient */
    __shared__ u32 s_slot[CUBE ? TGB_FAST_THREADS : 1];           /* CUBE: queue slots of the staged records */
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u32 n_rays = p_q_count[CUBE ? in_count_word : 0u];
    if (!CUBE && blockIdx.x == 0 && tid == 0) atomicAdd(&p_q_count[10], n_rays); /* rays of the frame, summed over its bands */
    if (CUBE && (u64)blockIdx.x * TGB_FAST_THREADS >= (u64)n_rays + TGB_FAST_THREADS) return; /* more CTAs than the list can feed */

    tgb_fast_ray r;
    r.ob = r.d = r.inv = r.r = r.posf = tgb_v3(0.0f, 0.0f, 0.0f);
    r.w = r.w_step = r.w_t = r.t_cur = 0.0f;
    r.vx = r.vy = r.vz = 0;
    r.cell = r.entry = r.flags = r.n_steps = 0;
    u32 kind = TGB_FAST_IDLE;
    /* warp-uniform */
    u32 pool_head = 0, pool_n = 0;       /* pool entries [pool_head, pool_n) hold a ready ray (filed by lanes 0 .. n - 1, taken in order) */
    u32 staged_n = 0;                    /* records in flight / arrived in s_rec (lanes 0 .. staged_n - 1) */
    u32 staged_base = 0;
    u32 c_next = 0, c_end = 0;           /* the warp's reserved chunk of the queue */
    bool drained = false;                /* no more chunks */
    u32 n_cells = 0, n_exact = 0;

    for (;;)
    {
        /* one REDUX counts the lanes that walk and the lanes whose ray is decided */
        const u32 counts = __reduce_add_sync(0xFFFFFFFFu, kind == TGB_FAST_WALK ? 1u : (kind == TGB_FAST_IDLE ? 0u : 0x100u));
        const u32 n_walking = counts & 0xFFu, n_decided = counts >> 8, n_idle = 32u - n_walking - n_decided;
        const bool more = pool_head < pool_n || staged_n != 0u || !drained || c_next < c_end;
        if (n_walking == 0 && n_decided == 0 && !more) break; /* queue drained, every ray decided and served */

        if (n_decided + (more ? n_idle : 0u) >= service_lanes || n_walking == 0)
        {
            /* ---- service: certain unoccluded rays return their ambient term, uncertain ones are handed over ---- */
            if (kind == TGB_FAST_UNOCCLUDED)
            {
                if (r.flags & TGB_FAST_UNCERTAIN) kind = TGB_FAST_EXACT;
                else
                {
                    f32* p_pixel = reinterpret_cast<f32*>(&p_out[s_cur[1][tid]]);
                    atomicAdd(p_pixel + 0, __uint_as_float(s_cur[2][tid]));
                    atomicAdd(p_pixel + 1, __uint_as_float(s_cur[3][tid]));
                    atomicAdd(p_pixel + 2, __uint_as_float(s_cur[4][tid]));
                    kind = TGB_FAST_IDLE;
                }
            }
            else if (kind == TGB_FAST_OCCLUDED) kind = TGB_FAST_IDLE;
            if (kind != TGB_FAST_WALK) { n_cells += r.n_steps; r.n_steps = 0; } /* cells the decided ray entered */
            for (u32 round = 0; round < 2u; round++)
            {
                /* hand-over: the slot goes to the list k_gi_trace_pool reads (warp-aggregated append); second round: shallow rays just taken */
                const u32 handed = __ballot_sync(0xFFFFFFFFu, kind == TGB_FAST_EXACT);
                if (handed)
                {
                    u32 base = 0;
                    const u32 leader = (u32)(__ffs(handed) - 1);
                    if (lane == leader) base = atomicAdd(&p_q_count[out_count_word], (u32)__popc(handed));
                    base = __shfl_sync(0xFFFFFFFFu, base, (int)leader);
                    if (kind == TGB_FAST_EXACT)
                    {
                        p_exact_list[base + (u32)__popc(handed & ((1u << lane) - 1u))] = s_cur[0][tid];
                        n_exact++;
                        kind = TGB_FAST_IDLE;
                    }
                }
                if (round == 1u) break;

                /* ---- the pool is empty: the staged records become ready rays, all lanes together ---- */
                if (pool_head >= pool_n)
                {
                    for (u32 attempt = 0; attempt < 2u && pool_head >= pool_n; attempt++)
                    {
                        if (staged_n != 0u)
                        {
                            tgb_cp_async_wait_all();
                            if (lane < staged_n)
                            {
                                const float4 q0 = s_rec[0][tid], q1 = s_rec[1][tid], q2 = s_rec[2][tid];
                                tgb_fast_ray n;
                                const u32 k0 = tgb_fast_start(&fr, tgb_v3(q0.x, q0.y, q0.z), tgb_v3(q1.x, q1.y, q1.z), q1.w, delta, &n, TILED);
                                u32* p = &s_pool[warp][0][lane];
                                p[P_OBX * 32] = __float_as_uint(n.ob.x); p[P_OBY * 32] = __float_as_uint(n.ob.y); p[P_OBZ * 32] = __float_as_uint(n.ob.z);
                                p[P_DX * 32] = __float_as_uint(q1.x); p[P_DY * 32] = __float_as_uint(q1.y); p[P_DZ * 32] = __float_as_uint(q1.z);
                                p[P_IX * 32] = __float_as_uint(n.inv.x); p[P_IY * 32] = __float_as_uint(n.inv.y); p[P_IZ * 32] = __float_as_uint(n.inv.z);
                                p[P_W * 32] = __float_as_uint(n.w); p[P_T * 32] = __float_as_uint(n.t_cur);
                                p[P_VOX * 32] = k0 == TGB_FAST_EXACT ? TGB_FAST_POOL_EXACT
                                                                      : ((u32)n.vx | ((u32)n.vy << 10) | ((u32)n.vz << 20) | ((n.flags & TGB_FAST_FIRST) ? TGB_FAST_POOL_FIRST : 0u));
                                p[P_SLOT * 32] = CUBE ? s_slot[tid] : staged_base + lane; p[P_PIXEL * 32] = __float_as_uint(q0.w);
                                p[P_AX * 32] = __float_as_uint(q2.x); p[P_AY * 32] = __float_as_uint(q2.y); p[P_AZ * 32] = __float_as_uint(q2.z);
                            }
                            __syncwarp();
                            pool_head = 0; pool_n = staged_n;
                            staged_n = 0;
                        }
                        /* stage the next 32 records of the warp's chunk (a new chunk when this one is used up) */
                        if (c_next >= c_end && !drained)
                        {
                            u32 nb = 0;
                            if (lane == 0) nb = atomicAdd(&p_q_count[CUBE ? in_count_word + 1u : 1u], CUBE ? 32u : TGB_FAST_CHUNK);
                            nb = __shfl_sync(0xFFFFFFFFu, nb, 0);
                            const u32 chunk = CUBE ? 32u : TGB_FAST_CHUNK; /* the few rays of a list are spread over the warps */
                            c_next = nb < n_rays ? nb : n_rays;
                            c_end = nb + chunk < n_rays ? nb + chunk : n_rays;
                            drained = nb + chunk >= n_rays;
                        }
                        if (c_next < c_end)
                        {
                            staged_n = c_end - c_next < 32u ? c_end - c_next : 32u;
                            staged_base = c_next;
                            c_next += staged_n;
                            if (lane < staged_n)
                            {
                                const u32 slot = CUBE ? p_in_list[staged_base + lane] : staged_base + lane;
                                if (CUBE) s_slot[tid] = slot;
                                tgb_cp_async16(&s_rec[0][tid], &p_q0[slot]);
                                tgb_cp_async16(&s_rec[1][tid], &p_q1[slot]);
                                tgb_cp_async16(&s_rec[2][tid], &p_q2[slot]);
                            }
                            tgb_cp_async_commit();
                        }
                    }
                }
                /* ---- idle lanes take a ready ray ---- */
                {
                    const u32 idle = __ballot_sync(0xFFFFFFFFu, kind == TGB_FAST_IDLE);
                    const u32 e = pool_head + (u32)__popc(idle & ((1u << lane) - 1u));   /* the idle lanes take the entries in order */
                    if (kind == TGB_FAST_IDLE && e < pool_n)
                    {
                        const u32* p = &s_pool[warp][0][e];
                        const u32 vox = p[P_VOX * 32];
                        r.ob = tgb_v3(__uint_as_float(p[P_OBX * 32]), __uint_as_float(p[P_OBY * 32]), __uint_as_float(p[P_OBZ * 32]));
                        r.d = tgb_v3(__uint_as_float(p[P_DX * 32]), __uint_as_float(p[P_DY * 32]), __uint_as_float(p[P_DZ * 32]));
                        r.inv = tgb_v3(__uint_as_float(p[P_IX * 32]), __uint_as_float(p[P_IY * 32]), __uint_as_float(p[P_IZ * 32]));
                        tgb_fast_derive(&r, __uint_as_float(p[P_W * 32]));
                        r.t_cur = __uint_as_float(p[P_T * 32]);
                        r.vx = (i32)(vox & 1023u); r.vy = (i32)((vox >> 10) & 1023u); r.vz = (i32)((vox >> 20) & 1023u);
                        r.cell = 0xFFFFFFFFu; r.entry = 0; r.n_steps = 0;
                        r.flags = (vox & TGB_FAST_POOL_FIRST) ? TGB_FAST_FIRST : 0u;
                        s_cur[0][tid] = p[P_SLOT * 32]; s_cur[1][tid] = p[P_PIXEL * 32];
                        s_cur[2][tid] = p[P_AX * 32]; s_cur[3][tid] = p[P_AY * 32]; s_cur[4][tid] = p[P_AZ * 32];
                        kind = (vox & TGB_FAST_POOL_EXACT) ? TGB_FAST_EXACT : TGB_FAST_WALK;
                    }
                    pool_head = pool_head + (u32)__popc(idle) < pool_n ? pool_head + (u32)__popc(idle) : pool_n;
                    __syncwarp();
                }
            }
            continue;
        }

        if (kind == TGB_FAST_WALK)
            kind = TILED ? tgb_fast_walk_tiled<CUBE>(&fr, &tiling, &r, steps, (u32*)0, (u32*)0, max_steps, max_steps_uncertain)
                         : tgb_fast_walk(&fr, &r, steps, (u32*)0, (u32*)0, max_steps, max_steps_uncertain);
    }
    /* [2] cells (empty boxes and voxels) entered by the fast walk in this frame; the exact kernel adds its look-ups there and counts its DDA steps in [3]; [15] rays handed over ([14], rays that needed the exact walk, is counted by the second stage) */
    n_cells = __reduce_add_sync(0xFFFFFFFFu, n_cells);
    n_exact = __reduce_add_sync(0xFFFFFFFFu, n_exact);
    if (lane == 0)
    {
        atomicAdd(reinterpret_cast<unsigned long long*>(p_q_count) + 1, (unsigned long long)n_cells);
        if (!CUBE) atomicAdd(&p_q_count[15], n_exact);
    }
}

/* ---- the coarser tiling (tgb_gi_fast.cuh, second half), rebuilt after every SVO build ---------------------------------------------- */

struct tgb_occ_cells { const u32* p; __device__ bool operator()(u32 x, u32 y, u32 z) const { return (p[(z << 10) | (y << 5) | x] & TGB_TOP_HAS_DATA) != 0u; } };
struct tgb_get_cells { const u32* p; __device__ u32 operator()(u32 x, u32 y, u32 z) const { return p[(z << 10) | (y << 5) | x]; } };
struct tgb_occ_bricks { const u32* p; __device__ bool operator()(u32 x, u32 y, u32 z) const { return p[(z << 4) | (y << 2) | x] != 0u; } };
struct tgb_get_bricks { const u32* p; __device__ u32 operator()(u32 x, u32 y, u32 z) const { return p[(z << 4) | (y << 2) | x]; } };

/* one thread per table cell, one launch per pass (a pass reads its neighbours' values of the pass before) */
__global__ void k_fast_tile_cells(const u32* __restrict__ p_grid, u32* __restrict__ p_pass1, u32* __restrict__ p_pass2, u32* __restrict__ p_cells, u32 pass)
{
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= TGB_TOP_GRID_CELLS) return;
    const u32 x = c & 31u, y = (c >> 5) & 31u, z = c >> 10;
    const tgb_occ_cells occ = { p_grid };
    if (pass == 1u) p_pass1[c] = tgb_tile_pass1(occ, TGB_TOP_GRID_DIM, x, y, z);
    else if (pass == 2u) { const tgb_get_cells g1 = { p_pass1 }; p_pass2[c] = tgb_tile_pass2(occ, g1, TGB_TOP_GRID_DIM, x, y, z); }
    else
    {
        const tgb_get_cells g2 = { p_pass2 };
        p_cells[c] = occ(x, y, z) ? (TGB_CELLS_LEAF | (p_grid[c] & TGB_TOP_POINTER_MASK)) : tgb_tile_entry<5>(p_pass2[c], tgb_tile_pass3(occ, g2, TGB_TOP_GRID_DIM, x, y, z));
    }
}

/* 64 threads per leaf block (one per 8^3 brick), four leaf blocks per CTA: which bricks hold a solid voxel, then the three passes in shared memory */
__global__ void __launch_bounds__(256) k_fast_tile_bricks(const u32* __restrict__ p_voxels, u32 n_leaves, u32* __restrict__ p_bricks)
{
    __shared__ u32 s_solid[4][64], s_p1[4][64], s_p2[4][64];
    const u32 g = threadIdx.x >> 6, b = threadIdx.x & 63u, leaf = blockIdx.x * 4u + g;
    const u32 bx = b & 3u, by = (b >> 2) & 3u, bz = b >> 4;
    u32 any = 0;
    if (leaf < n_leaves)
    {
        const u32* p_block = p_voxels + (u64)leaf * TG_SVO_BLOCK_WORDS;
        for (u32 i = 0; i < 64u; i++) any |= p_block[32u * (8u * bz + (i >> 3)) + 8u * by + (i & 7u)];
        any = (any >> (8u * bx)) & 0xFFu;
    }
    s_solid[g][b] = any;
    __syncthreads();
    const tgb_occ_bricks occ = { s_solid[g] };
    s_p1[g][b] = tgb_tile_pass1(occ, 4u, bx, by, bz);
    __syncthreads();
    const tgb_get_bricks g1 = { s_p1[g] };
    s_p2[g][b] = tgb_tile_pass2(occ, g1, 4u, bx, by, bz);
    __syncthreads();
    const tgb_get_bricks g2 = { s_p2[g] };
    if (leaf < n_leaves)
        p_bricks[(u64)leaf * 64u + b] = any ? TGB_BRICK_SOLID : tgb_tile_brick_entry(s_p2[g][b], tgb_tile_pass3(occ, g2, 4u, bx, by, bz));
}

/* the leaf blocks' voxels once more with y, and with z, as the bit index (tgb_fast_tiling::p_columns): one CTA per block, one warp per z slice (y copy) and
 * per y line of slices (z copy), a 32 x 32 bit transpose by 32 ballots each */
__global__ void __launch_bounds__(1024) k_fast_tile_columns(const u32* __restrict__ p_voxels, u32* __restrict__ p_columns_y, u32* __restrict__ p_columns_z)
{
    const u32 lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const u32* p_block = p_voxels + (u64)blockIdx.x * TG_SVO_BLOCK_WORDS;
    const u32 row_y = p_block[32u * w + lane];    /* slice z = w, row y = lane */
    const u32 row_z = p_block[32u * lane + w];    /* slice z = lane, row y = w */
    u32 out_y = 0, out_z = 0;
#pragma unroll
    for (u32 x = 0; x < 32u; x++)
    {
        const u32 by = __ballot_sync(0xFFFFFFFFu, (row_y >> x) & 1u);   /* bit y = voxel (x, y, w) */
        const u32 bz = __ballot_sync(0xFFFFFFFFu, (row_z >> x) & 1u);   /* bit z = voxel (x, w, z) */
        if (lane == x) { out_y = by; out_z = bz; }
    }
    p_columns_y[(u64)blockIdx.x * TG_SVO_BLOCK_WORDS + 32u * w + lane] = out_y;   /* word 32 z + x */
    p_columns_z[(u64)blockIdx.x * TG_SVO_BLOCK_WORDS + 32u * w + lane] = out_z;   /* word 32 y + x */
}

/* DELTA0 of the certificate. TGB_GI_FAST_DELTA_PERCENT (tests, margin studies) scales it: the product runs at 100, a frame that still equals the exact
 * kernel's at 25 shows a fourfold margin on that frame's rays */
extern "C" f32 tgbd_gi_fast_delta(void)
{
    return TGB_FAST_DELTA * 0.01f * (f32)max(0, min(1000, tgbd_env_int("TGB_GI_FAST_DELTA_PERCENT", 100)));
}

extern "C" void tgbd_gi_fast_tiling_get(struct tgb_device* d, tgb_fast_tiling* p_tiling)
{
    p_tiling->p_cells = d->svo.d_fast_cells; p_tiling->p_bricks = d->svo.d_fast_bricks;
    p_tiling->p_columns = d->svo.d_fast_columns; p_tiling->columns_stride = d->svo.voxel_word_capacity;
}

/* called behind k_svo_flatten (tgb_svo.cu) on the stream of the build when the tiled walk is selected, else lazily by the first trace */
extern "C" b32 tgbd_gi_fast_tiling_build(struct tgb_device* d, cudaStream_t st)
{
    tgb_svo_device* s = &d->svo;
    if (!s->d_fast_cells)
    {
        TGB_CUDA(cudaMalloc(&s->d_fast_cells, (u64)3 * TGB_TOP_GRID_CELLS * sizeof(u32)));
        TGB_CUDA(cudaMalloc(&s->d_fast_bricks, (u64)s->leaf_capacity * 64u * sizeof(u32)));
        TGB_CUDA(cudaMalloc(&s->d_fast_columns, (u64)2 * s->voxel_word_capacity * sizeof(u32)));
    }
    u32* p1 = s->d_fast_cells + TGB_TOP_GRID_CELLS, *p2 = s->d_fast_cells + 2 * TGB_TOP_GRID_CELLS;
    for (u32 pass = 1; pass <= 3u; pass++)
    {
        k_fast_tile_cells<<<TGB_TOP_GRID_CELLS / 256, 256, 0, st>>>(s->d_top_grid, p1, p2, s->d_fast_cells, pass);
        TGB_LAUNCH_CHECK(d);
    }
    if (s->n_leaves)
    {
        k_fast_tile_bricks<<<(s->n_leaves + 3u) / 4u, 256, 0, st>>>(s->d_voxels, s->n_leaves, s->d_fast_bricks);
        TGB_LAUNCH_CHECK(d);
        k_fast_tile_columns<<<s->n_leaves, 1024, 0, st>>>(s->d_voxels, s->d_fast_columns, s->d_fast_columns + s->voxel_word_capacity);
        TGB_LAUNCH_CHECK(d);
    }
    s->fast_tiling_valid = TG_TRUE;
    return TG_TRUE;
}

/*
 * One band of rays (called by tgbd__shade_launch, tgb_shade.cu, after k_shade queued them): the fast walk over the whole queue, then
 * the exact kernel over the slots it handed over (p_q_count[12] of them, read on the device).
 */
__global__ void k_gi_reset_lists(u32* __restrict__ p_q_count)
{
    if (threadIdx.x < 2u) { p_q_count[12u + threadIdx.x] = 0u; p_q_count[16u + threadIdx.x] = 0u; } /* handed over / fetched: first pass, careful pass */
}

extern "C" b32 tgbd_gi_fast_trace(struct tgb_device* d, f32 far_plane, b32 tiled)
{
    if (tiled && !d->svo.fast_tiling_valid && !tgbd_gi_fast_tiling_build(d, d->stream)) return TG_FALSE;
    const u32 ctas_per_sm = (u32)max(1, min(16, tgbd_env_int("TGB_GI_FAST_CTAS_PER_SM", 8)));
    const u32 service_lanes = (u32)max(1, min(32, tgbd_env_int("TGB_GI_FAST_SERVICE_LANES", 8)));
    const u32 steps = (u32)max(1, tgbd_env_int("TGB_GI_FAST_STEPS", tiled ? 4 : 8)); /* cells per walk phase: measured (profiles/r04g_sweep_full.jsonl) */
    const u32 max_steps = (u32)max(1, tgbd_env_int("TGB_GI_FAST_MAX_STEPS", (i32)TGB_FAST_MAX_STEPS)), max_steps_uncertain = (u32)max(1, tgbd_env_int("TGB_GI_FAST_MAX_STEPS_UNCERTAIN", (i32)TGB_FAST_MAX_STEPS_UNCERTAIN));
    tgb_gi_frame fr;
    tgb_gi_frame_init(&fr, d->svo.bmin, d->svo.bmax, far_plane, d->svo.d_top_grid, d->svo.d_voxels);
    k_gi_reset_lists<<<1, 32, 0, d->stream>>>(d->d_gi_count);
    TGB_LAUNCH_CHECK(d);
    const f32 delta = tgbd_gi_fast_delta();
    tgb_fast_tiling tiling;
    tgbd_gi_fast_tiling_get(d, &tiling);
    u32* p_list_a = d->d_gi_exact, *p_list_b = d->d_gi_exact + (u64)d->width * d->height;
#define TGB_FAST_ARGS(IN, IN_WORD, OUT, OUT_WORD, SVC, STEPS, CAP, CAP_U) fr, tiling, IN, IN_WORD, OUT_WORD, d->d_gi_q0, d->d_gi_q1, d->d_gi_q2, d->d_gi_count, OUT, d->d_radiance, SVC, STEPS, CAP, CAP_U, delta
    if (tiled) k_gi_trace_fast<true, false><<<d->n_sms * ctas_per_sm, TGB_FAST_THREADS, 0, d->stream>>>(TGB_FAST_ARGS((const u32*)NULL, 0u, p_list_a, 12u, service_lanes, steps, max_steps, max_steps_uncertain));
    else       k_gi_trace_fast<false, false><<<d->n_sms * ctas_per_sm, TGB_FAST_THREADS, 0, d->stream>>>(TGB_FAST_ARGS((const u32*)NULL, 0u, p_list_a, 12u, service_lanes, steps, max_steps, max_steps_uncertain));
    TGB_LAUNCH_CHECK(d);
    if (tiled && tgbd_env_int("TGB_GI_FAST_CAREFUL", 0))
    {
        /* the careful pass over what was handed over; its own hand-overs go to the shader's arithmetic. OFF: it decides four handed-over rays of
         * five, but what the second stage takes is the chain of its longest ray, and the longest of the remaining fifth is as long as
         * the longest of all (profiles/r04f_*: 1.25 ms for the stage with it, 1.01 without) */
        const u32 careful_ctas = (u32)max(1, min(16, tgbd_env_int("TGB_GI_FAST_CAREFUL_CTAS_PER_SM", 4)));
        k_gi_trace_fast<true, true><<<d->n_sms * careful_ctas, TGB_FAST_THREADS, 0, d->stream>>>(TGB_FAST_ARGS(p_list_a, 12u, p_list_b, 16u, (u32)max(1, min(32, tgbd_env_int("TGB_GI_FAST_CAREFUL_SERVICE_LANES", 1))), steps, max_steps, max_steps_uncertain));
        TGB_LAUNCH_CHECK(d);
        return tgbd_gi_pool_trace_list(d, far_plane, p_list_b, 16u);
    }
#undef TGB_FAST_ARGS
    return tgbd_gi_pool_trace_list(d, far_plane, p_list_a, 12u);
}

/*
 * tgb_procedural.cu -- the reference's procedural object fill on the device.
 *
 * tg_raytracer_create_object (graphics/vulkan/tgvk_raytracer.c:868-943, "// TODO: gen on GPU") evaluates, on one CPU
 * thread, three simplex-noise samples per voxel (math/tg_math.c:182-302) and packs the `solid` bits LSB-first, x fastest,
 * 16 u32 per cluster. Here one thread produces one 32-bit word (8 x-values x 4 y-values of one z): the two terrain
 * samples depend on (x, z) only and are evaluated once per x, the cave sample once per voxel. The arithmetic is the
 * reference's, operation for operation (this TU is built with -fmad=false like the others; float -> int conversions
 * truncate like the C casts), so the bits equal the oracle's restatement (oracle/tgo_procedural.c), which
 * tests/test_reference_pins.py pins against the reference's own tgm_simplex_noise.
 */
#include "tgb_device.cuh"

/* math/tg_math.c:11-15 */
__constant__ signed char c_simplex_gradients[12][3] = {
    {  1,  1,  0 }, { -1,  1,  0 }, {  1, -1,  0 }, { -1, -1,  0 },
    {  1,  0,  1 }, { -1,  0,  1 }, {  1,  0, -1 }, { -1,  0, -1 },
    {  0,  1,  1 }, {  0, -1,  1 }, {  0,  1, -1 }, {  0, -1, -1 }
};
/* math/tg_math.c:17-83 (the reference stores the 256 entries twice; indexed modulo 256 here) */
__constant__ unsigned char c_simplex_permutation[256] = {
    151, 160, 137,  91,  90,  15, 131,  13, 201,  95,  96,  53, 194, 233,   7, 225,
    140,  36, 103,  30,  69, 142,   8,  99,  37, 240,  21,  10,  23, 190,   6, 148,
    247, 120, 234,  75,   0,  26, 197,  62,  94, 252, 219, 203, 117,  35,  11,  32,
     57, 177,  33,  88, 237, 149,  56,  87, 174,  20, 125, 136, 171, 168,  68, 175,
     74, 165,  71, 134, 139,  48,  27, 166,  77, 146, 158, 231,  83, 111, 229, 122,
     60, 211, 133, 230, 220, 105,  92,  41,  55,  46, 245,  40, 244, 102, 143,  54,
     65,  25,  63, 161,   1, 216,  80,  73, 209,  76, 132, 187, 208,  89,  18, 169,
    200, 196, 135, 130, 116, 188, 159,  86, 164, 100, 109, 198, 173, 186,   3,  64,
     52, 217, 226, 250, 124, 123,   5, 202,  38, 147, 118, 126, 255,  82,  85, 212,
    207, 206,  59, 227,  47,  16,  58,  17, 182, 189,  28,  42, 223, 183, 170, 213,
    119, 248, 152,   2,  44, 154, 163,  70, 221, 153, 101, 155, 167,  43, 172,   9,
    129,  22,  39, 253,  19,  98, 108, 110,  79, 113, 224, 232, 178, 185, 112, 104,
    218, 246,  97, 228, 251,  34, 242, 193, 238, 210, 144,  12, 191, 179, 162, 241,
     81,  51, 145, 235, 249,  14, 239, 107,  49, 192, 214,  31, 181, 199, 106, 157,
    184,  84, 204, 176, 115, 121,  50,  45, 127,   4, 150, 254, 138, 236, 205,  93,
    222, 114,  67,  29,  24,  72, 243, 141, 128, 195,  78,  66, 215,  61, 156, 180
};

__device__ __forceinline__ i32 tgb_perm(i32 i) { return (i32)c_simplex_permutation[i & 255]; }
__device__ __forceinline__ i32 tgb_fastfloor(f32 x) { return x > 0.0f ? (i32)x : (i32)x - 1; } /* math/tg_math.c:184 */

__device__ __forceinline__ f32 tgb_simplex_corner(f32 x, f32 y, f32 z, i32 gi)
{
    /* math/tg_math.c:247-258 */
    f32 t = 0.5f - x * x - y * y - z * z;
    if (t < 0.0f) return 0.0f;
    t *= t;
    const f32 dot = (f32)c_simplex_gradients[gi][0] * x + (f32)c_simplex_gradients[gi][1] * y + (f32)c_simplex_gradients[gi][2] * z;
    return t * t * dot;
}

/* math/tg_math.c:182-302 */
__device__ f32 tgb_simplex_noise(f32 x, f32 y, f32 z)
{
    const f32 s = (x + y + z) * 0.333333343f;
    const i32 i = tgb_fastfloor(x + s);
    const i32 j = tgb_fastfloor(y + s);
    const i32 k = tgb_fastfloor(z + s);

    const f32 g3 = 0.166666672f;
    const f32 t = (f32)(i + j + k) * g3;
    const f32 x0 = x - ((f32)i - t);
    const f32 y0 = y - ((f32)j - t);
    const f32 z0 = z - ((f32)k - t);

    i32 i1, j1, k1, i2, j2, k2;
    if (x0 >= y0)
    {
        if (y0 >= z0)      { i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 1; k2 = 0; }
        else if (x0 >= z0) { i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 0; k2 = 1; }
        else               { i1 = 0; j1 = 0; k1 = 1; i2 = 1; j2 = 0; k2 = 1; }
    }
    else
    {
        if (y0 < z0)       { i1 = 0; j1 = 0; k1 = 1; i2 = 0; j2 = 1; k2 = 1; }
        else if (x0 < z0)  { i1 = 0; j1 = 1; k1 = 0; i2 = 0; j2 = 1; k2 = 1; }
        else               { i1 = 0; j1 = 1; k1 = 0; i2 = 1; j2 = 1; k2 = 0; }
    }

    const f32 x1 = x0 - (f32)i1 + g3,        y1 = y0 - (f32)j1 + g3,        z1 = z0 - (f32)k1 + g3;
    const f32 x2 = x0 - (f32)i2 + 2.0f * g3, y2 = y0 - (f32)j2 + 2.0f * g3, z2 = z0 - (f32)k2 + 2.0f * g3;
    const f32 x3 = x0 - 1.0f + 3.0f * g3,    y3 = y0 - 1.0f + 3.0f * g3,    z3 = z0 - 1.0f + 3.0f * g3;

    const i32 ii = i & 255, jj = j & 255, kk = k & 255;
    const i32 gi0 = tgb_perm(ii      + tgb_perm(jj      + tgb_perm(kk     ))) % 12;
    const i32 gi1 = tgb_perm(ii + i1 + tgb_perm(jj + j1 + tgb_perm(kk + k1))) % 12;
    const i32 gi2 = tgb_perm(ii + i2 + tgb_perm(jj + j2 + tgb_perm(kk + k2))) % 12;
    const i32 gi3 = tgb_perm(ii +  1 + tgb_perm(jj +  1 + tgb_perm(kk +  1))) % 12;

    const f32 n0 = tgb_simplex_corner(x0, y0, z0, gi0);
    const f32 n1 = tgb_simplex_corner(x1, y1, z1, gi1);
    const f32 n2 = tgb_simplex_corner(x2, y2, z2, gi2);
    const f32 n3 = tgb_simplex_corner(x3, y3, z3, gi3);
    return 32.0f * (n0 + n1 + n2 + n3);
}

/* math/tg_math.c:584-606: max(low, min(high, v)) with C ternaries */
__device__ __forceinline__ f32 tgb_clamp_c(f32 v, f32 low, f32 high)
{
    const f32 m = high < v ? high : v;
    return low > m ? low : m;
}
/* math/tg_math.c:614-618 */
__device__ __forceinline__ i32 tgb_round_to_i32(f32 v) { return v >= 0.0f ? (i32)(v + 0.5f) : (i32)(-(fabsf(v) + 0.5f)); }

/*
 * One thread = one u32 of one cluster. `p_dst_by_pointer != NULL`: the word goes to the cluster the pointer table names
 * (p_masks[cluster_idx * 16 + word]); otherwise to a dense staging array in pointer order (tgb200_procedural_solid_bits).
 */
__global__ void __launch_bounds__(128) k_procedural_fill(u32 object_idx, u32 nx, u32 ny, u32 nz, const u32* __restrict__ p_cluster_pointers, u32 first_pointer,
                                                         u32* __restrict__ p_masks, u32* __restrict__ p_dense)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 rel = t >> 4, word = t & 15u;
    if (rel >= nx * ny * nz) return;
    const u32 cx = rel % nx, cy = (rel / nx) % ny, cz = rel / (nx * ny);
    /* bit b of word w is voxel 32 w + b of the cluster: x = b & 7, y = 4 (w & 1) + (b >> 3), z = w >> 1 (tgvk_raytracer.c:894-937) */
    const u32 voxel_z = 8u * cz + (word >> 1);
    const f32 zf = (f32)voxel_z;
    u32 bits = 0;
    for (u32 bx = 0; bx < 8; bx++)
    {
        const u32 voxel_x = 8u * cx + bx;
        const f32 xf = (f32)voxel_x + (f32)object_idx * 1024.0f;
        /* tgvk_raytracer.c:904-906: the terrain height does not depend on y */
        const f32 n_hills0 = tgb_simplex_noise(xf * 0.008f, 0.0f, zf * 0.008f);
        const f32 n_hills1 = tgb_simplex_noise(xf * 0.2f, 0.0f, zf * 0.2f);
        const f32 n_hills = n_hills0 + 0.005f * n_hills1;
        for (u32 by = 0; by < 4; by++)
        {
            const u32 voxel_y = 8u * cy + 4u * (word & 1u) + by;
            const f32 yf = (f32)voxel_y;
            /* :908-913 */
            const f32 s_caves = 0.06f;
            const f32 unclamped_noise_caves = tgb_simplex_noise(s_caves * xf, s_caves * yf, s_caves * zf);
            const f32 n_caves = tgb_clamp_c(unclamped_noise_caves, -1.0f, 0.0f);
            /* :915-921 */
            const f32 noise = (n_hills * 64.0f) - ((f32)voxel_y - 8.0f) + (10.0f * n_caves);
            const f32 noise_clamped = tgb_clamp_c(noise, -1.0f, 1.0f);
            const f32 f0 = (noise_clamped + 1.0f) * 0.5f;
            const f32 f1 = 254.0f * f0;
            const i32 f2 = -(i32)(signed char)(tgb_round_to_i32(f1) - 127);
            const bool solid = f2 <= 0 || voxel_y == 0;
            bits |= (solid ? 1u : 0u) << (8u * by + bx);
        }
    }
    if (p_dense) p_dense[(u64)rel * TG_CLUSTER_MASK_WORDS + word] = bits;
    else         p_masks[(u64)p_cluster_pointers[first_pointer + rel] * TG_CLUSTER_MASK_WORDS + word] = bits;
}

/* the bits of object `object_idx`'s clusters [first_pointer, first_pointer + nx ny nz) straight into the resident mask array */
extern "C" b32 tgbd_procedural_fill(struct tgb_device* d, u32 object_idx, u32 nx, u32 ny, u32 nz, u32 first_pointer)
{
    TGB_CUDA(cudaSetDevice(d->device));
    const u64 n_threads = (u64)nx * ny * nz * 16u;
    if (n_threads == 0) return TG_TRUE;
    k_procedural_fill<<<(u32)((n_threads + 127) / 128), 128, 0, d->stream>>>(object_idx, nx, ny, nz, d->d_cluster_pointers, first_pointer, d->d_masks, NULL);
    TGB_LAUNCH_CHECK(d);
    return TG_TRUE;
}

/* the same bits into host memory, clusters in pointer order (no raytracer needed: a scratch buffer on `device`) */
extern "C" b32 tgbd_procedural_bits_to_host(i32 device, u32 object_idx, u32 nx, u32 ny, u32 nz, u32* p_out)
{
    TGB_CUDA(cudaSetDevice(device));
    const u64 n_words = (u64)nx * ny * nz * TG_CLUSTER_MASK_WORDS;
    if (n_words == 0) return TG_TRUE;
    u32* d_dense = NULL;
    TGB_CUDA(cudaMalloc(&d_dense, n_words * sizeof(u32)));
    k_procedural_fill<<<(u32)((n_words + 127) / 128), 128>>>(object_idx, nx, ny, nz, NULL, 0, NULL, d_dense);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(p_out, d_dense, n_words * sizeof(u32), cudaMemcpyDeviceToHost);
    cudaFree(d_dense);
    if (e != cudaSuccess) { tgb_set_error("procedural fill: %s", cudaGetErrorString(e)); return TG_FALSE; }
    return TG_TRUE;
}

/* ---- seeded random fill (BASELINE.json's "random solid bits" configs; SURVEY.md section 8d synthetic inputs) ---------------- */
/*
 * One thread = one cluster: state0 = hash_u32(object_seed ^ hash_u32(rel_cluster)) | 1 (math/tg_math.c:809-820 = util.inc:47-56),
 * each of the 16 mask words is the AND of `k` successive xorshift32 draws (math/tg_math.c:328-338): density 2^-k. The same
 * definition as tg_b200/scenes.py random_solid_bits and tgb200_synthetic_solid_bits (host), which tests compare bit for bit.
 * Four 16-byte stores per thread; a warp writes 2 KiB contiguously when the object's cluster indices are one run.
 */
__global__ void __launch_bounds__(128) k_synthetic_fill(u32 object_seed, u32 k, u32 n_clusters, const u32* __restrict__ p_cluster_pointers, u32 first_pointer, u32* __restrict__ p_masks)
{
    const u32 rel = blockIdx.x * blockDim.x + threadIdx.x;
    if (rel >= n_clusters) return;
    u32 state = tgb_hash_u32(object_seed ^ tgb_hash_u32(rel)) | 1u;
    uint4* p_dst = reinterpret_cast<uint4*>(p_masks + (u64)p_cluster_pointers[first_pointer + rel] * TG_CLUSTER_MASK_WORDS);
#pragma unroll
    for (u32 q = 0; q < 4; q++)
    {
        u32 w[4];
#pragma unroll
        for (u32 i = 0; i < 4; i++)
        {
            u32 word = 0xFFFFFFFFu;
            for (u32 j = 0; j < k; j++) word &= tgb_xorshift32(&state);
            w[i] = word;
        }
        p_dst[q] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

extern "C" b32 tgbd_synthetic_fill(struct tgb_device* d, u32 object_seed, u32 k, u32 n_clusters, u32 first_pointer)
{
    TGB_CUDA(cudaSetDevice(d->device));
    if (n_clusters == 0) return TG_TRUE;
    k_synthetic_fill<<<(n_clusters + 127) / 128, 128, 0, d->stream>>>(object_seed, k, n_clusters, d->d_cluster_pointers, first_pointer, d->d_masks);
    TGB_LAUNCH_CHECK(d);
    return TG_TRUE;
}

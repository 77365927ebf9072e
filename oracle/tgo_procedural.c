/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 * Restatement of the reference's procedural object fill: tgm_simplex_noise (math/tg_math.c:182-302, tables :11-83) and
 * the per-voxel terrain rule of tg_raytracer_create_object (graphics/vulkan/tgvk_raytracer.c:868-943), plus exported
 * twins of the tgo_math.h routines so tests/test_reference_pins.py can put each of them next to the reference's own
 * function compiled into oracle/_ref/libtg_ref.so.
 */
#include "tgo.h"
#include "tgo_math.h"

/* math/tg_math.c:11-15 */
static const i8 tgo_simplex_gradients[12][3] = {
    {  1,  1,  0 }, { -1,  1,  0 }, {  1, -1,  0 }, { -1, -1,  0 },
    {  1,  0,  1 }, { -1,  0,  1 }, {  1,  0, -1 }, { -1,  0, -1 },
    {  0,  1,  1 }, {  0, -1,  1 }, {  0,  1, -1 }, {  0, -1, -1 }
};

/* math/tg_math.c:17-83: Ken Perlin's permutation, stored twice by the reference; indexed modulo 256 here */
static const u8 tgo_simplex_permutation[256] = {
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
static inline i32 tgo_perm(i32 i) { return (i32)tgo_simplex_permutation[i & 255]; } /* the doubled table: index < 512 */

static inline i32 tgo_fastfloor(f32 x) { return x > 0.0f ? (i32)x : (i32)x - 1; } /* math/tg_math.c:184 */

static inline f32 tgo_simplex_corner(f32 x, f32 y, f32 z, i32 gi)
{
    /* math/tg_math.c:247-258 (and its three copies): t = 0.5 - x^2 - y^2 - z^2, left to right */
    f32 t = 0.5f - x * x - y * y - z * z;
    if (t < 0.0f) return 0.0f;
    t *= t;
    const f32 dot = (f32)tgo_simplex_gradients[gi][0] * x + (f32)tgo_simplex_gradients[gi][1] * y + (f32)tgo_simplex_gradients[gi][2] * z;
    return t * t * dot;
}

/* math/tg_math.c:182-302 */
f32 tgo_simplex_noise(f32 x, f32 y, f32 z)
{
    const f32 s = (x + y + z) * 0.333333343f;
    const i32 i = tgo_fastfloor(x + s);
    const i32 j = tgo_fastfloor(y + s);
    const i32 k = tgo_fastfloor(z + s);

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

    /* :229-237: x0 - i1 + g3 is (x0 - (f32)i1) + g3; 2.0f * g3 and 3.0f * g3 are float products */
    const f32 x1 = x0 - (f32)i1 + g3,        y1 = y0 - (f32)j1 + g3,        z1 = z0 - (f32)k1 + g3;
    const f32 x2 = x0 - (f32)i2 + 2.0f * g3, y2 = y0 - (f32)j2 + 2.0f * g3, z2 = z0 - (f32)k2 + 2.0f * g3;
    const f32 x3 = x0 - 1.0f + 3.0f * g3,    y3 = y0 - 1.0f + 3.0f * g3,    z3 = z0 - 1.0f + 3.0f * g3;

    const i32 ii = i & 255, jj = j & 255, kk = k & 255;
    const i32 gi0 = tgo_perm(ii      + tgo_perm(jj      + tgo_perm(kk     ))) % 12;
    const i32 gi1 = tgo_perm(ii + i1 + tgo_perm(jj + j1 + tgo_perm(kk + k1))) % 12;
    const i32 gi2 = tgo_perm(ii + i2 + tgo_perm(jj + j2 + tgo_perm(kk + k2))) % 12;
    const i32 gi3 = tgo_perm(ii +  1 + tgo_perm(jj +  1 + tgo_perm(kk +  1))) % 12;

    const f32 n0 = tgo_simplex_corner(x0, y0, z0, gi0);
    const f32 n1 = tgo_simplex_corner(x1, y1, z1, gi1);
    const f32 n2 = tgo_simplex_corner(x2, y2, z2, gi2);
    const f32 n3 = tgo_simplex_corner(x3, y3, z3, gi3);
    return 32.0f * (n0 + n1 + n2 + n3);
}

/* math/tg_math.c:584-588: max(low, min(high, v)) with the C ternaries of tgm_f32_min / tgm_f32_max (:  v0 < v1 ? v0 : v1) */
static inline f32 tgo_f32_clamp_c(f32 v, f32 low, f32 high)
{
    const f32 m = high < v ? high : v;
    return low > m ? low : m;
}

/* math/tg_math.c:614-618 */
static inline i32 tgo_round_to_i32(f32 v) { return v >= 0.0f ? (i32)(v + 0.5f) : (i32)(-((v < 0.0f ? -v : v) + 0.5f)); }

/* tgvk_raytracer.c:900-928: is voxel (voxel_x, voxel_y, voxel_z) of object `object_idx` solid? */
b32 tgo_procedural_voxel_is_solid(u32 object_idx, u32 voxel_x, u32 voxel_y, u32 voxel_z)
{
    const f32 xf = (f32)voxel_x + (f32)object_idx * 1024.0f; /* u32 * float: the index is converted first */
    const f32 yf = (f32)voxel_y;
    const f32 zf = (f32)voxel_z;

    const f32 n_hills0 = tgo_simplex_noise(xf * 0.008f, 0.0f, zf * 0.008f);
    const f32 n_hills1 = tgo_simplex_noise(xf * 0.2f, 0.0f, zf * 0.2f);
    const f32 n_hills = n_hills0 + 0.005f * n_hills1;

    const f32 s_caves = 0.06f;
    const f32 unclamped_noise_caves = tgo_simplex_noise(s_caves * xf, s_caves * yf, s_caves * zf);
    const f32 n_caves = tgo_f32_clamp_c(unclamped_noise_caves, -1.0f, 0.0f);

    const f32 noise = (n_hills * 64.0f) - ((f32)voxel_y - 8.0f) + (10.0f * n_caves);
    const f32 noise_clamped = tgo_f32_clamp_c(noise, -1.0f, 1.0f);
    const f32 f0 = (noise_clamped + 1.0f) * 0.5f;
    const f32 f1 = 254.0f * f0;
    const i8 f2 = (i8)(-(i8)(tgo_round_to_i32(f1) - 127));
    return f2 <= 0 || voxel_y == 0;
}

/* tgvk_raytracer.c:871-943: 16 u32 per cluster, clusters in pointer order (x fastest, then y, then z), bit 64 z + 8 y + x */
void tgo_procedural_solid_bits(u32 object_idx, v3u dims, u32* p_out)
{
    const u32 n = dims.x * dims.y * dims.z;
#pragma omp parallel for schedule(dynamic, 16)
    for (u32 rel = 0; rel < n; rel++)
    {
        const u32 cx = rel % dims.x, cy = (rel / dims.x) % dims.y, cz = rel / (dims.x * dims.y);
        for (u32 w = 0; w < 16; w++)
        {
            u32 bits = 0;
            for (u32 b = 0; b < 32; b++)
            {
                const u32 v = 32 * w + b;
                if (tgo_procedural_voxel_is_solid(object_idx, 8 * cx + (v & 7u), 8 * cy + ((v >> 3) & 7u), 8 * cz + (v >> 6))) bits |= 1u << b;
            }
            p_out[(u64)rel * 16u + w] = bits;
        }
    }
}

/* ---- exported twins of tgo_math.h for the reference pins -------------------------------------------------- */
m4  tgo_pin_m4_mul(m4 a, m4 b) { return tgo_m4_mul(a, b); }
m4  tgo_pin_m4_inverse(m4 m) { return tgo_m4_inverse(m); }
m4  tgo_pin_m4_angle_axis(f32 angle_in_radians, v3 axis) { return tgo_m4_angle_axis(angle_in_radians, axis); }
m4  tgo_pin_m4_euler(f32 pitch, f32 yaw, f32 roll) { return tgo_m4_euler(pitch, yaw, roll); }
m4  tgo_pin_m4_perspective(f32 fov_y, f32 aspect, f32 n, f32 f) { return tgo_m4_perspective(fov_y, aspect, n, f); }
m4  tgo_pin_m4_translate(v3 v) { return tgo_m4_translate(v); }
v4  tgo_pin_m4_mulv4(m4 m, v4 v) { return tgo_m4_mulv4(m, v); }
v3  tgo_pin_v3_normalized(v3 v) { return tgo_v3_normalized(v); }
v3  tgo_pin_v3_lerp(v3 a, v3 b, f32 t) { return tgo_v3_lerp(a, b, t); }
u32 tgo_pin_xorshift32_next(u32* p_state) { return tgo_xorshift32_next(p_state); }
f32 tgo_pin_xorshift32_next_f32(u32* p_state) { return tgo_xorshift32_next_f32(p_state); }
f32 tgo_pin_xorshift32_next_f32_range(u32* p_state, f32 lo, f32 hi) { return tgo_xorshift32_next_f32_range(p_state, lo, hi); }
b32 tgo_pin_intersect_ray_aabb_c(v3 o, v3 d, v3 bmin, v3 bmax, f32* p_enter, f32* p_exit) { return tgo_intersect_ray_aabb_c(o, d, bmin, bmax, p_enter, p_exit); }

/*
 * tgb_math.h -- float32 math shared by the C host code and the CUDA kernels of libtgb200.
 *
 * The visibility buffer and the SVO must match the reference's arithmetic bit for bit, so every
 * function spells out the reference's operation order (paths under /root/reference/tg/src, cited
 * per function) and the translation units that include this are built WITHOUT FMA contraction
 * (nvcc -fmad=false, gcc -ffp-contract=off), with IEEE division and square root. Trigonometry is
 * host-only (sinf/cosf/tanf differ between libm and CUDA; the reference also evaluates them on the
 * CPU: tgvk_raytracer.c:841,1171-1178).
 *
 * GLSL builtins are pinned as: min(x,y) = y<x?y:x, max(x,y) = x<y?y:x, mix(a,b,t) = a*(1-t)+b*t,
 * normalize(v) = v / sqrt(dot(v,v)), inverse(mat4) = cofactor expansion (tgm_m4_inverse).
 */
#ifndef TGB_MATH_H
#define TGB_MATH_H

#include <math.h>
#include "../../include/tg_types.h"

#ifdef __CUDACC__
#define TGB_HD __host__ __device__ __forceinline__
#else
#define TGB_HD static inline
#endif

TGB_HD f32 tgb_min(f32 x, f32 y) { return y < x ? y : x; }
TGB_HD f32 tgb_max(f32 x, f32 y) { return x < y ? y : x; }
TGB_HD f32 tgb_clamp(f32 x, f32 lo, f32 hi) { return tgb_min(tgb_max(x, lo), hi); }
TGB_HD f32 tgb_mix(f32 a, f32 b, f32 t) { return a * (1.0f - t) + b * t; }
TGB_HD f32 tgb_sign(f32 x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

TGB_HD v3 tgb_v3(f32 x, f32 y, f32 z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
TGB_HD v3 tgb_add(v3 a, v3 b) { return tgb_v3(a.x + b.x, a.y + b.y, a.z + b.z); }
TGB_HD v3 tgb_sub(v3 a, v3 b) { return tgb_v3(a.x - b.x, a.y - b.y, a.z - b.z); }
TGB_HD v3 tgb_mul(v3 a, v3 b) { return tgb_v3(a.x * b.x, a.y * b.y, a.z * b.z); }
TGB_HD v3 tgb_scale(v3 a, f32 f) { return tgb_v3(a.x * f, a.y * f, a.z * f); }
TGB_HD v3 tgb_neg(v3 a) { return tgb_v3(-a.x, -a.y, -a.z); }
/* math/tg_math.c:1064-1068 */
TGB_HD f32 tgb_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* math/tg_math.c:1106-1110 */
TGB_HD f32 tgb_length(v3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }
/* math/tg_math.c:1175-1184 */
TGB_HD v3 tgb_normalize(v3 v)
{
    const f32 len = tgb_length(v);
    return tgb_v3(v.x / len, v.y / len, v.z / len);
}
TGB_HD v3 tgb_divf(v3 v, f32 f) { return tgb_v3(v.x / f, v.y / f, v.z / f); }
/* math/tg_math.c:1118-1140 (C ternaries, used by the SVO builder) */
TGB_HD v3 tgb_cmax(v3 a, v3 b) { return tgb_v3(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z); }
TGB_HD v3 tgb_cmin(v3 a, v3 b) { return tgb_v3(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z); }
TGB_HD v3 tgb_mix3(v3 a, v3 b, f32 t) { return tgb_v3(tgb_mix(a.x, b.x, t), tgb_mix(a.y, b.y, t), tgb_mix(a.z, b.z, t)); }

TGB_HD m4 tgb_m4_identity(void)
{
    m4 r;
    r.m00 = 1.0f; r.m10 = 0.0f; r.m20 = 0.0f; r.m30 = 0.0f;
    r.m01 = 0.0f; r.m11 = 1.0f; r.m21 = 0.0f; r.m31 = 0.0f;
    r.m02 = 0.0f; r.m12 = 0.0f; r.m22 = 1.0f; r.m32 = 0.0f;
    r.m03 = 0.0f; r.m13 = 0.0f; r.m23 = 0.0f; r.m33 = 1.0f;
    return r;
}

/* math/tg_math.c:2669-2694 */
TGB_HD m4 tgb_m4_translate(v3 t)
{
    m4 r = tgb_m4_identity();
    r.m03 = t.x; r.m13 = t.y; r.m23 = t.z;
    return r;
}

/* math/tg_math.c:2374-2399: each entry = ((p0 + p1) + p2) + p3 */
#define TGB_ROWCOL(a, b, i, j) (a.m##i##0 * b.m0##j + a.m##i##1 * b.m1##j + a.m##i##2 * b.m2##j + a.m##i##3 * b.m3##j)
TGB_HD m4 tgb_m4_mul(m4 a, m4 b)
{
    m4 r;
    r.m00 = TGB_ROWCOL(a, b, 0, 0); r.m10 = TGB_ROWCOL(a, b, 1, 0); r.m20 = TGB_ROWCOL(a, b, 2, 0); r.m30 = TGB_ROWCOL(a, b, 3, 0);
    r.m01 = TGB_ROWCOL(a, b, 0, 1); r.m11 = TGB_ROWCOL(a, b, 1, 1); r.m21 = TGB_ROWCOL(a, b, 2, 1); r.m31 = TGB_ROWCOL(a, b, 3, 1);
    r.m02 = TGB_ROWCOL(a, b, 0, 2); r.m12 = TGB_ROWCOL(a, b, 1, 2); r.m22 = TGB_ROWCOL(a, b, 2, 2); r.m32 = TGB_ROWCOL(a, b, 3, 2);
    r.m03 = TGB_ROWCOL(a, b, 0, 3); r.m13 = TGB_ROWCOL(a, b, 1, 3); r.m23 = TGB_ROWCOL(a, b, 2, 3); r.m33 = TGB_ROWCOL(a, b, 3, 3);
    return r;
}
#undef TGB_ROWCOL

/* math/tg_math.c:2428-2438, xyz of m * (v, w) */
TGB_HD v3 tgb_m4_transform(m4 m, v3 v, f32 w)
{
    v3 r;
    r.x = v.x * m.m00 + v.y * m.m01 + v.z * m.m02 + w * m.m03;
    r.y = v.x * m.m10 + v.y * m.m11 + v.z * m.m12 + w * m.m13;
    r.z = v.x * m.m20 + v.y * m.m21 + v.z * m.m22 + w * m.m23;
    return r;
}

/* math/tg_math.c:2185-2234: 2x2 minors s_<rows>_<cols>, then cofactors scaled by 1/det */
TGB_HD m4 tgb_m4_inverse(m4 m)
{
    const f32 s23_23 = m.m22 * m.m33 - m.m23 * m.m32;
    const f32 s13_23 = m.m21 * m.m33 - m.m23 * m.m31;
    const f32 s12_23 = m.m21 * m.m32 - m.m22 * m.m31;
    const f32 s03_23 = m.m20 * m.m33 - m.m23 * m.m30;
    const f32 s02_23 = m.m20 * m.m32 - m.m22 * m.m30;
    const f32 s01_23 = m.m20 * m.m31 - m.m21 * m.m30;
    const f32 s23_13 = m.m12 * m.m33 - m.m13 * m.m32;
    const f32 s13_13 = m.m11 * m.m33 - m.m13 * m.m31;
    const f32 s12_13 = m.m11 * m.m32 - m.m12 * m.m31;
    const f32 s23_12 = m.m12 * m.m23 - m.m13 * m.m22;
    const f32 s13_12 = m.m11 * m.m23 - m.m13 * m.m21;
    const f32 s12_12 = m.m11 * m.m22 - m.m12 * m.m21;
    const f32 s03_13 = m.m10 * m.m33 - m.m13 * m.m30;
    const f32 s02_13 = m.m10 * m.m32 - m.m12 * m.m30;
    const f32 s03_12 = m.m10 * m.m23 - m.m13 * m.m20;
    const f32 s02_12 = m.m10 * m.m22 - m.m12 * m.m20;
    const f32 s01_13 = m.m10 * m.m31 - m.m11 * m.m30;
    const f32 s01_12 = m.m10 * m.m21 - m.m11 * m.m20;

    const f32 inv_det = 1.0f / (
        m.m00 * (m.m11 * s23_23 - m.m12 * s13_23 + m.m13 * s12_23) -
        m.m01 * (m.m10 * s23_23 - m.m12 * s03_23 + m.m13 * s02_23) +
        m.m02 * (m.m10 * s13_23 - m.m11 * s03_23 + m.m13 * s01_23) -
        m.m03 * (m.m10 * s12_23 - m.m11 * s02_23 + m.m12 * s01_23));

    m4 r;
    r.m00 = inv_det *  (m.m11 * s23_23 - m.m12 * s13_23 + m.m13 * s12_23);
    r.m01 = inv_det * -(m.m01 * s23_23 - m.m02 * s13_23 + m.m03 * s12_23);
    r.m02 = inv_det *  (m.m01 * s23_13 - m.m02 * s13_13 + m.m03 * s12_13);
    r.m03 = inv_det * -(m.m01 * s23_12 - m.m02 * s13_12 + m.m03 * s12_12);
    r.m10 = inv_det * -(m.m10 * s23_23 - m.m12 * s03_23 + m.m13 * s02_23);
    r.m11 = inv_det *  (m.m00 * s23_23 - m.m02 * s03_23 + m.m03 * s02_23);
    r.m12 = inv_det * -(m.m00 * s23_13 - m.m02 * s03_13 + m.m03 * s02_13);
    r.m13 = inv_det *  (m.m00 * s23_12 - m.m02 * s03_12 + m.m03 * s02_12);
    r.m20 = inv_det *  (m.m10 * s13_23 - m.m11 * s03_23 + m.m13 * s01_23);
    r.m21 = inv_det * -(m.m00 * s13_23 - m.m01 * s03_23 + m.m03 * s01_23);
    r.m22 = inv_det *  (m.m00 * s13_13 - m.m01 * s03_13 + m.m03 * s01_13);
    r.m23 = inv_det * -(m.m00 * s13_12 - m.m01 * s03_12 + m.m03 * s01_12);
    r.m30 = inv_det * -(m.m10 * s12_23 - m.m11 * s02_23 + m.m12 * s01_23);
    r.m31 = inv_det *  (m.m00 * s12_23 - m.m01 * s02_23 + m.m02 * s01_23);
    r.m32 = inv_det * -(m.m00 * s12_13 - m.m01 * s02_13 + m.m02 * s01_13);
    r.m33 = inv_det *  (m.m00 * s12_12 - m.m01 * s02_12 + m.m02 * s01_12);
    return r;
}

/*
 * assets/shaders/raytracer/collide.inc:3-24 (== physics/tg_physics.c:394-406 up to min/max form).
 * True division; a zero direction component yields -/+F32_MAX.
 */
TGB_HD b32 tgb_ray_aabb(v3 o, v3 d, v3 bmin, v3 bmax, f32* p_enter, f32* p_exit)
{
    const f32 a_x = (d.x == 0.0f) ? TG_F32_MIN : ((bmin.x - o.x) / d.x);
    const f32 a_y = (d.y == 0.0f) ? TG_F32_MIN : ((bmin.y - o.y) / d.y);
    const f32 a_z = (d.z == 0.0f) ? TG_F32_MIN : ((bmin.z - o.z) / d.z);
    const f32 b_x = (d.x == 0.0f) ? TG_F32_MAX : ((bmax.x - o.x) / d.x);
    const f32 b_y = (d.y == 0.0f) ? TG_F32_MAX : ((bmax.y - o.y) / d.y);
    const f32 b_z = (d.z == 0.0f) ? TG_F32_MAX : ((bmax.z - o.z) / d.z);
    const f32 enter = tgb_max(tgb_max(tgb_min(a_x, b_x), tgb_min(a_y, b_y)), tgb_min(a_z, b_z));
    const f32 exit  = tgb_min(tgb_min(tgb_max(a_x, b_x), tgb_max(a_y, b_y)), tgb_max(a_z, b_z));
    *p_enter = enter;
    *p_exit = exit;
    return exit > 0.0f && enter <= exit;
}

/* math/tg_math.c:328-338 == assets/shaders/util.inc:11-19 */
TGB_HD u32 tgb_xorshift32(u32* p_state)
{
    u32 s = *p_state;
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    *p_state = s;
    return s;
}
/* assets/shaders/util.inc:21-31 */
TGB_HD f32 tgb_xorshift32_range(u32* p_state, f32 lo, f32 hi)
{
    return ((f32)tgb_xorshift32(p_state) / (f32)TG_U32_MAX) * (hi - lo) + lo;
}
/* assets/shaders/util.inc:47-56 */
TGB_HD u32 tgb_hash_u32(u32 v)
{
    v ^= v >> 16;
    v *= 0x85ebca6bu;
    v ^= v >> 13;
    v *= 0xc2b2ae35u;
    v ^= v >> 16;
    return v;
}

#ifndef __CUDACC__
/* ---- host-only: trigonometry from libm ------------------------------------------------------ */

/* math/tg_math.h:9,27 */
#define TGB_PI 3.14159274f
static inline f32 tgb_deg2rad(f32 degrees) { return degrees * ((TGB_PI * 2.0f) / 360.0f); }

/* math/tg_math.c:1870-1910 */
static inline m4 tgb_m4_angle_axis(f32 angle_in_radians, v3 axis)
{
    const f32 c = cosf(angle_in_radians);
    const f32 s = sinf(angle_in_radians);
    const f32 d = 1.0f - c;
    const f32 x = axis.x * d;
    const f32 y = axis.y * d;
    const f32 z = axis.z * d;
    const f32 axay = x * axis.y;
    const f32 axaz = x * axis.z;
    const f32 ayaz = y * axis.z;
    m4 r;
    r.m00 = c + x * axis.x;    r.m10 = axay + s * axis.z; r.m20 = axaz - s * axis.y; r.m30 = 0.0f;
    r.m01 = axay - s * axis.z; r.m11 = c + y * axis.y;    r.m21 = ayaz + s * axis.x; r.m31 = 0.0f;
    r.m02 = axaz + s * axis.y; r.m12 = ayaz - s * axis.x; r.m22 = c + z * axis.z;    r.m32 = 0.0f;
    r.m03 = 0.0f;              r.m13 = 0.0f;              r.m23 = 0.0f;              r.m33 = 1.0f;
    return r;
}

/* math/tg_math.c:2501-2589, 2020-2028: euler = Z * (Y * X) */
static inline m4 tgb_m4_euler(f32 pitch, f32 yaw, f32 roll)
{
    m4 x = tgb_m4_identity(), y = tgb_m4_identity(), z = tgb_m4_identity();
    const f32 cx = cosf(pitch), sx = sinf(pitch);
    const f32 cy = cosf(yaw),   sy = sinf(yaw);
    const f32 cz = cosf(roll),  sz = sinf(roll);
    x.m11 = cx; x.m21 = sx; x.m12 = -sx; x.m22 = cx;
    y.m00 = cy; y.m20 = -sy; y.m02 = sy; y.m22 = cy;
    z.m00 = cz; z.m10 = sz; z.m01 = -sz; z.m11 = cz;
    return tgb_m4_mul(z, tgb_m4_mul(y, x));
}

/* math/tg_math.c:2469-2499 */
static inline m4 tgb_m4_perspective(f32 fov_y_in_radians, f32 aspect, f32 n, f32 f)
{
    const f32 tan_half_fov_y = tanf(fov_y_in_radians / 2.0f);
    m4 r;
    r.m00 = 1.0f / (aspect * tan_half_fov_y); r.m10 = 0.0f; r.m20 = 0.0f; r.m30 = 0.0f;
    r.m01 = 0.0f; r.m11 = -1.0f / tan_half_fov_y; r.m21 = 0.0f; r.m31 = 0.0f;
    r.m02 = 0.0f; r.m12 = 0.0f; r.m22 = f / (n - f); r.m32 = -1.0f;
    r.m03 = 0.0f; r.m13 = 0.0f; r.m23 = -(2.0f * f * n) / (f - n); r.m33 = 0.0f;
    return r;
}
#endif /* !__CUDACC__ */

#endif

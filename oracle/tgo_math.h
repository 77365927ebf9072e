/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY. Nothing under tg_b200/ may include, link or call this.
 *
 * tgo_math.h: scalar float32 restatement of the subset of the reference's math that the voxel
 * rendering path uses. Paths relative to /root/reference/tg/src. Every function keeps the
 * reference's operation ORDER because results are compared bit for bit; build with
 * `-ffp-contract=off -fno-fast-math` (see oracle/Makefile).
 *
 * Pinned conventions where GLSL is implementation-defined:
 *   mix(a,b,t)  = a*(1-t) + b*t            (== tgm_v3_lerp, math/tg_math.c:1091-1098)
 *   inverse(m4) = cofactor form            (tgm_m4_inverse, math/tg_math.c:2185-2234)
 *   normalize   = v / sqrtf(x*x+y*y+z*z)   (tgm_v3_normalized, math/tg_math.c:1106-1110,1175-1184)
 *   min(x,y)    = y < x ? y : x,  max(x,y) = x < y ? y : x   (GLSL 4.50 spec 8.3)
 *   float->int  = C truncation
 */
#ifndef TGO_MATH_H
#define TGO_MATH_H

#include <math.h>
#include "../include/tg_types.h"

static inline f32 tgo_min(f32 x, f32 y) { return y < x ? y : x; }
static inline f32 tgo_max(f32 x, f32 y) { return x < y ? y : x; }
static inline f32 tgo_clamp(f32 x, f32 lo, f32 hi) { return tgo_min(tgo_max(x, lo), hi); }
static inline f32 tgo_mix(f32 a, f32 b, f32 t) { return a * (1.0f - t) + b * t; }
static inline f32 tgo_sign(f32 x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

static inline v3 tgo_v3(f32 x, f32 y, f32 z) { v3 r = { x, y, z }; return r; }
static inline v3 tgo_v3_add(v3 a, v3 b) { return tgo_v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 tgo_v3_sub(v3 a, v3 b) { return tgo_v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 tgo_v3_mul(v3 a, v3 b) { return tgo_v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 tgo_v3_mulf(v3 a, f32 f) { return tgo_v3(a.x * f, a.y * f, a.z * f); }
static inline v3 tgo_v3_neg(v3 a) { return tgo_v3(-a.x, -a.y, -a.z); }
/* math/tg_math.c:1064-1068 */
static inline f32 tgo_v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* math/tg_math.c:1106-1110 */
static inline f32 tgo_v3_mag(v3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }
/* math/tg_math.c:1175-1184 */
static inline v3 tgo_v3_normalized(v3 v)
{
    const f32 mag = tgo_v3_mag(v);
    return tgo_v3(v.x / mag, v.y / mag, v.z / mag);
}
/* math/tg_math.c:1053-1062 */
static inline v3 tgo_v3_divf(v3 v, f32 f) { return tgo_v3(v.x / f, v.y / f, v.z / f); }
/* math/tg_math.c:1118-1140: C ternaries (NOT the GLSL forms) */
static inline v3 tgo_v3_max(v3 a, v3 b) { return tgo_v3(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z); }
static inline v3 tgo_v3_min(v3 a, v3 b) { return tgo_v3(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z); }
static inline v3 tgo_v3_floor(v3 a) { return tgo_v3(floorf(a.x), floorf(a.y), floorf(a.z)); }
static inline v3 tgo_v3_ceil(v3 a) { return tgo_v3(ceilf(a.x), ceilf(a.y), ceilf(a.z)); }
/* math/tg_math.c:1091-1098 */
static inline v3 tgo_v3_lerp(v3 a, v3 b, f32 t)
{
    return tgo_v3((1.0f - t) * a.x + t * b.x, (1.0f - t) * a.y + t * b.y, (1.0f - t) * a.z + t * b.z);
}
/* GLSL mix() on vec3 with the pinned scalar form */
static inline v3 tgo_v3_mix(v3 a, v3 b, f32 t) { return tgo_v3(tgo_mix(a.x, b.x, t), tgo_mix(a.y, b.y, t), tgo_mix(a.z, b.z, t)); }

static inline m4 tgo_m4_identity(void)
{
    m4 r = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    return r;
}

/* math/tg_math.c:2669-2694 */
static inline m4 tgo_m4_translate(v3 v)
{
    m4 r = tgo_m4_identity();
    r.m03 = v.x;
    r.m13 = v.y;
    r.m23 = v.z;
    return r;
}

/* math/tg_math.c:2374-2399: every entry is the left-to-right sum of four products */
static inline m4 tgo_m4_mul(m4 a, m4 b)
{
    m4 r;
    r.m00 = a.m00 * b.m00 + a.m01 * b.m10 + a.m02 * b.m20 + a.m03 * b.m30;
    r.m10 = a.m10 * b.m00 + a.m11 * b.m10 + a.m12 * b.m20 + a.m13 * b.m30;
    r.m20 = a.m20 * b.m00 + a.m21 * b.m10 + a.m22 * b.m20 + a.m23 * b.m30;
    r.m30 = a.m30 * b.m00 + a.m31 * b.m10 + a.m32 * b.m20 + a.m33 * b.m30;

    r.m01 = a.m00 * b.m01 + a.m01 * b.m11 + a.m02 * b.m21 + a.m03 * b.m31;
    r.m11 = a.m10 * b.m01 + a.m11 * b.m11 + a.m12 * b.m21 + a.m13 * b.m31;
    r.m21 = a.m20 * b.m01 + a.m21 * b.m11 + a.m22 * b.m21 + a.m23 * b.m31;
    r.m31 = a.m30 * b.m01 + a.m31 * b.m11 + a.m32 * b.m21 + a.m33 * b.m31;

    r.m02 = a.m00 * b.m02 + a.m01 * b.m12 + a.m02 * b.m22 + a.m03 * b.m32;
    r.m12 = a.m10 * b.m02 + a.m11 * b.m12 + a.m12 * b.m22 + a.m13 * b.m32;
    r.m22 = a.m20 * b.m02 + a.m21 * b.m12 + a.m22 * b.m22 + a.m23 * b.m32;
    r.m32 = a.m30 * b.m02 + a.m31 * b.m12 + a.m32 * b.m22 + a.m33 * b.m32;

    r.m03 = a.m00 * b.m03 + a.m01 * b.m13 + a.m02 * b.m23 + a.m03 * b.m33;
    r.m13 = a.m10 * b.m03 + a.m11 * b.m13 + a.m12 * b.m23 + a.m13 * b.m33;
    r.m23 = a.m20 * b.m03 + a.m21 * b.m13 + a.m22 * b.m23 + a.m23 * b.m33;
    r.m33 = a.m30 * b.m03 + a.m31 * b.m13 + a.m32 * b.m23 + a.m33 * b.m33;
    return r;
}

/* math/tg_math.c:2428-2438 */
static inline v4 tgo_m4_mulv4(m4 m, v4 v)
{
    v4 r;
    r.x = v.x * m.m00 + v.y * m.m01 + v.z * m.m02 + v.w * m.m03;
    r.y = v.x * m.m10 + v.y * m.m11 + v.z * m.m12 + v.w * m.m13;
    r.z = v.x * m.m20 + v.y * m.m21 + v.z * m.m22 + v.w * m.m23;
    r.w = v.x * m.m30 + v.y * m.m31 + v.z * m.m32 + v.w * m.m33;
    return r;
}
static inline v3 tgo_m4_mulv3w(m4 m, v3 v, f32 w)
{
    v4 q = { v.x, v.y, v.z, w };
    q = tgo_m4_mulv4(m, q);
    return tgo_v3(q.x, q.y, q.z);
}

/* math/tg_math.c:2185-2234 (cofactor expansion, det = 1/determinant) */
static inline m4 tgo_m4_inverse(m4 m)
{
    m4 r;
    const f32 m2323 = m.m22 * m.m33 - m.m23 * m.m32;
    const f32 m1323 = m.m21 * m.m33 - m.m23 * m.m31;
    const f32 m1223 = m.m21 * m.m32 - m.m22 * m.m31;
    const f32 m0323 = m.m20 * m.m33 - m.m23 * m.m30;
    const f32 m0223 = m.m20 * m.m32 - m.m22 * m.m30;
    const f32 m0123 = m.m20 * m.m31 - m.m21 * m.m30;
    const f32 m2313 = m.m12 * m.m33 - m.m13 * m.m32;
    const f32 m1313 = m.m11 * m.m33 - m.m13 * m.m31;
    const f32 m1213 = m.m11 * m.m32 - m.m12 * m.m31;
    const f32 m2312 = m.m12 * m.m23 - m.m13 * m.m22;
    const f32 m1312 = m.m11 * m.m23 - m.m13 * m.m21;
    const f32 m1212 = m.m11 * m.m22 - m.m12 * m.m21;
    const f32 m0313 = m.m10 * m.m33 - m.m13 * m.m30;
    const f32 m0213 = m.m10 * m.m32 - m.m12 * m.m30;
    const f32 m0312 = m.m10 * m.m23 - m.m13 * m.m20;
    const f32 m0212 = m.m10 * m.m22 - m.m12 * m.m20;
    const f32 m0113 = m.m10 * m.m31 - m.m11 * m.m30;
    const f32 m0112 = m.m10 * m.m21 - m.m11 * m.m20;

    const f32 det = 1.0f / (
        m.m00 * (m.m11 * m2323 - m.m12 * m1323 + m.m13 * m1223) -
        m.m01 * (m.m10 * m2323 - m.m12 * m0323 + m.m13 * m0223) +
        m.m02 * (m.m10 * m1323 - m.m11 * m0323 + m.m13 * m0123) -
        m.m03 * (m.m10 * m1223 - m.m11 * m0223 + m.m12 * m0123));

    r.m00 = det *  (m.m11 * m2323 - m.m12 * m1323 + m.m13 * m1223);
    r.m01 = det * -(m.m01 * m2323 - m.m02 * m1323 + m.m03 * m1223);
    r.m02 = det *  (m.m01 * m2313 - m.m02 * m1313 + m.m03 * m1213);
    r.m03 = det * -(m.m01 * m2312 - m.m02 * m1312 + m.m03 * m1212);
    r.m10 = det * -(m.m10 * m2323 - m.m12 * m0323 + m.m13 * m0223);
    r.m11 = det *  (m.m00 * m2323 - m.m02 * m0323 + m.m03 * m0223);
    r.m12 = det * -(m.m00 * m2313 - m.m02 * m0313 + m.m03 * m0213);
    r.m13 = det *  (m.m00 * m2312 - m.m02 * m0312 + m.m03 * m0212);
    r.m20 = det *  (m.m10 * m1323 - m.m11 * m0323 + m.m13 * m0123);
    r.m21 = det * -(m.m00 * m1323 - m.m01 * m0323 + m.m03 * m0123);
    r.m22 = det *  (m.m00 * m1313 - m.m01 * m0313 + m.m03 * m0113);
    r.m23 = det * -(m.m00 * m1312 - m.m01 * m0312 + m.m03 * m0112);
    r.m30 = det * -(m.m10 * m1223 - m.m11 * m0223 + m.m12 * m0123);
    r.m31 = det *  (m.m00 * m1223 - m.m01 * m0223 + m.m02 * m0123);
    r.m32 = det * -(m.m00 * m1213 - m.m01 * m0213 + m.m02 * m0113);
    r.m33 = det *  (m.m00 * m1212 - m.m01 * m0212 + m.m02 * m0112);
    return r;
}

/* math/tg_math.c:1870-1910 (Lengyel); sinf/cosf from libm, HOST ONLY */
static inline m4 tgo_m4_angle_axis(f32 angle_in_radians, v3 axis)
{
    m4 r;
    const f32 c = cosf(angle_in_radians);
    const f32 s = sinf(angle_in_radians);
    const f32 d = 1.0f - c;

    const f32 x = axis.x * d;
    const f32 y = axis.y * d;
    const f32 z = axis.z * d;
    const f32 axay = x * axis.y;
    const f32 axaz = x * axis.z;
    const f32 ayaz = y * axis.z;

    r.m00 = c + x * axis.x;
    r.m10 = axay + s * axis.z;
    r.m20 = axaz - s * axis.y;
    r.m30 = 0.0f;

    r.m01 = axay - s * axis.z;
    r.m11 = c + y * axis.y;
    r.m21 = ayaz + s * axis.x;
    r.m31 = 0.0f;

    r.m02 = axaz + s * axis.y;
    r.m12 = ayaz - s * axis.x;
    r.m22 = c + z * axis.z;
    r.m32 = 0.0f;

    r.m03 = 0.0f;
    r.m13 = 0.0f;
    r.m23 = 0.0f;
    r.m33 = 1.0f;
    return r;
}

/* math/tg_math.c:2501-2589 */
static inline m4 tgo_m4_rotate_x(f32 a)
{
    m4 r = tgo_m4_identity();
    const f32 c = cosf(a), s = sinf(a);
    r.m11 = c;  r.m21 = s;
    r.m12 = -s; r.m22 = c;
    return r;
}
static inline m4 tgo_m4_rotate_y(f32 a)
{
    m4 r = tgo_m4_identity();
    const f32 c = cosf(a), s = sinf(a);
    r.m00 = c; r.m20 = -s;
    r.m02 = s; r.m22 = c;
    return r;
}
static inline m4 tgo_m4_rotate_z(f32 a)
{
    m4 r = tgo_m4_identity();
    const f32 c = cosf(a), s = sinf(a);
    r.m00 = c;  r.m10 = s;
    r.m01 = -s; r.m11 = c;
    return r;
}
/* math/tg_math.c:2020-2028: Z * (Y * X) */
static inline m4 tgo_m4_euler(f32 pitch, f32 yaw, f32 roll)
{
    const m4 x = tgo_m4_rotate_x(pitch);
    const m4 y = tgo_m4_rotate_y(yaw);
    const m4 z = tgo_m4_rotate_z(roll);
    return tgo_m4_mul(z, tgo_m4_mul(y, x));
}

/* math/tg_math.c:2469-2499 */
static inline m4 tgo_m4_perspective(f32 fov_y_in_radians, f32 aspect, f32 n, f32 f)
{
    m4 r = { 0 };
    const f32 tan_half_fov_y = tanf(fov_y_in_radians / 2.0f);
    const f32 a = f / (n - f);
    const f32 b = -(2.0f * f * n) / (f - n);
    r.m00 = 1.0f / (aspect * tan_half_fov_y);
    r.m11 = -1.0f / tan_half_fov_y;
    r.m22 = a;
    r.m32 = -1.0f;
    r.m23 = b;
    return r;
}

/* math/tg_math.c:328-338 == shaders/util.inc:11-19 */
static inline u32 tgo_xorshift32_next(u32* p_state)
{
    u32 r = *p_state;
    r ^= r << 13;
    r ^= r >> 17;
    r ^= r << 5;
    *p_state = r;
    return r;
}
/* shaders/util.inc:21-31 */
static inline f32 tgo_xorshift32_next_f32(u32* p_state) { return (f32)tgo_xorshift32_next(p_state) / (f32)TG_U32_MAX; }
static inline f32 tgo_xorshift32_next_f32_range(u32* p_state, f32 lo, f32 hi) { return tgo_xorshift32_next_f32(p_state) * (hi - lo) + lo; }
/* shaders/util.inc:47-56 (murmur3 finaliser) */
static inline u32 tgo_hash_u32(u32 v)
{
    u32 r = v;
    r ^= r >> 16;
    r *= 0x85ebca6bu;
    r ^= r >> 13;
    r *= 0xc2b2ae35u;
    r ^= r >> 16;
    return r;
}

/*
 * shaders/raytracer/collide.inc:3-24 (GLSL; the visibility / shading / SVO-traversal shaders).
 * True division; zero direction component -> -/+F32_MAX; GLSL min/max.
 */
static inline b32 tgo_intersect_ray_aabb_glsl(v3 o, v3 d, v3 bmin, v3 bmax, f32* p_enter, f32* p_exit)
{
    const f32 vec0_x = (d.x == 0.0f) ? TG_F32_MIN : ((bmin.x - o.x) / d.x);
    const f32 vec0_y = (d.y == 0.0f) ? TG_F32_MIN : ((bmin.y - o.y) / d.y);
    const f32 vec0_z = (d.z == 0.0f) ? TG_F32_MIN : ((bmin.z - o.z) / d.z);

    const f32 vec1_x = (d.x == 0.0f) ? TG_F32_MAX : ((bmax.x - o.x) / d.x);
    const f32 vec1_y = (d.y == 0.0f) ? TG_F32_MAX : ((bmax.y - o.y) / d.y);
    const f32 vec1_z = (d.z == 0.0f) ? TG_F32_MAX : ((bmax.z - o.z) / d.z);

    const f32 n_x = tgo_min(vec0_x, vec1_x);
    const f32 n_y = tgo_min(vec0_y, vec1_y);
    const f32 n_z = tgo_min(vec0_z, vec1_z);

    const f32 f_x = tgo_max(vec0_x, vec1_x);
    const f32 f_y = tgo_max(vec0_y, vec1_y);
    const f32 f_z = tgo_max(vec0_z, vec1_z);

    *p_enter = tgo_max(tgo_max(n_x, n_y), n_z);
    *p_exit  = tgo_min(tgo_min(f_x, f_y), f_z);
    return *p_exit > 0.0f && *p_enter <= *p_exit;
}

/* physics/tg_physics.c:394-406 (C twin; TG_MAX/TG_MIN macros = C ternaries, math/tg_math.h:17-22) */
static inline b32 tgo_intersect_ray_aabb_c(v3 o, v3 d, v3 bmin, v3 bmax, f32* p_enter, f32* p_exit)
{
    v3 vec0, vec1;
    vec0.x = (d.x == 0.0f) ? TG_F32_MIN : ((bmin.x - o.x) / d.x);
    vec0.y = (d.y == 0.0f) ? TG_F32_MIN : ((bmin.y - o.y) / d.y);
    vec0.z = (d.z == 0.0f) ? TG_F32_MIN : ((bmin.z - o.z) / d.z);
    vec1.x = (d.x == 0.0f) ? TG_F32_MAX : ((bmax.x - o.x) / d.x);
    vec1.y = (d.y == 0.0f) ? TG_F32_MAX : ((bmax.y - o.y) / d.y);
    vec1.z = (d.z == 0.0f) ? TG_F32_MAX : ((bmax.z - o.z) / d.z);
    const v3 n = tgo_v3_min(vec0, vec1);
    const v3 f = tgo_v3_max(vec0, vec1);
#define TGO_CMAX(a, b) ((a) > (b) ? (a) : (b))
#define TGO_CMIN(a, b) ((a) < (b) ? (a) : (b))
    *p_enter = TGO_CMAX(TGO_CMAX(n.x, n.y), n.z);
    *p_exit  = TGO_CMIN(TGO_CMIN(f.x, f.y), f.z);
#undef TGO_CMAX
#undef TGO_CMIN
    return *p_exit > 0.0f && *p_enter <= *p_exit;
}

#endif

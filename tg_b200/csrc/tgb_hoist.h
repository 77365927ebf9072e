/*
 * tgb_hoist.h -- per-object factorisation of the reference's world-space -> cluster-space chain.
 *
 * visibility.frag:57-69 (and shading.frag:169-181, tg_sparse_voxel_octree.c:53-57) rebuild, PER
 * FRAGMENT,   ws2ms = T(-off) * T(half) * inverse(rotation) * T(-translation)   (left-associative
 * tgm_m4_mul products, math/tg_math.c:2374-2399) and then  o_ms = ws2ms * (camera, 1),
 * d_ms = normalize(ws2ms * (dir_ws, 0)).  With A = T(-off)*T(half), B = A*R, C = B*T(-t), writing
 * the four-term sums out shows that only the translation column depends on the cluster:
 *
 *   A      = [ I | a ],  a = half - off                      (small integers: exact)
 *   B.mi3  = ((A.mi0*R.m03 + A.mi1*R.m13) + A.mi2*R.m23) + a_i*R.m33       = p_i + a_i*r33
 *   C.mi3  = ((B.mi0*-tx + B.mi1*-ty) + B.mi2*-tz) + B.mi3*1               = q_i + B.mi3
 *   o_ms_i = ((cam.x*C.mi0 + cam.y*C.mi1) + cam.z*C.mi2) + 1*C.mi3         = oc_i + C.mi3
 *
 * and the 3x3 part C.mij (j<3) is the same for every cluster of the object up to the sign of an
 * exact zero (terms a_i*R.m3j and B.mi3*0 are +-0 because rotation's last row is (0,0,0,1)); a
 * signed zero never changes a packed word (it only feeds ==0 / <0 / >0 tests and sums).
 * So: run the reference chain ONCE per object for cluster 0 with the generic routines, keep
 * (c, p, q, oc, r33), and per cluster evaluate three mul + nine add with the reference's own
 * operation order. tests/test_host_logic.py::test_per_object_factorisation_equals_the_full_chain checks this against the oracle's full chain bit for bit.
 */
#ifndef TGB_HOIST_H
#define TGB_HOIST_H

#include "tgb_math.h"

/*
 * Per-frame record of one object. First block: exact chain factors. Second block: conservative
 * culling data filled by k_cull_objects (tgb_visibility.cu).
 */
typedef struct tgb_object_frame
{
    f32 c[9];       /* C.mij, row-major [i*3+j] */
    f32 p[3];
    f32 q[3];
    f32 oc[3];
    f32 half[3];    /* 4 * n_cluster_pointers_per_dim */
    f32 r33;        /* inverse(rotation).m33 */
    u32 nx, ny, nz;
    u32 first_cluster_pointer;
    /* conservative */
    f32 og[3];      /* camera in the object's grid frame (== o_ms of cluster 0) */
    f32 eps;        /* candidate-enumeration slack, voxels */
    i32 x0, y0, x1, y1; /* inclusive pixel rectangle */
    u32 min_depth24;    /* lower bound of any depth24 this object can write */
    u32 object_idx;
} tgb_object_frame;

TGB_HD void tgb_hoist_object(const tg_object_data* p_object, v3 camera, tgb_object_frame* f)
{
    const v3 dims = tgb_v3((f32)p_object->n_cluster_pointers_per_dim.x, (f32)p_object->n_cluster_pointers_per_dim.y, (f32)p_object->n_cluster_pointers_per_dim.z);
    /* visibility.frag:45-48 */
    const v3 cluster_half_extent = tgb_mul(tgb_v3(8.0f, 8.0f, 8.0f), tgb_v3(0.5f, 0.5f, 0.5f));
    const v3 half = tgb_mul(cluster_half_extent, dims);
    const v3 t = p_object->translation;

    const m4 R  = tgb_m4_inverse(p_object->rotation);
    const m4 A0 = tgb_m4_mul(tgb_m4_translate(tgb_neg(tgb_v3(0.0f, 0.0f, 0.0f))), tgb_m4_translate(half));
    const m4 B0 = tgb_m4_mul(A0, R);
    const m4 C0 = tgb_m4_mul(B0, tgb_m4_translate(tgb_neg(t)));

    f->c[0] = C0.m00; f->c[1] = C0.m01; f->c[2] = C0.m02;
    f->c[3] = C0.m10; f->c[4] = C0.m11; f->c[5] = C0.m12;
    f->c[6] = C0.m20; f->c[7] = C0.m21; f->c[8] = C0.m22;

    f->p[0] = (A0.m00 * R.m03 + A0.m01 * R.m13) + A0.m02 * R.m23;
    f->p[1] = (A0.m10 * R.m03 + A0.m11 * R.m13) + A0.m12 * R.m23;
    f->p[2] = (A0.m20 * R.m03 + A0.m21 * R.m13) + A0.m22 * R.m23;
    f->r33 = R.m33;

    const f32 ntx = -t.x, nty = -t.y, ntz = -t.z;
    f->q[0] = (B0.m00 * ntx + B0.m01 * nty) + B0.m02 * ntz;
    f->q[1] = (B0.m10 * ntx + B0.m11 * nty) + B0.m12 * ntz;
    f->q[2] = (B0.m20 * ntx + B0.m21 * nty) + B0.m22 * ntz;

    f->oc[0] = (camera.x * C0.m00 + camera.y * C0.m01) + camera.z * C0.m02;
    f->oc[1] = (camera.x * C0.m10 + camera.y * C0.m11) + camera.z * C0.m12;
    f->oc[2] = (camera.x * C0.m20 + camera.y * C0.m21) + camera.z * C0.m22;

    f->half[0] = half.x; f->half[1] = half.y; f->half[2] = half.z;
    f->nx = p_object->n_cluster_pointers_per_dim.x;
    f->ny = p_object->n_cluster_pointers_per_dim.y;
    f->nz = p_object->n_cluster_pointers_per_dim.z;
    f->first_cluster_pointer = p_object->first_cluster_pointer;
}

/* o_ms of the cluster at grid coordinates (cx, cy, cz) of the object */
TGB_HD v3 tgb_hoist_cluster_origin(const tgb_object_frame* f, u32 cx, u32 cy, u32 cz)
{
    const f32 ax = f->half[0] - (f32)(cx * 8u);
    const f32 ay = f->half[1] - (f32)(cy * 8u);
    const f32 az = f->half[2] - (f32)(cz * 8u);
    v3 o;
    o.x = f->oc[0] + (f->q[0] + (f->p[0] + ax * f->r33));
    o.y = f->oc[1] + (f->q[1] + (f->p[1] + ay * f->r33));
    o.z = f->oc[2] + (f->q[2] + (f->p[2] + az * f->r33));
    return o;
}

/* d_ms = normalize(ws2ms * (dir_ws, 0)) */
TGB_HD v3 tgb_hoist_direction(const tgb_object_frame* f, v3 dir_ws)
{
    v3 raw;
    raw.x = (dir_ws.x * f->c[0] + dir_ws.y * f->c[1]) + dir_ws.z * f->c[2];
    raw.y = (dir_ws.x * f->c[3] + dir_ws.y * f->c[4]) + dir_ws.z * f->c[5];
    raw.z = (dir_ws.x * f->c[6] + dir_ws.y * f->c[7]) + dir_ws.z * f->c[8];
    return tgb_normalize(raw);
}

/* common.inc:40-46 with gl_FragCoord = pixel centre (visibility.frag:32-33) */
TGB_HD v3 tgb_pixel_direction(const tg_camera_rays* p_cam, u32 w, u32 h, u32 px, u32 py)
{
    const f32 fx =        ((f32)px + 0.5f) / (f32)w;
    const f32 fy = 1.0f - ((f32)py + 0.5f) / (f32)h;
    const v3 bl = tgb_v3(p_cam->ray_bl.x, p_cam->ray_bl.y, p_cam->ray_bl.z);
    const v3 br = tgb_v3(p_cam->ray_br.x, p_cam->ray_br.y, p_cam->ray_br.z);
    const v3 tr = tgb_v3(p_cam->ray_tr.x, p_cam->ray_tr.y, p_cam->ray_tr.z);
    const v3 tl = tgb_v3(p_cam->ray_tl.x, p_cam->ray_tl.y, p_cam->ray_tl.z);
    return tgb_mix3(tgb_mix3(bl, tl, fy), tgb_mix3(br, tr, fy), fx);
}

#endif

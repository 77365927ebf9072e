/*
 * tgb_shade.cu -- K3, the shading pass with the composed 1-bounce GI term.
 *
 * Replaces the reference's full-screen fragment pass
 *   shading            assets/shaders/raytracer/shading.frag:114-337 (BRDF helpers :53-110)
 *   SVO traversal      assets/shaders/raytracer/svo_functions.inc:1-329 (bound by shading.frag:40-43 but never
 *                      called there; the only GPU caller is debug_visibility_svo.frag:52)
 *   secondary rays     tgvk_raytracer.c:1405-1431 (CPU sketch under `if (0)`), TODO.h:33-43, util.inc:1-56 (RNG, hash)
 * One thread per pixel: unpack the visibility word, fetch material + object, re-derive the primary ray in the winning
 * cluster's space with the per-object factorisation (tgb_hoist.h, once per object per frame in k_object_frames
 * instead of once per fragment), face normal, Cook-Torrance term, then ONE secondary ray per hit pixel through the
 * replicated 1-bit SVO; a miss keeps the ambient term, a hit removes it (GI spec pinned in DESIGN.md, oracle twin:
 * oracle/tgo_shade.c). Floating point: GI radiance tolerance is 1e-3 relative (BASELINE.json), the arithmetic below
 * nevertheless follows the shader's operation order (this TU is built with -fmad=false like the others).
 */
#include "tgb_device.cuh"
#include "tgb_gi_fast.cuh"

#define TGB_PI_F               3.14159265358979323846f

/* ---- per-object frames for ALL objects, indexed by object idx -------------------------------- */
__global__ void k_object_frames(const tg_object_data* __restrict__ p_objects, u32 object_capacity, v3 camera, tgb_object_frame* __restrict__ p_frames)
{
    const u32 object_idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (object_idx >= object_capacity) return;
    const tg_object_data o = p_objects[object_idx];
    if (o.n_cluster_pointers_per_dim.x == 0 || o.n_cluster_pointers_per_dim.y == 0 || o.n_cluster_pointers_per_dim.z == 0) return;
    tgb_object_frame f;
    tgb_hoist_object(&o, camera, &f);
    f.object_idx = object_idx;
    f.og[0] = f.og[1] = f.og[2] = 0.0f; f.eps = 0.0f; f.x0 = f.y0 = f.x1 = f.y1 = 0; f.min_depth24 = 0;
    p_frames[object_idx] = f;
}

/* ---- BRDF: shading.frag:53-110 ------------------------------------------------------------------ */
__device__ __forceinline__ f32 tgb_distribution_ggx(f32 clamped_n_dot_h, f32 roughness)
{
    const f32 a = roughness * roughness;
    const f32 a_sqr = a * a;
    f32 denom = (clamped_n_dot_h * clamped_n_dot_h * (a_sqr - 1.0f) + 1.0f);
    denom = TGB_PI_F * denom * denom;
    return a_sqr / denom;
}
__device__ __forceinline__ f32 tgb_geometry_schlick_ggx(f32 n_dot_v, f32 roughness)
{
    const f32 r = (roughness + 1.0f);
    const f32 k = (r * r) / 8.0f;
    const f32 denom = n_dot_v * (1.0f - k) + k;
    return n_dot_v / denom;
}
__device__ __forceinline__ f32 tgb_pow5(f32 x) { return ((x * x) * (x * x)) * x; } /* pow(x, 5.0) pinned, oracle/tgo_shade.c */
__device__ __forceinline__ f32 tgb_fresnel_schlick(f32 u, f32 roughness)
{
    const f32 f_lambda = 1.0f - roughness;
    return f_lambda + (1.0f - f_lambda) * tgb_pow5(1.0f - u);
}
__device__ __forceinline__ v3 tgb_shade_brdf(v3 n, v3 v, v3 l, v3 diffuse_albedo, v3 specular_albedo, f32 metallic, f32 roughness, v3 radiance)
{
    const v3 h = tgb_normalize(tgb_add(v, l));
    const f32 clamped_h_dot_n = tgb_clamp(tgb_dot(h, n), 0.0f, 1.0f);
    const f32 clamped_l_dot_n = tgb_clamp(tgb_dot(n, l), 0.0f, 1.0f);
    const f32 clamped_n_dot_v = tgb_clamp(tgb_dot(n, v), 0.0f, 1.0f);
    const f32 d = tgb_distribution_ggx(clamped_h_dot_n, roughness);
    const f32 f = tgb_fresnel_schlick(clamped_n_dot_v, roughness);
    const f32 g = tgb_geometry_schlick_ggx(clamped_l_dot_n, roughness) * tgb_geometry_schlick_ggx(clamped_n_dot_v, roughness);
    const f32 dfg = d * f * g;
    const f32 denominator = 4.0f * clamped_n_dot_v * clamped_l_dot_n;
    const v3 specular = tgb_scale(specular_albedo, dfg / tgb_max(denominator, 0.001f));
    const f32 kd = (1.0f - f) * (1.0f - metallic);
    const v3 diffuse = tgb_divf(tgb_scale(diffuse_albedo, kd), TGB_PI_F);
    return tgb_scale(tgb_mul(tgb_add(diffuse, specular), radiance), clamped_l_dot_n);
}

__device__ __forceinline__ float4 tgb_hash_color(u32 v)
{
    const u32 h0 = tgb_hash_u32(v), h1 = tgb_hash_u32(h0), h2 = tgb_hash_u32(h1);
    return make_float4((f32)h0 / 4294967295.0f, (f32)h1 / 4294967295.0f, (f32)h2 / 4294967295.0f, 1.0f);
}

/* ---- SVO traversal: svo_functions.inc:1-329 ---------------------------------------------------- */
struct tgb_svo_view
{
    const u32* __restrict__ p_nodes;
    const u32* __restrict__ p_leaf_data; /* 65 u32 per leaf: n, cluster_idcs[64] */
    const u32* __restrict__ p_voxels;    /* 1024 u32 per leaf */
    v3 bmin, bmax;
};

/* ---- K3a: per-pixel shading, secondary rays that enter the SVO box are queued ---------------------- */
struct tgb_shade_args
{
    const u64* __restrict__ p_vis;
    float4* __restrict__ p_out;
    const u32* __restrict__ p_cluster_pointers;
    const u32* __restrict__ p_c2o;
    const tg_object_data* __restrict__ p_objects;
    const tgb_object_frame* __restrict__ p_frames; /* by object idx */
    const u8* __restrict__ p_lut_idx;
    const u32* __restrict__ p_color_lut;
    tgb_svo_view svo;
    tg_camera_rays cam;
    u32 w, h;
    u32 global_pointer_base, n_local_pointers;
    u32 gi_enabled, frame_seed, debug_visualization;
    u32 y0, y1; /* VIRTUAL rows [y0, y1) are shaded by this launch (a band of this rank's screen tile; tgb_rows.h) */
    u32 mat_y0; /* first virtual row of the tile p_mat describes */
    u32 n_ranks, tile_rows; /* virtual <-> physical row mapping */
    /* GI ray queue (SoA): origin.xyz + pixel | direction.xyz + enter of the root slab test | ambient.rgb */
    float4* __restrict__ p_q0;
    float4* __restrict__ p_q1;
    float4* __restrict__ p_q2;
    u32* __restrict__ p_q_count; /* [0] rays queued, [1] rays fetched */
    const u64* __restrict__ p_mat; /* RESOLVED mode: this tile's owner-resolved material words, row y0 first */
    /* FAST: the certified walk over the coarser tiling (tgb_gi_fast.cuh) takes its first steps here, and only rays still undecided are queued */
    tgb_gi_frame fast_frame;
    tgb_fast_tiling fast_tiling;
    u32 fast_steps;
    f32 fast_delta;
};

/* returns true when a secondary ray has to be traced; *p_color is then the pixel WITHOUT its ambient term */
/*
 * RESOLVED (multi-GPU): the winning cluster may live on another GPU, so its material arrived as a word resolved by
 * the owner ((global object idx + 1) << 32 | packed colour) and the object tables (a.p_objects, a.p_frames) are the
 * all-gathered global ones whose first_cluster_pointer is global. Everything downstream is the same arithmetic.
 */
template <bool RESOLVED>
__device__ __forceinline__ bool tgb_shade_pixel(const tgb_shade_args& a, u32 px, u32 py, u32 vy, float4* p_color, v3* p_origin, v3* p_dir, v3* p_ambient, f32* p_root_enter)
{
    const u64 pixel = (u64)vy * a.w + px; /* buffers keep rows in virtual order; the ray and the RNG seed use the physical row py */

    /* shading.frag:116-120 */
    const u64 packed_data = a.p_vis[pixel];
    const f32 depth_24b           = (f32)(u32)(packed_data >> TG_VIS_DEPTH_SHIFT) / TG_VIS_DEPTH_SCALE;
    const u32 cluster_pointer_31b = (u32)(packed_data >> TG_VIS_POINTER_SHIFT) & 2147483647u;
    const u32 voxel_idx_9b        = (u32)(packed_data) & 511u;

    if (!(depth_24b < 1.0f)) { *p_color = make_float4(1.0f, 0.0f, 1.0f, 1.0f); return false; } /* :335 */

    if (a.debug_visualization == TG_DEBUG_SHOW_BLOCKS)
    {
        /* The word comes from the SVO primary-ray pass (tgb_debug_svo.cu; debug_visibility_svo.frag): its pointer field is a leaf NODE
         * index, which shading.frag:122-126,247-256 runs through the cluster-pointer table and hashes; what the shader computes in
         * between is dead for this view. A node index beyond the live pointer range reads as cluster 0 (the reference reads whatever
         * its SSBO holds there). Sharded: the pointer table is distributed, the node index itself is hashed (documented in
         * tg_raytracer.h). */
        const u32 cluster_idx_of_node = RESOLVED ? cluster_pointer_31b : (cluster_pointer_31b < a.n_local_pointers ? __ldg(&a.p_cluster_pointers[cluster_pointer_31b]) : 0u);
        *p_color = tgb_hash_color(cluster_idx_of_node);
        return false;
    }

    u32 local_pointer, cluster_idx, object_idx, color_lut_idx, packed_color;
    if (RESOLVED)
    {
        const u64 mat = a.p_mat[(u64)(vy - a.mat_y0) * a.w + px];
        if (mat == 0) { *p_color = make_float4(0.0f, 0.0f, 0.0f, 0.0f); return false; } /* no rank owns this pointer: inconsistent shards */
        local_pointer = cluster_pointer_31b; /* global pointer against globalised object records */
        cluster_idx = cluster_pointer_31b;   /* debug views only */
        object_idx = (u32)(mat >> 32) - 1u;
        color_lut_idx = 0;                   /* debug views only */
        packed_color = (u32)mat;
    }
    else
    {
        local_pointer = cluster_pointer_31b - a.global_pointer_base;
        if (local_pointer >= a.n_local_pointers) { *p_color = make_float4(0.0f, 0.0f, 0.0f, 0.0f); return false; } /* another shard's cluster */
        cluster_idx = __ldg(&a.p_cluster_pointers[local_pointer]);
        object_idx = __ldg(&a.p_c2o[cluster_idx]);
        /* :128-135; per-object LUT (Q2) */
        color_lut_idx = __ldg(&a.p_lut_idx[(u64)cluster_idx * 512u + voxel_idx_9b]);
        packed_color = __ldg(&a.p_color_lut[a.p_objects[object_idx].lut_idx * 256u + color_lut_idx]);
    }
    const tgb_object_frame& f = a.p_frames[object_idx];
    const tg_object_data& obj = a.p_objects[object_idx];
    const f32 color_r = (f32)( packed_color >> 24        ) / 255.0f;
    const f32 color_g = (f32)((packed_color >> 16) & 0xffu) / 255.0f;
    const f32 color_b = (f32)((packed_color >>  8) & 0xffu) / 255.0f;

    /* :144-181: the primary ray in the winning cluster's space */
    const u32 rel = local_pointer - f.first_cluster_pointer;
    const u32 cx = rel % f.nx, cy = (rel / f.nx) % f.ny, cz = rel / (f.nx * f.ny);
    const v3 ray_origin_ws = tgb_v3(a.cam.camera.x, a.cam.camera.y, a.cam.camera.z);
    const v3 ray_origin_ms = tgb_hoist_cluster_origin(&f, cx, cy, cz);
    const v3 ray_direction_ws = tgb_pixel_direction(&a.cam, a.w, a.h, px, py);
    const v3 ray_direction_ms = tgb_hoist_direction(&f, ray_direction_ws);

    /* :183-228 */
    const v3 voxel_min = tgb_v3((f32)(voxel_idx_9b % 8u), (f32)((voxel_idx_9b / 8u) % 8u), (f32)(voxel_idx_9b / 64u));
    const v3 voxel_max = tgb_add(voxel_min, tgb_v3(1.0f, 1.0f, 1.0f));
    v3 normal_ws = tgb_v3(0.0f, 0.0f, 0.0f);
    f32 enter, exit;
    if (tgb_ray_aabb(ray_origin_ms, ray_direction_ms, voxel_min, voxel_max, &enter, &exit))
    {
        const v3 hit_position_ms = enter > 0.0f ? tgb_add(ray_origin_ms, tgb_scale(ray_direction_ms, enter)) : ray_origin_ms;
        const v3 voxel_center_ms = tgb_add(voxel_min, tgb_v3(0.5f, 0.5f, 0.5f));
        v3 n = tgb_sub(hit_position_ms, voxel_center_ms);
        if (fabsf(n.x) > fabsf(n.y))
        {
            n.y = 0.0f;
            if (fabsf(n.x) > fabsf(n.z)) { n.x = tgb_sign(n.x); n.z = 0.0f; }
            else                         { n.z = tgb_sign(n.z); n.x = 0.0f; }
        }
        else
        {
            n.x = 0.0f;
            if (fabsf(n.y) > fabsf(n.z)) { n.y = tgb_sign(n.y); n.z = 0.0f; }
            else                         { n.z = tgb_sign(n.z); n.y = 0.0f; }
        }
        normal_ws = tgb_normalize(tgb_m4_transform(obj.rotation, n, 0.0f));
    }

    /* :231 -- un-normalised direction (Q3) */
    const v3 hit_position_ws = tgb_add(ray_origin_ws, tgb_scale(ray_direction_ws, depth_24b * a.cam.far_plane));

    /* :233-300 debug views */
    switch (a.debug_visualization)
    {
    case TG_DEBUG_SHOW_OBJECT_INDEX:    *p_color = tgb_hash_color(object_idx); return false;
    case TG_DEBUG_SHOW_DEPTH:           { const f32 g = tgb_min(1.0f, 8.0f * depth_24b); *p_color = make_float4(g, g, g, 1.0f); return false; }
    case TG_DEBUG_SHOW_CLUSTER_INDEX:
    case TG_DEBUG_SHOW_BLOCKS:          *p_color = tgb_hash_color(cluster_idx); return false;
    case TG_DEBUG_SHOW_VOXEL_INDEX:     *p_color = tgb_hash_color(voxel_idx_9b); return false;
    case TG_DEBUG_SHOW_COLOR_LUT_INDEX: *p_color = tgb_hash_color(color_lut_idx); return false;
    case TG_DEBUG_SHOW_COLOR:           *p_color = make_float4(color_r, color_g, color_b, 1.0f); return false;
    case TG_DEBUG_SHOW_NORMAL:          *p_color = make_float4(normal_ws.x * 0.5f + 0.5f, normal_ws.y * 0.5f + 0.5f, normal_ws.z * 0.5f + 0.5f, 1.0f); return false;
    default: break;
    }

    /* :285-316 */
    const f32 metallic = 0.1f;
    const v3 v = tgb_normalize(tgb_sub(ray_origin_ws, hit_position_ws));
    const v3 l = tgb_normalize(tgb_v3(0.0f, 0.8f, 0.3f));
    const v3 albedo = a.debug_visualization == TG_DEBUG_SHOW_SHADING ? tgb_v3(1.0f, 1.0f, 1.0f) : tgb_v3(color_r, color_g, color_b);
    const v3 specular_albedo = tgb_mix3(tgb_v3(0.04f, 0.04f, 0.04f), albedo, metallic);
    const f32 roughness = 0.8f;
    const v3 lo = tgb_shade_brdf(normal_ws, v, l, albedo, specular_albedo, metallic, roughness, tgb_v3(3.0f, 3.0f, 3.0f));
    const v3 ambient = tgb_scale(albedo, 0.1f);

    /* composed GI term (DESIGN.md "GI spec"; oracle/tgo_shade.c): one secondary ray towards the sky hemisphere */
    if (a.gi_enabled && a.debug_visualization == TG_DEBUG_SHOW_NONE && (normal_ws.x != 0.0f || normal_ws.y != 0.0f || normal_ws.z != 0.0f))
    {
        const u32 pixel_idx = a.w * py + px;
        u32 rng = tgb_hash_u32(pixel_idx ^ tgb_hash_u32(a.frame_seed)) | 1u;
        v3 dir = normal_ws;
        for (u32 attempt = 0; attempt < 32; attempt++)
        {
            v3 c;
            c.x = tgb_xorshift32_range(&rng, -1.0f, 1.0f);
            c.y = tgb_xorshift32_range(&rng, -1.0f, 1.0f);
            c.z = tgb_xorshift32_range(&rng, -1.0f, 1.0f);
            /* The shader normalises every candidate and tests dot(c, normal) > 0. The sign of that dot product is the sign of
             * dot(raw, normal) whenever the latter is clear of rounding: |raw| <= sqrt(3), |normal| = 1, so the un-normalised
             * sum carries an error below 5e-7 and, divided by the length, the normalised one is off by less than 4e-7 -- a
             * candidate with s < -1e-5 is rejected by the shader too, one with s > 1e-5 accepted, and only the sliver in
             * between (and NaN) needs the shader's own comparison. Half of the candidates skip the square root and divisions. */
            const f32 s = (c.x * normal_ws.x + c.y * normal_ws.y) + c.z * normal_ws.z;
            if (s < -1e-5f) continue;
            c = tgb_normalize(c);
            if (s > 1e-5f || tgb_dot(c, normal_ws) > 0.0f) { dir = c; break; }
        }
        const v3 origin = tgb_add(hit_position_ws, tgb_scale(dir, 1.73205080757f));
        /* svo_functions.inc:27-31: a ray that misses the root box is unoccluded; only the others are traced */
        const v3 extent = tgb_sub(a.svo.bmax, a.svo.bmin);
        const v3 center = tgb_add(tgb_scale(extent, 0.5f), a.svo.bmin);
        f32 e0, e1;
        if (tgb_ray_aabb(tgb_sub(origin, center), dir, a.svo.bmin, a.svo.bmax, &e0, &e1))
        {
            *p_color = make_float4(lo.x, lo.y, lo.z, 1.0f); /* ambient * 0 + lo, unless the ray escapes */
            *p_origin = origin; *p_dir = dir; *p_ambient = ambient; *p_root_enter = e0;
            return true;
        }
    }
    *p_color = make_float4(ambient.x + lo.x, ambient.y + lo.y, ambient.z + lo.z, 1.0f);
    return false;
}

/*
 * FAST (TGB_GI_KERNEL=4): two secondary rays of three are decided by the FIRST cell the certified walk enters (tgb_gi_fast.cuh: the
 * ray starts in the free space above the objects and the box of free cells around it reaches the root's border, or it starts next to a
 * solid voxel) -- those never see the queue: 48 bytes written and read again, a slot reserved, a set-up and a service per ray saved.
 * The pixel gets `ambient + lo` here exactly as the trace kernels would add it (one float addition per channel, commutative). Rays
 * the first steps leave undecided, or decide without certainty, are queued as before and start again from their record.
 */
template <bool RESOLVED, bool FAST, int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS) k_shade(const tgb_shade_args a)
{
    /* 8x4 pixel blocks per warp like K1: neighbouring pixels share clusters, objects and material bytes */
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u32 px = blockIdx.x * 16u + (warp & 1u) * 8u + (lane & 7u);
    const u32 vy = a.y0 + blockIdx.y * 16u + (warp >> 1) * 4u + (lane >> 3);  /* bands are 16 rows: a CTA stays inside one */
    const u32 py = tgb_row_to_physical(vy, a.n_ranks, a.tile_rows);
    const bool in_tile = px < a.w && vy < a.y1 && py < a.h;

    float4 color = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    v3 origin = tgb_v3(0.0f, 0.0f, 0.0f), dir = origin, ambient = origin;
    f32 root_enter = 0.0f; /* `enter` of the slab test against the SVO root (svo_functions.inc:27-31), queued for k_gi_trace_flat */
    bool trace = in_tile && tgb_shade_pixel<RESOLVED>(a, px, py, vy, &color, &origin, &dir, &ambient, &root_enter);
    u32 n_cells = 0, n_decided = 0;
    if (FAST)
    {
        bool decided = false;
        /* a tree k_svo_flatten could not tabulate (an uploaded tree with leaves off depth 5) has no tiling either: every ray is queued and
         * the stack machine k_gi_trace traces the queue */
        if (trace && a.fast_frame.p_grid[TGB_TOP_GRID_CELLS] != 0u)
        {
            tgb_fast_ray r;
            u32 kind = tgb_fast_start(&a.fast_frame, origin, dir, root_enter, a.fast_delta, &r, true);
            if (kind == TGB_FAST_WALK) kind = tgb_fast_walk_tiled<false>(&a.fast_frame, &a.fast_tiling, &r, a.fast_steps, (u32*)0, (u32*)0);
            n_cells = r.n_steps;
            if (kind == TGB_FAST_OCCLUDED) decided = true;
            else if (kind == TGB_FAST_UNOCCLUDED && !(r.flags & TGB_FAST_UNCERTAIN))
            {
                decided = true;
                color.x = ambient.x + color.x; color.y = ambient.y + color.y; color.z = ambient.z + color.z;
            }
            trace = !decided;
        }
        n_decided = (u32)__popc(__ballot_sync(0xFFFFFFFFu, decided));
        n_cells = __reduce_add_sync(0xFFFFFFFFu, decided ? n_cells : 0u);
    }
    if (in_tile) a.p_out[(u64)vy * a.w + px] = color;

    /* CTA-aggregated append to the ray queue: one atomic per 16x16 tile. (One per warp was 259 k atomics per 4K frame on a single address, a
     * fifth of this kernel's stall samples once the counters of the FAST path were added to the same line: profiles/r04b_k_shade.) */
    __shared__ u32 s_count[8], s_cells[8], s_decided[8], s_base;
    const u32 m = __ballot_sync(0xFFFFFFFFu, trace);
    if (lane == 0) { s_count[warp] = (u32)__popc(m); if (FAST) { s_cells[warp] = n_cells; s_decided[warp] = n_decided; } }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        u32 total = 0, cells = 0, dec = 0;
        for (u32 i = 0; i < 8u; i++) { total += s_count[i]; if (FAST) { cells += s_cells[i]; dec += s_decided[i]; } }
        s_base = total ? atomicAdd(&a.p_q_count[0], total) : 0u;
        if (FAST && dec)
        {
            /* the frame's counters: [10] rays of the frame (the trace kernel adds the queued ones), u64 [1] cells entered */
            atomicAdd(&a.p_q_count[10], dec);
            atomicAdd(reinterpret_cast<unsigned long long*>(a.p_q_count) + 1, (unsigned long long)cells);
        }
    }
    __syncthreads();
    if (trace)
    {
        u32 slot = s_base + (u32)__popc(m & ((1u << lane) - 1u));
        for (u32 i = 0; i < warp; i++) slot += s_count[i];
        a.p_q0[slot] = make_float4(origin.x, origin.y, origin.z, __uint_as_float(a.w * vy + px)); /* where the ambient term goes */
        a.p_q1[slot] = make_float4(dir.x, dir.y, dir.z, root_enter);
        a.p_q2[slot] = make_float4(ambient.x, ambient.y, ambient.z, 0.0f);
    }
}

/* ---- K3b: the queued secondary rays through the SVO ------------------------------------------------------ */
/*
 * tg_svo_traverse (svo_functions.inc:1-329) as a per-lane state machine inside a persistent kernel.
 *
 * Secondary rays are incoherent (random hemisphere directions): inside one warp some lanes descend the tree, some
 * walk a 32^3 leaf block voxel by voxel, some advance to the next node, and trip counts differ by two orders of
 * magnitude. Run as one loop per thread, a warp executes the union of all those paths with a handful of lanes active
 * (measured: 4.3 of 32, profiles/r01c). Here every ray is in one of three states and each warp iteration executes the
 * ONE phase most of its lanes are waiting for, so the lanes that execute it do so together:
 *   NODE  one visit of the shader's while loop up to the decision (:44-110, 262-270): read the node on top of the
 *         stack, pick the octant; inner child -> push; leaf with data -> set up the DDA; otherwise -> ADV
 *   ADV   advance to the far border of the child and pop every stacked node the ray has left (:279-324), then visit
 *   DDA   up to TGB_GI_DDA_STEPS steps of the leaf DDA (:111-257); solid voxel before the far plane -> hit; leaving
 *         the block -> ADV
 * One REDUX per iteration counts the lanes of each kind. A lane whose ray finished fetches the next queued ray. The (node, min, max) stack of svo.inc:4 lives in shared
 * memory ([level][component][thread], conflict-free), the top entry is cached in registers. The arithmetic that
 * moves `position` or decides a comparison is the shader's, operation for operation.
 */
#define TGB_GI_THREADS   128
#define TGB_GI_DDA_STEPS 8
#define TGB_GI_FLAT_CTAS_PER_SM 8
enum { TGB_ST_IDLE = 0, TGB_ST_NODE = 1, TGB_ST_DDA = 2, TGB_ST_ADV = 3 };

__global__ void __launch_bounds__(TGB_GI_THREADS) k_gi_trace(const tgb_svo_view svo, const u32* __restrict__ p_flat_ok, f32 far_plane, const float4* __restrict__ p_q0, const float4* __restrict__ p_q1,
                                                             const float4* __restrict__ p_q2, u32* __restrict__ p_q_count, float4* __restrict__ p_out)
{
    if (p_flat_ok != NULL && p_flat_ok[0] != 0) return; /* the stackless kernel k_gi_trace_flat traces this frame */

    /* stack entries 1..4 (entry 0 is the root: node 0, the SVO box) */
    __shared__ f32 s_box[4][6][TGB_GI_THREADS];
    __shared__ u32 s_idx[4][TGB_GI_THREADS];

    const u32 tid = threadIdx.x, lane = tid & 31u;
    const u32 n_rays = p_q_count[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(&p_q_count[10], n_rays); atomicAdd(&p_q_count[14], n_rays); } /* rays of the frame, summed over its bands; all of them traced exactly */
    const v3 extent = tgb_sub(svo.bmax, svo.bmin);
    const v3 center = tgb_add(tgb_scale(extent, 0.5f), svo.bmin); /* svo_functions.inc:3-8 */

    u32 state = TGB_ST_IDLE, slot = 0, iterations = 0, stack_size = 0, top_idx = 0;
    v3 o = tgb_v3(0.0f, 0.0f, 0.0f), d = o, position = o, top_min = o, top_max = o, child_min = o, child_max = o;
    f32 t_max_x = 0.0f, t_max_y = 0.0f, t_max_z = 0.0f, t_delta_x = 0.0f, t_delta_y = 0.0f, t_delta_z = 0.0f;
    i32 x = 0, y = 0, z = 0;
    const u32* __restrict__ p_block = svo.p_voxels;
    bool exhausted = false;
    u32 n_visits = 0, n_steps = 0, n_advances = 0; /* work counters (reported through tgb200_timings) */

    for (;;)
    {
        /* one REDUX counts the lanes per kind: idle | in a leaf DDA | in the tree (visit or advance pending) */
        const u32 counts = __reduce_add_sync(0xFFFFFFFFu, 1u << (8u * state));
        const u32 n_idle = counts & 0xFFu, n_node = (counts >> 8) & 0xFFu, n_dda = (counts >> 16) & 0xFFu, n_adv = counts >> 24;

        /* ---- refill ---- */
        if (!exhausted && n_idle >= 8u)
        {
            const u32 idle = __ballot_sync(0xFFFFFFFFu, state == TGB_ST_IDLE);
            u32 base = 0;
            const u32 leader = (u32)(__ffs(idle) - 1);
            if (lane == leader) base = atomicAdd(&p_q_count[1], n_idle);
            base = __shfl_sync(0xFFFFFFFFu, base, (int)leader);
            if (state == TGB_ST_IDLE)
            {
                const u32 mine = base + (u32)__popc(idle & ((1u << lane) - 1u));
                if (mine < n_rays)
                {
                    slot = mine;
                    const float4 q0 = p_q0[mine], q1 = p_q1[mine];
                    o = tgb_sub(tgb_v3(q0.x, q0.y, q0.z), center);
                    d = tgb_v3(q1.x, q1.y, q1.z);
                    f32 enter, exit;
                    if (tgb_ray_aabb(o, d, svo.bmin, svo.bmax, &enter, &exit)) /* :27-31; true: tested before queueing */
                    {
                        position = o;
                        if (enter > 0.0f) position = tgb_add(position, tgb_scale(d, enter));
                        top_min = svo.bmin; top_max = svo.bmax; top_idx = 0;
                        stack_size = 1; iterations = 0;
                        /* :139-176: the DDA increments depend on the ray only */
                        t_delta_x = d.x > 0.0f ? 1.0f / d.x : (d.x < 0.0f ? 1.0f / -d.x : TG_F32_MAX);
                        t_delta_y = d.y > 0.0f ? 1.0f / d.y : (d.y < 0.0f ? 1.0f / -d.y : TG_F32_MAX);
                        t_delta_z = d.z > 0.0f ? 1.0f / d.z : (d.z < 0.0f ? 1.0f / -d.z : TG_F32_MAX);
                        state = TGB_ST_NODE;
                    }
                }
            }
            exhausted = base + n_idle >= n_rays;
            continue;
        }
        if (n_idle == 32u) break; /* queue drained and every ray finished */

        u32 finished = 0; /* 1 = occluded, 2 = unoccluded */
        if (n_dda >= n_node && n_dda >= n_adv)
        {
            if (state == TGB_ST_DDA)
            {
                /* :178-257; the step is written with selects (no branch on which axis advances): adding +0 to the
                 * other two t_max leaves them bit-identical */
                const i32 step_x = d.x > 0.0f ? 1 : (d.x < 0.0f ? -1 : 0);
                const i32 step_y = d.y > 0.0f ? 1 : (d.y < 0.0f ? -1 : 0);
                const i32 step_z = d.z > 0.0f ? 1 : (d.z < 0.0f ? -1 : 0);
#pragma unroll 1
                for (u32 k = 0; k < TGB_GI_DDA_STEPS; k++)
                {
                    /* a block row is one word: bit 1024 z + 32 y + x (the builder only makes 32^3 blocks) */
                    const u32 bits = __ldg(&p_block[32 * z + y]);
                    n_steps++;
                    if ((bits >> x) & 1u)
                    {
                        const v3 voxel_min = tgb_add(child_min, tgb_v3((f32)x, (f32)y, (f32)z));
                        const v3 voxel_max = tgb_add(child_min, tgb_v3((f32)(x + 1), (f32)(y + 1), (f32)(z + 1)));
                        f32 enter, exit;
                        tgb_ray_aabb(o, d, voxel_min, voxel_max, &enter, &exit);
                        /* :219-256 result = enter / far; only result < 1 ends the shader's loop, otherwise it advances */
                        if (enter / far_plane < 1.0f) finished = 1u; else state = TGB_ST_ADV;
                        break;
                    }
                    const bool xy = t_max_x < t_max_y;
                    const bool go_x = xy & (t_max_x < t_max_z);
                    const bool go_y = !xy & (t_max_y < t_max_z);
                    const bool go_z = !(go_x | go_y);
                    t_max_x = go_x ? t_max_x + t_delta_x : t_max_x;
                    t_max_y = go_y ? t_max_y + t_delta_y : t_max_y;
                    t_max_z = go_z ? t_max_z + t_delta_z : t_max_z;
                    x += go_x ? step_x : 0;
                    y += go_y ? step_y : 0;
                    z += go_z ? step_z : 0;
                    if ((u32)(x | y | z) > 31u) { state = TGB_ST_ADV; break; } /* left the block: a coordinate is -1 or 32 */
                }
            }
        }
        else if (n_adv > n_node)
        {
            bool popping = false;
            if (state == TGB_ST_ADV)
            {
                /* :279-294 advance to the far border of the child */
                n_advances++;
                const f32 exit = tgb_exit_distance(child_min, child_max, position, d);
                position = tgb_add(position, tgb_scale(d, exit + TG_F32_EPSILON));
                state = TGB_ST_NODE;
                popping = !tgb_still_inside(top_min, top_max, position, d);
                stack_size -= popping ? 1u : 0u;
            }
            /* :296-324: pop while the ray has left the stacked node; the lanes of this phase pop in lock-step */
            while (__any_sync(0xFFFFFFFFu, popping))
            {
                if (popping)
                {
                    if (stack_size == 0) { finished = 2u; state = TGB_ST_IDLE; popping = false; }
                    else
                    {
                        const u32 e = stack_size >= 2u ? stack_size - 2u : 0u;
                        const bool root = stack_size == 1u;
                        top_min = root ? svo.bmin : tgb_v3(s_box[e][0][tid], s_box[e][1][tid], s_box[e][2][tid]);
                        top_max = root ? svo.bmax : tgb_v3(s_box[e][3][tid], s_box[e][4][tid], s_box[e][5][tid]);
                        top_idx = root ? 0u : s_idx[e][tid];
                        popping = !tgb_still_inside(top_min, top_max, position, d);
                        stack_size -= popping ? 1u : 0u;
                    }
                }
            }
        }
        else
        {
            if (state == TGB_ST_NODE)
            {
                /* one visit of the shader's while loop: :44-110, 262-270 */
                n_visits++;
                if (++iterations > TGB_TRAVERSE_MAX_ITERS) finished = 2u;
                else
                {
                    const u32 node_data = __ldg(&svo.p_nodes[top_idx]);
                    const u32 child_pointer =  node_data        & 0xFFFFu;
                    const u32 valid_mask    = (node_data >> 16) & 0xFFu;
                    const u32 leaf_mask     = (node_data >> 24) & 0xFFu;
                    /* :57-80 */
                    /* with selects: child_min = parent_min [+ extent], child_max = (parent_min + extent) [+ extent], the shader's sums */
                    const v3 child_extent = tgb_scale(tgb_sub(top_max, top_min), 0.5f);
                    const v3 mid = tgb_add(top_min, child_extent);
                    const bool ux = (mid.x < position.x) | ((position.x == mid.x) & (d.x > 0.0f));
                    const bool uy = (mid.y < position.y) | ((position.y == mid.y) & (d.y > 0.0f));
                    const bool uz = (mid.z < position.z) | ((position.z == mid.z) & (d.z > 0.0f));
                    const u32 oct = (ux ? 1u : 0u) | (uy ? 2u : 0u) | (uz ? 4u : 0u);
                    child_min = tgb_v3(ux ? mid.x : top_min.x, uy ? mid.y : top_min.y, uz ? mid.z : top_min.z);
                    child_max = tgb_v3(ux ? mid.x + child_extent.x : mid.x, uy ? mid.y + child_extent.y : mid.y, uz ? mid.z + child_extent.z : mid.z);
                    state = TGB_ST_ADV;
                    if ((valid_mask & (1u << oct)) != 0)
                    {
                        /* :86-91 */
                        const u32 child_idx = top_idx + child_pointer + (u32)__popc(valid_mask & ((1u << oct) - 1u));
                        if ((leaf_mask & (1u << oct)) == 0)
                        {
                            /* :262-270 push */
                            if (stack_size >= TG_SVO_TRAVERSE_STACK_CAPACITY) finished = 2u; /* malformed tree */
                            else
                            {
                                const u32 e = stack_size - 1u; /* new entry `stack_size` is stored in slot stack_size - 1 */
                                s_box[e][0][tid] = child_min.x; s_box[e][1][tid] = child_min.y; s_box[e][2][tid] = child_min.z;
                                s_box[e][3][tid] = child_max.x; s_box[e][4][tid] = child_max.y; s_box[e][5][tid] = child_max.z;
                                s_idx[e][tid] = child_idx;
                                stack_size++;
                                top_min = child_min; top_max = child_max; top_idx = child_idx;
                                state = TGB_ST_NODE;
                            }
                        }
                        else
                        {
                            const u32 data_pointer = __ldg(&svo.p_nodes[child_idx]);
                            if (__ldg(&svo.p_leaf_data[(u64)data_pointer * 65u]) != 0)
                            {
                                /* :111-176 (blocks are 32^3: tgbd_svo_build / tgbd_svo_set only accept a 1024^3 box) */
                                p_block = svo.p_voxels + (u64)data_pointer * TG_SVO_BLOCK_WORDS;
                                v3 hit = position;
                                v3 xyz = tgb_v3(tgb_clamp(floorf(hit.x), child_min.x, child_max.x - 1.0f),
                                                tgb_clamp(floorf(hit.y), child_min.y, child_max.y - 1.0f),
                                                tgb_clamp(floorf(hit.z), child_min.z, child_max.z - 1.0f));
                                hit = tgb_sub(hit, child_min);
                                xyz = tgb_sub(xyz, child_min);
                                x = (i32)xyz.x; y = (i32)xyz.y; z = (i32)xyz.z;
                                t_max_x = TG_F32_MAX; t_max_y = TG_F32_MAX; t_max_z = TG_F32_MAX;
                                if (d.x > 0.0f)      t_max_x = ((f32)(x + 1) - hit.x) / d.x;
                                else if (d.x < 0.0f) t_max_x = (hit.x - (f32)x) / -d.x;
                                if (d.y > 0.0f)      t_max_y = ((f32)(y + 1) - hit.y) / d.y;
                                else if (d.y < 0.0f) t_max_y = (hit.y - (f32)y) / -d.y;
                                if (d.z > 0.0f)      t_max_z = ((f32)(z + 1) - hit.z) / d.z;
                                else if (d.z < 0.0f) t_max_z = (hit.z - (f32)z) / -d.z;
                                state = TGB_ST_DDA;
                            }
                        }
                    }
                }
            }
        }

        if (finished)
        {
#ifdef TGB_GI_HISTOGRAM
            atomicMax(&p_q_count[8], iterations);
            atomicAdd(&p_q_count[16 + (31 - __clz(iterations | 1u))], 1u);
#endif
            if (finished == 2u)
            {
                /* unoccluded: the ambient term comes back (ambient * 1 + lo) */
                const float4 q0 = p_q0[slot], q2 = p_q2[slot];
                const u32 pixel = __float_as_uint(q0.w);
                float4 c = p_out[pixel];
                c.x = q2.x + c.x; c.y = q2.y + c.y; c.z = q2.z + c.z;
                p_out[pixel] = c;
            }
            state = TGB_ST_IDLE;
        }
    }
    /* [2] node visits, [3] DDA steps, [4] advances of this frame */
    n_visits = __reduce_add_sync(0xFFFFFFFFu, n_visits);
    n_steps = __reduce_add_sync(0xFFFFFFFFu, n_steps);
    n_advances = __reduce_add_sync(0xFFFFFFFFu, n_advances);
    if (lane == 0)
    {
        atomicAdd(reinterpret_cast<unsigned long long*>(p_q_count) + 1, (unsigned long long)n_visits);
        atomicAdd(reinterpret_cast<unsigned long long*>(p_q_count) + 2, (unsigned long long)n_steps);
        atomicAdd(reinterpret_cast<unsigned long long*>(p_q_count) + 3, (unsigned long long)n_advances);
    }
}

/* ---- K3b, stackless: the same traversal over the flattened tree ------------------------------------------------ */
/*
 * tg_svo_traverse keeps a stack only to find, after every advance, the deepest stacked node that still contains
 * `position`, and then descends again to the terminal node (an invalid octant or a leaf) around it. Both steps are a
 * function of `position` and the shader's comparison rules alone:
 *   - the pop loop (:296-324) ends the ray when the ROOT fails `exit > epsilon`: every stacked box lies inside the root
 *     and IEEE subtraction / division are monotone, so a nested box that passes the test implies the root passes it;
 *   - the descent (:57-91) picks per axis the upper half iff mid < p || (p == mid && d > 0); five levels of that rule
 *     select one 32^3 cell of the box, and the terminal node around it is tabulated per cell by k_svo_flatten
 *     (tgb_svo.cu). Whether the descent starts at the root or at a stacked ancestor is immaterial: the ancestor
 *     contains `position` under the same rule (it passed the pop test on every axis the ray moves along).
 * What moves `position` -- the far-border distance of the terminal box and `position += (exit + epsilon) * d`
 * (:279-294) -- and the leaf DDA (:111-257) are the shader's operations, so hit / miss decisions are the oracle's.
 * The box corners must be multiples of 32 (the chain mid = min + extent / 2 is then exact and equals min + 32 * cell);
 * other boxes, and trees k_svo_flatten could not tabulate, take the stack kernel above.
 * Two kinds of lanes remain: TREE (advance + look-up) and DDA (up to TGB_FL_DDA_STEPS voxel steps); each warp iteration
 * runs the phase the majority waits for.
 */
/*
 * Lane kinds. TREE and DDA are the two working phases; HIT (a solid voxel was found: the shader's slab test against it
 * decides, :219-256), MISS (the ray left the root: the ambient term comes back) and IDLE (needs a ray) are rare events
 * per iteration, so they wait until a quarter of the warp needs service and are then handled together.
 */
enum { TGB_FL_IDLE = 0, TGB_FL_TREE = 1, TGB_FL_DDA = 2, TGB_FL_HIT = 3, TGB_FL_MISS = 4 };
#define TGB_FL_SERVICE_LANES 16u /* measured best with 16 DDA steps and 4 tree cells per phase */
#define TGB_FL_DDA_STEPS 16
#define TGB_FL_TREE_REPS 4 /* cells a ray may cross per tree phase: 1.548 -> 1.506 ms for the stage with 16 service lanes (sweeps of this round) */

template <int DDA_STEPS>
__global__ void __launch_bounds__(TGB_GI_THREADS) k_gi_trace_flat(const tgb_svo_view svo, const u32* __restrict__ p_grid, f32 far_plane,
                                                                  const float4* __restrict__ p_q0, const float4* __restrict__ p_q1, const float4* __restrict__ p_q2,
                                                                  u32* __restrict__ p_q_count, float4* __restrict__ p_out, u32 service_lanes, u32 dda_bias, u32 tree_reps)
{
    if (p_grid[TGB_TOP_GRID_CELLS] == 0) return; /* not tabulated: k_gi_trace runs */

    const u32 lane = threadIdx.x & 31u;
    const u32 n_rays = p_q_count[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(&p_q_count[10], n_rays); atomicAdd(&p_q_count[14], n_rays); } /* rays of the frame, summed over its bands; all of them traced exactly */
    const v3 extent = tgb_sub(svo.bmax, svo.bmin);
    const v3 center = tgb_add(tgb_scale(extent, 0.5f), svo.bmin); /* svo_functions.inc:3-8 */
    const v3 box_mid = tgb_scale(tgb_add(svo.bmin, svo.bmax), 0.5f);
    const i32 min_cell_x = (i32)(svo.bmin.x * 0.03125f), min_cell_y = (i32)(svo.bmin.y * 0.03125f), min_cell_z = (i32)(svo.bmin.z * 0.03125f);

    u32 kind = TGB_FL_IDLE, slot = 0, iterations = 0;
    bool advance_pending = false, setup_pending = false, border_pending = false, exotic = false;
    v3 d = tgb_v3(0.0f, 0.0f, 0.0f), position = d, child_min = d;
    f32 child_size = 0.0f;
    f32 t_max_x = 0.0f, t_max_y = 0.0f, t_max_z = 0.0f, t_delta_x = 0.0f, t_delta_y = 0.0f, t_delta_z = 0.0f;
    i32 x = 0, y = 0, z = 0;
    const u32* __restrict__ p_block = svo.p_voxels;
    bool exhausted = false;
    u32 n_visits = 0, n_steps = 0, n_advances = 0;

    for (;;)
    {
        const u32 counts = __reduce_add_sync(0xFFFFFFFFu, 1u << (6u * kind));
        const u32 n_idle = counts & 63u, n_tree = (counts >> 6) & 63u, n_dda = (counts >> 12) & 63u, n_hit = (counts >> 18) & 63u, n_miss = (counts >> 24) & 63u;
        const u32 n_service = n_hit + n_miss + (exhausted ? 0u : n_idle);
        const u32 n_working = n_tree + n_dda;
        if (n_working == 0 && n_service == 0) break; /* queue drained and every ray finished */

        if (n_service >= service_lanes || n_working == 0)
        {
            /* ---- service: unoccluded rays return their ambient term, voxel hits are decided, idle lanes fetch rays ---- */
            if (kind == TGB_FL_MISS && border_pending)
            {
                /* :296-324 for a ray within one unit of a root face: still inside -> back to the tree, position already advanced */
                border_pending = false;
                if (tgb_still_inside(svo.bmin, svo.bmax, position, d)) { kind = TGB_FL_TREE; advance_pending = false; }
            }
            if (kind == TGB_FL_MISS)
            {
                /* ambient * 1 + lo; float addition commutes and the reductions do not stall the lane */
                const float4 q0 = p_q0[slot], q2 = p_q2[slot];
                f32* p_pixel = reinterpret_cast<f32*>(&p_out[__float_as_uint(q0.w)]);
                atomicAdd(p_pixel + 0, q2.x);
                atomicAdd(p_pixel + 1, q2.y);
                atomicAdd(p_pixel + 2, q2.z);
                kind = TGB_FL_IDLE;
            }
            else if (kind == TGB_FL_HIT)
            {
                /* :219-256: result = enter / far of the slab test against the voxel. Only `enter` matters: the largest of the
                 * three near-plane quotients, and min((lo - o) / d, (hi - o) / d) is the quotient of the plane the ray meets
                 * first (division by d is monotone), so three divisions give the shader's value. */
                const float4 q0 = p_q0[slot];
                const v3 o = tgb_sub(tgb_v3(q0.x, q0.y, q0.z), center);
                const v3 lo = tgb_add(child_min, tgb_v3((f32)x, (f32)y, (f32)z));
                const v3 hi = tgb_add(child_min, tgb_v3((f32)(x + 1), (f32)(y + 1), (f32)(z + 1)));
                const f32 ex = d.x == 0.0f ? TG_F32_MIN : ((d.x > 0.0f ? lo.x : hi.x) - o.x) / d.x;
                const f32 ey = d.y == 0.0f ? TG_F32_MIN : ((d.y > 0.0f ? lo.y : hi.y) - o.y) / d.y;
                const f32 ez = d.z == 0.0f ? TG_F32_MIN : ((d.z > 0.0f ? lo.z : hi.z) - o.z) / d.z;
                const f32 enter = tgb_max(tgb_max(ex, ey), ez);
                /* only result < 1 ends the shader's loop (occluded), otherwise it advances past the leaf */
                kind = (enter / far_plane < 1.0f) ? TGB_FL_IDLE : TGB_FL_TREE;
            }
            if (!exhausted)
            {
                const u32 idle = __ballot_sync(0xFFFFFFFFu, kind == TGB_FL_IDLE);
                if (idle)
                {
                    const u32 n = (u32)__popc(idle);
                    u32 base = 0;
                    const u32 leader = (u32)(__ffs(idle) - 1);
                    if (lane == leader) base = atomicAdd(&p_q_count[1], n);
                    base = __shfl_sync(0xFFFFFFFFu, base, (int)leader);
                    const u32 mine = base + (u32)__popc(idle & ((1u << lane) - 1u));
                    if (kind == TGB_FL_IDLE && mine < n_rays)
                    {
                        slot = mine;
                        const float4 q0 = p_q0[mine], q1 = p_q1[mine];
                        d = tgb_v3(q1.x, q1.y, q1.z);
                        /* :27-31: k_shade made the slab test against the root and queued its `enter` */
                        position = tgb_sub(tgb_v3(q0.x, q0.y, q0.z), center);
                        if (q1.w > 0.0f) position = tgb_add(position, tgb_scale(d, q1.w));
                        iterations = 0;
                        advance_pending = false;
                        /* :139-176: the DDA increments 1 / |d| depend on the ray only (rcp.rn == IEEE 1 / x) */
                        const f32 ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
                        t_delta_x = ax != 0.0f ? __frcp_rn(ax) : TG_F32_MAX;
                        t_delta_y = ay != 0.0f ? __frcp_rn(ay) : TG_F32_MAX;
                        t_delta_z = az != 0.0f ? __frcp_rn(az) : TG_F32_MAX;
                        exotic = (ax != 0.0f && ax < 1e-30f) || (ay != 0.0f && ay < 1e-30f) || (az != 0.0f && az < 1e-30f);
                        kind = TGB_FL_TREE;
                    }
                    exhausted = base + n >= n_rays;
                }
            }
            continue;
        }

        if (n_dda + dda_bias > n_tree && n_dda > 0)
        {
            if (kind == TGB_FL_DDA)
            {
                if (setup_pending)
                {
                    /* :111-176: the lanes that entered a leaf since the last DDA phase set up together */
                    setup_pending = false;
                    v3 hit = position;
                    v3 xyz = tgb_v3(tgb_clamp(floorf(hit.x), child_min.x, (child_min.x + child_size) - 1.0f),
                                    tgb_clamp(floorf(hit.y), child_min.y, (child_min.y + child_size) - 1.0f),
                                    tgb_clamp(floorf(hit.z), child_min.z, (child_min.z + child_size) - 1.0f));
                    hit = tgb_sub(hit, child_min);
                    xyz = tgb_sub(xyz, child_min);
                    x = (i32)xyz.x; y = (i32)xyz.y; z = (i32)xyz.z;
                    t_max_x = TG_F32_MAX; t_max_y = TG_F32_MAX; t_max_z = TG_F32_MAX;
                    if (d.x > 0.0f)      t_max_x = ((f32)(x + 1) - hit.x) / d.x;
                    else if (d.x < 0.0f) t_max_x = (hit.x - (f32)x) / -d.x;
                    if (d.y > 0.0f)      t_max_y = ((f32)(y + 1) - hit.y) / d.y;
                    else if (d.y < 0.0f) t_max_y = (hit.y - (f32)y) / -d.y;
                    if (d.z > 0.0f)      t_max_z = ((f32)(z + 1) - hit.z) / d.z;
                    else if (d.z < 0.0f) t_max_z = (hit.z - (f32)z) / -d.z;
                }
                /* :178-257, steps written with selects (adding +0 to the other two t_max leaves them bit-identical) */
                const i32 step_x = d.x > 0.0f ? 1 : (d.x < 0.0f ? -1 : 0);
                const i32 step_y = d.y > 0.0f ? 1 : (d.y < 0.0f ? -1 : 0);
                const i32 step_z = d.z > 0.0f ? 1 : (d.z < 0.0f ? -1 : 0);
                u32 bits = __ldg(&p_block[32 * z + y]); /* a block row is one word: bit 1024 z + 32 y + x; re-read only when the row changes */
#pragma unroll 1
                for (u32 k = 0; k < (u32)DDA_STEPS; k++)
                {
                    n_steps++;
                    if ((bits >> x) & 1u) { kind = TGB_FL_HIT; break; }
                    const bool xy = t_max_x < t_max_y;
                    const bool go_x = xy & (t_max_x < t_max_z);
                    const bool go_y = !xy & (t_max_y < t_max_z);
                    const bool go_z = !(go_x | go_y);
                    t_max_x = go_x ? t_max_x + t_delta_x : t_max_x;
                    t_max_y = go_y ? t_max_y + t_delta_y : t_max_y;
                    t_max_z = go_z ? t_max_z + t_delta_z : t_max_z;
                    x += go_x ? step_x : 0;
                    y += go_y ? step_y : 0;
                    z += go_z ? step_z : 0;
                    if ((u32)(x | y | z) > 31u) { kind = TGB_FL_TREE; break; } /* left the block: a coordinate is -1 or 32 */
                    if (!go_x) bits = __ldg(&p_block[32 * z + y]);
                }
            }
        }
        else
#pragma unroll 1
        for (u32 rep = 0; rep < tree_reps && kind == TGB_FL_TREE; rep++) /* a ray crossing empty cells stays in the tree phase: up to tree_reps cells per phase */
        {
            if (advance_pending)
            {
                /* :279-294 advance to the far border of the terminal box, :296-324 the ray ends when it has left the root */
                n_advances++;
                const f32 exit = tgb_exit_distance_rcp(child_min, child_size, position, d, t_delta_x, t_delta_y, t_delta_z, exotic);
                position = tgb_add(position, tgb_scale(d, exit + TG_F32_EPSILON));
                /* a position at least one unit inside every face passes the pop test (exit >= 1 / |d| >= ~1 > epsilon) without evaluating it */
                const f32 off = fmaxf(fmaxf(fabsf(position.x - box_mid.x), fabsf(position.y - box_mid.y)), fabsf(position.z - box_mid.z));
                if (!(off < 0.5f * (f32)TG_SVO_SIDE_LENGTH - 1.0f)) { kind = TGB_FL_MISS; border_pending = true; } /* the few rays near a face: the test itself runs in the service phase */
            }
            advance_pending = true;
            if (kind == TGB_FL_TREE)
            {
                if (++iterations > TGB_TRAVERSE_MAX_ITERS) kind = TGB_FL_MISS;
                else
                {
                    /* :44-110: the terminal node around `position` */
                    n_visits++;
                    const u32 cx = tgb_cell_axis(position.x, d.x, min_cell_x);
                    const u32 cy = tgb_cell_axis(position.y, d.y, min_cell_y);
                    const u32 cz = tgb_cell_axis(position.z, d.z, min_cell_z);
                    const u32 entry = __ldg(&p_grid[(cz << 10) | (cy << 5) | cx]);
                    const u32 level = (entry >> TGB_TOP_LEVEL_SHIFT) & 7u;
                    const u32 cells = 16u >> level;              /* side of the terminal box in cells */
                    const u32 keep = ~(cells - 1u);
                    child_size = (f32)(cells << 5);
                    child_min = tgb_v3(svo.bmin.x + (f32)((cx & keep) << 5), svo.bmin.y + (f32)((cy & keep) << 5), svo.bmin.z + (f32)((cz & keep) << 5));
                    if (entry & TGB_TOP_HAS_DATA)
                    {
                        p_block = svo.p_voxels + (u64)(entry & TGB_TOP_POINTER_MASK) * TG_SVO_BLOCK_WORDS;
                        setup_pending = true;
                        kind = TGB_FL_DDA;
                    }
                }
            }
        }
    }
    /* [2] look-ups, [3] DDA steps, [4] advances of this frame */
    n_visits = __reduce_add_sync(0xFFFFFFFFu, n_visits);
    n_steps = __reduce_add_sync(0xFFFFFFFFu, n_steps);
    n_advances = __reduce_add_sync(0xFFFFFFFFu, n_advances);
    if (lane == 0)
    {
        atomicAdd(reinterpret_cast<unsigned long long*>(p_q_count) + 1, (unsigned long long)n_visits);
        atomicAdd(reinterpret_cast<unsigned long long*>(p_q_count) + 2, (unsigned long long)n_steps);
        atomicAdd(reinterpret_cast<unsigned long long*>(p_q_count) + 3, (unsigned long long)n_advances);
    }
}

/* ---- owner-resolved materials (multi-GPU) ------------------------------------------------------------------ */
/*
 * SURVEY.md section 8e "second exchange": the LUT-index bytes (512 B per cluster) exist only on the GPU that owns the
 * cluster. After the visibility merge every rank looks at every pixel; where the winning pointer is its own it writes
 * ((global object idx + 1) << 32 | packed colour), elsewhere 0. A max-reduce-scatter by screen tile then hands each rank the
 * resolved words of exactly the rows it shades.
 */
__global__ void __launch_bounds__(256) k_resolve_material(const u64* __restrict__ p_vis, u64 n_pixels, u64 n_padded, const u32* __restrict__ p_cluster_pointers, const u32* __restrict__ p_c2o,
                                                          const tg_object_data* __restrict__ p_objects, const u8* __restrict__ p_lut_idx, const u32* __restrict__ p_color_lut,
                                                          u32 global_pointer_base, u32 n_local_pointers, u32 global_object_base, u64* __restrict__ p_mat)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_padded) return;
    u64 word = 0;
    if (i < n_pixels)
    {
        const u64 packed_data = p_vis[i];
        const u32 depth24 = (u32)(packed_data >> TG_VIS_DEPTH_SHIFT);
        const u32 local_pointer = ((u32)(packed_data >> TG_VIS_POINTER_SHIFT) & 2147483647u) - global_pointer_base;
        if ((f32)depth24 / TG_VIS_DEPTH_SCALE < 1.0f && local_pointer < n_local_pointers)
        {
            const u32 cluster_idx = __ldg(&p_cluster_pointers[local_pointer]);
            const u32 object_idx = __ldg(&p_c2o[cluster_idx]);
            const u32 color_lut_idx = __ldg(&p_lut_idx[(u64)cluster_idx * 512u + ((u32)packed_data & 511u)]);
            const u32 packed_color = __ldg(&p_color_lut[p_objects[object_idx].lut_idx * 256u + color_lut_idx]);
            word = ((u64)(global_object_base + object_idx + 1u) << 32) | (u64)packed_color; /* + 1: a resolved word is never 0 */
        }
    }
    p_mat[i] = word;
}

/* the same for the flagged 16x16 tiles only (sharded frame on the peer-memory path) */
__global__ void __launch_bounds__(256) k_resolve_material_tiles(const u64* __restrict__ p_vis, u32 w, u32 tiles_x, const u32* __restrict__ p_tile_flags,
                                                                const u32* __restrict__ p_cluster_pointers, const u32* __restrict__ p_c2o,
                                                                const tg_object_data* __restrict__ p_objects, const u8* __restrict__ p_lut_idx, const u32* __restrict__ p_color_lut,
                                                                u32 global_pointer_base, u32 n_local_pointers, u32 global_object_base, u64* __restrict__ p_mat)
{
    /* a CTA walks 8 neighbouring tiles of its band: most of a rank's tiles have no hit, and a launch of one CTA per tile costs more than the flagged tiles' work */
    for (u32 t = 0; t < 8u; t++)
    {
        const u32 tile_x = blockIdx.x * 8u + t;
        if (tile_x >= tiles_x) break;
        if (p_tile_flags[blockIdx.y * tiles_x + tile_x] == 0u) continue; /* no hit: nobody reads this tile's material words */
        const u32 x = tile_x * 16u + (threadIdx.x & 15u), vy = blockIdx.y * 16u + (threadIdx.x >> 4);
        if (x >= w) continue;
        const u64 i = (u64)vy * w + x;
        const u64 packed_data = p_vis[i];
        const u32 depth24 = (u32)(packed_data >> TG_VIS_DEPTH_SHIFT);
        const u32 local_pointer = ((u32)(packed_data >> TG_VIS_POINTER_SHIFT) & 2147483647u) - global_pointer_base;
        u64 word = 0;
        if ((f32)depth24 / TG_VIS_DEPTH_SCALE < 1.0f && local_pointer < n_local_pointers)
        {
            const u32 cluster_idx = __ldg(&p_cluster_pointers[local_pointer]);
            const u32 object_idx = __ldg(&p_c2o[cluster_idx]);
            const u32 color_lut_idx = __ldg(&p_lut_idx[(u64)cluster_idx * 512u + ((u32)packed_data & 511u)]);
            const u32 packed_color = __ldg(&p_color_lut[p_objects[object_idx].lut_idx * 256u + color_lut_idx]);
            word = ((u64)(global_object_base + object_idx + 1u) << 32) | (u64)packed_color;
        }
        p_mat[i] = word;
    }
}

/* ---- present pass: present.frag:9-12 + the swapchain's format conversion --------------------------------------- */
/*
 * The reference ends a frame by sampling the HDR target 1:1 into the swapchain image (present.frag: out_color = texture(..);
 * tgvk_raytracer.c:560-630,1524-1553), whose format is VK_FORMAT_B8G8R8A8_UNORM (tgvk_core.c:4239-4246). The conversion
 * is the fixed-function float -> UNORM8 one: NaN -> 0, clamp to [0, 1], round(c * 255) to nearest (ties to even); memory
 * order B, G, R, A = one little-endian u32  a << 24 | r << 16 | g << 8 | b. One thread = one pixel, 16 B in, 4 B out.
 */
__device__ __forceinline__ u32 tgb_unorm8(f32 c)
{
    f32 v = c > 0.0f ? c : 0.0f; /* NaN and negatives -> 0 */
    v = v > 1.0f ? 1.0f : v;
    return __float2uint_rn(v * 255.0f);
}

__global__ void __launch_bounds__(256) k_present(const float4* __restrict__ p_radiance, u32* __restrict__ p_out, u64 first, u64 n)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 c = p_radiance[first + i];
    p_out[first + i] = (tgb_unorm8(c.w) << 24) | (tgb_unorm8(c.x) << 16) | (tgb_unorm8(c.y) << 8) | tgb_unorm8(c.z);
}

static b32 tgbd__present_buffer(struct tgb_device* d, u32 buf)
{
    if (!d->d_present_pair[buf])
    {
        const u64 n_bytes = (u64)d->width * d->tile_rows * (d->n_ranks ? d->n_ranks : 1) * sizeof(u32);
        TGB_CUDA(cudaMalloc(&d->d_present_pair[buf], n_bytes));
    }
    return TG_TRUE;
}

/* the whole frame (the current radiance buffer) presented into caller memory; synchronous */
extern "C" b32 tgbd_read_present(struct tgb_device* d, u32* p_out)
{
    TGB_CUDA(cudaSetDevice(d->device));
    if (!tgbd__present_buffer(d, d->radiance_flip)) return TG_FALSE;
    const u64 n = (u64)d->width * d->tile_rows * (d->n_ranks ? d->n_ranks : 1); /* the padded frame, virtual row order; p_out must hold it */
    k_present<<<(u32)((n + 255) / 256), 256, 0, d->stream>>>(d->d_radiance, d->d_present_pair[d->radiance_flip], 0, n);
    TGB_LAUNCH_CHECK(d);
    TGB_CUDA(cudaMemcpyAsync(p_out, d->d_present_pair[d->radiance_flip], n * sizeof(u32), cudaMemcpyDeviceToHost, d->stream));
    TGB_CUDA(cudaStreamSynchronize(d->stream));
    return TG_TRUE;
}

static b32 tgbd__shade_launch(struct tgb_device* d, const tg_camera_rays* p_cam, bool resolved, u32 n_local_pointers, u32 gi_enabled, u32 frame_seed, u32 debug_visualization, u32 y0, u32 y1)
{
    if (d->p_sink)
    {
        /* double-buffered radiance: this frame goes to the buffer the frame before last used, whose copies are (long) done */
        if (!d->d_radiance_pair[1])
        {
            const u64 n_bytes = (u64)d->width * d->tile_rows * (d->n_ranks ? d->n_ranks : 1) * sizeof(float4);
            TGB_CUDA(cudaMalloc(&d->d_radiance_pair[1], n_bytes));
            TGB_CUDA(cudaMemsetAsync(d->d_radiance_pair[1], 0, n_bytes, d->stream));
        }
        d->radiance_flip ^= 1u;
        d->d_radiance = d->d_radiance_pair[d->radiance_flip];
    }
    const u32 buf = d->radiance_flip;
    const bool present = d->p_sink && d->sink_format == TGB200_SINK_BGRA8;
    if (present && !tgbd__present_buffer(d, buf)) return TG_FALSE;
    tgb_shade_args a;
    /* after the peer-memory merge the merged words of this rank's tile live in d_vis_tile, addressed with whole-frame pixel indices */
    a.p_vis = (resolved && d->tile_merged && !d->vis_merged) ? d->d_vis_tile - (u64)d->rank * d->tile_rows * d->width : d->d_vis;
    a.p_out = d->d_radiance;
    a.p_cluster_pointers = d->d_cluster_pointers;
    a.p_c2o = d->d_c2o;
    a.p_objects = resolved ? d->d_objects_global : d->d_objects;
    a.p_frames = resolved ? d->d_frames_global : d->d_frames_all;
    a.p_lut_idx = d->d_lut_idx;
    a.p_color_lut = d->d_color_lut;
    a.svo.p_nodes = d->svo.d_nodes;
    a.svo.p_leaf_data = d->svo.d_leaf_data;
    a.svo.p_voxels = d->svo.d_voxels;
    a.svo.bmin = d->svo.bmin;
    a.svo.bmax = d->svo.bmax;
    a.cam = *p_cam;
    a.w = d->width; a.h = d->height;
    a.n_ranks = d->n_ranks ? d->n_ranks : 1; a.tile_rows = d->tile_rows;
    a.global_pointer_base = d->global_pointer_base;
    a.n_local_pointers = n_local_pointers;
    a.gi_enabled = gi_enabled; a.frame_seed = frame_seed; a.debug_visualization = debug_visualization;
    a.mat_y0 = y0;
    a.p_q0 = d->d_gi_q0; a.p_q1 = d->d_gi_q1; a.p_q2 = d->d_gi_q2; a.p_q_count = d->d_gi_count;
    a.p_mat = d->d_mat_tile;
    const bool gi = gi_enabled && debug_visualization == TG_DEBUG_SHOW_NONE;
    if (gi) { k_set_words<<<1, 32, 0, d->stream>>>(d->d_gi_count, 32, 0u); TGB_LAUNCH_CHECK(d); }

    /* box corners on the 32-unit lattice: the flattened tree is exact (k_gi_trace_flat); the stack kernel runs only when the tree could not be tabulated */
    bool flat = d->gi_traversal != 1;
    {
        const f32 c[6] = { a.svo.bmin.x, a.svo.bmin.y, a.svo.bmin.z, a.svo.bmax.x, a.svo.bmax.y, a.svo.bmax.z };
        for (int i = 0; i < 6; i++) flat = flat && fmodf(c[i], 32.0f) == 0.0f && fabsf(c[i]) <= 4194304.0f; /* corners on the 32-unit cell lattice */
    }
    const int gi_ctas = max(1, min(16, tgbd_env_int("TGB_GI_CTAS_PER_SM", TGB_GI_FLAT_CTAS_PER_SM))); /* persistent CTAs per SM (tuning only) */
    /* TGB_GI_KERNEL=4: the first TGB_GI_SHADE_STEPS cells of the certified walk are entered by k_shade itself (0: every ray is queued) */
    const int shade_min_ctas = tgbd_env_int("TGB_SHADE_MIN_CTAS", 5); /* measured: 0.904 ms for the stage with 5, 0.923 with 4 (profiles/r04k_sweep_full.jsonl) */
    a.fast_steps = 0; a.fast_delta = tgbd_gi_fast_delta();
    if (gi && flat && tgbd_env_int("TGB_GI_KERNEL", TGB_GI_KERNEL_DEFAULT) == 4)
    {
        a.fast_steps = (u32)max(0, min(64, tgbd_env_int("TGB_GI_SHADE_STEPS", 1)));
        if (a.fast_steps)
        {
            if (!d->svo.fast_tiling_valid && !tgbd_gi_fast_tiling_build(d, d->stream)) return TG_FALSE;
            tgb_gi_frame_init(&a.fast_frame, d->svo.bmin, d->svo.bmax, p_cam->far_plane, d->svo.d_top_grid, d->svo.d_voxels);
            tgbd_gi_fast_tiling_get(d, &a.fast_tiling);
        }
    }

    /*
     * With a frame sink the rows are shaded in bands and every finished band is copied to the caller's memory on the copy
     * stream while the next band is shaded (the 133 MB RGBA32F frame takes as long over PCIe as the whole frame takes to
     * render). A band that would overwrite rows whose copy from the previous frame is still in flight waits for that copy.
     */
    const u32 n_bands = d->p_sink ? d->sink_bands : 1u;
    const u32 band_rows = (((y1 - y0) + n_bands - 1u) / n_bands + 15u) & ~15u; /* whole 16-row tiles of k_shade */
    b32 new_pending[TGB_MAX_BANDS];
    for (u32 k = 0; k < TGB_MAX_BANDS; k++) new_pending[k] = TG_FALSE;
    for (u32 b = 0; b < n_bands; b++)
    {
        const u32 by0 = y0 + b * band_rows;
        if (by0 >= y1) break;
        const u32 by1 = by0 + band_rows < y1 ? by0 + band_rows : y1;
        for (u32 k = 0; k < TGB_MAX_BANDS; k++)
        {
            if (d->band_copy_pending[buf][k] && d->band_row0[buf][k] < by1 && by0 < d->band_row1[buf][k])
            {
                TGB_CUDA(cudaStreamWaitEvent(d->stream, d->ev_band_copied[buf][k], 0));
                d->band_copy_pending[buf][k] = TG_FALSE;
            }
        }
        a.y0 = by0; a.y1 = by1;
        if (gi && b > 0) { k_set_words<<<1, 32, 0, d->stream>>>(d->d_gi_count, 2, 0u); TGB_LAUNCH_CHECK(d); } /* queued / fetched; the work counters accumulate over the bands */
        const dim3 grid((d->width + 15) / 16, (by1 - by0 + 15) / 16);
        /* FAST: 4 CTAs per SM (64 registers) or 5 (48 registers, a few spilled words): TGB_SHADE_MIN_CTAS, measured */
        if (a.fast_steps && shade_min_ctas >= 5) { if (resolved) k_shade<true, true, 5><<<grid, 256, 0, d->stream>>>(a); else k_shade<false, true, 5><<<grid, 256, 0, d->stream>>>(a); }
        else if (a.fast_steps)                    { if (resolved) k_shade<true, true, 4><<<grid, 256, 0, d->stream>>>(a); else k_shade<false, true, 4><<<grid, 256, 0, d->stream>>>(a); }
        else                                      { if (resolved) k_shade<true, false, 1><<<grid, 256, 0, d->stream>>>(a); else k_shade<false, false, 1><<<grid, 256, 0, d->stream>>>(a); }
        TGB_LAUNCH_CHECK(d);
        if (gi)
        {
            /* persistent: a few CTAs per SM, each lane pulls rays until the queue is empty (count read on the device) */
            /* TGB_GI_KERNEL: 4 (default) = the certified fast walk over the coarser tiling (tgb_gi_fast.cu; k_shade entered its first cell above), the shader's
             * own arithmetic only on the rays it hands over; 3 = its predecessor over the octree's cells (measured slower than 2: profiles/r03e_*);
             * 2 = the exact kernel on every ray, several rays per lane (tgb_gi_pool.cu): the reference frame of the others; 1 = the exact kernel, one ray per
             * lane (k_gi_trace_flat). All produce the same radiance bits. */
            const int gi_kernel = tgbd_env_int("TGB_GI_KERNEL", TGB_GI_KERNEL_DEFAULT);
            if (flat && (gi_kernel == 3 || gi_kernel == 4))
            {
                if (!tgbd_gi_fast_trace(d, p_cam->far_plane, gi_kernel == 4)) return TG_FALSE;
            }
            else if (flat && gi_kernel == 2)
            {
                if (!tgbd_gi_pool_trace(d, p_cam->far_plane)) return TG_FALSE;
            }
            else if (flat)
            {
                /* scheduling knobs (tuning only): DDA steps per phase, lanes that trigger a service phase, bias of the majority vote towards the DDA */
                const int dda_steps = tgbd_env_int("TGB_GI_DDA_STEPS", TGB_FL_DDA_STEPS);
                const u32 service_lanes = (u32)tgbd_env_int("TGB_GI_SERVICE_LANES", (i32)TGB_FL_SERVICE_LANES), dda_bias = (u32)tgbd_env_int("TGB_GI_DDA_BIAS", 0),
                                 tree_reps = (u32)max(1, tgbd_env_int("TGB_GI_TREE_REPS", TGB_FL_TREE_REPS));
                const dim3 gi_grid(d->n_sms * (u32)gi_ctas);
#define TGB_GI_LAUNCH(K) k_gi_trace_flat<K><<<gi_grid, TGB_GI_THREADS, 0, d->stream>>>(a.svo, d->svo.d_top_grid, p_cam->far_plane, d->d_gi_q0, d->d_gi_q1, d->d_gi_q2, \
                                                                                      d->d_gi_count, d->d_radiance, service_lanes, dda_bias, tree_reps)
                if (dda_steps <= 4) TGB_GI_LAUNCH(4); else if (dda_steps <= 8) TGB_GI_LAUNCH(8); else if (dda_steps <= 16) TGB_GI_LAUNCH(16); else TGB_GI_LAUNCH(64);
#undef TGB_GI_LAUNCH
                TGB_LAUNCH_CHECK(d);
            }
            k_gi_trace<<<d->n_sms * 8, TGB_GI_THREADS, 0, d->stream>>>(a.svo, flat ? d->svo.d_top_grid + TGB_TOP_GRID_CELLS : NULL, p_cam->far_plane, d->d_gi_q0, d->d_gi_q1, d->d_gi_q2,
                                                                       d->d_gi_count, d->d_radiance);
            TGB_LAUNCH_CHECK(d);
        }
        if (d->p_sink)
        {
            const u64 first_px = (u64)by0 * d->width, n_px = (u64)(by1 - by0) * d->width;
            if (present)
            {
                k_present<<<(u32)((n_px + 255) / 256), 256, 0, d->stream>>>(d->d_radiance, d->d_present_pair[buf], first_px, n_px);
                TGB_LAUNCH_CHECK(d);
            }
            TGB_CUDA(cudaEventRecord(d->ev_band[b], d->stream));
            TGB_CUDA(cudaStreamWaitEvent(d->copy_stream, d->ev_band[b], 0));
            if (present) TGB_CUDA(cudaMemcpyAsync((u32*)d->p_sink + (u64)(by0 - y0) * d->width, d->d_present_pair[buf] + first_px, n_px * sizeof(u32), cudaMemcpyDeviceToHost, d->copy_stream));
            else         TGB_CUDA(cudaMemcpyAsync(d->p_sink + (u64)(by0 - y0) * d->width * 4u, d->d_radiance + first_px, n_px * sizeof(float4), cudaMemcpyDeviceToHost, d->copy_stream));
            TGB_CUDA(cudaEventRecord(d->ev_band_copied[buf][b], d->copy_stream));
            new_pending[b] = TG_TRUE;
            /* an older copy still tracked under this index (the band layout changed): keep the union of the row ranges, the re-recorded event covers both */
            d->band_row0[buf][b] = d->band_copy_pending[buf][b] && d->band_row0[buf][b] < by0 ? d->band_row0[buf][b] : by0;
            d->band_row1[buf][b] = d->band_copy_pending[buf][b] && d->band_row1[buf][b] > by1 ? d->band_row1[buf][b] : by1;
        }
    }
    if (d->p_sink)
    {
        for (u32 k = 0; k < TGB_MAX_BANDS; k++) d->band_copy_pending[buf][k] = d->band_copy_pending[buf][k] || new_pending[k];
        d->n_frames_sunk++;
        TGB_CUDA(cudaEventRecord(d->ev_frame_copied[d->n_frames_sunk % TGB_FRAME_RING], d->copy_stream));
    }
    d->gi_stats_valid = gi ? TG_TRUE : TG_FALSE;
    return TG_TRUE;
}

extern "C" b32 tgbd_render_shading(struct tgb_device* d, const tg_camera_rays* p_cam, u32 n_local_pointers, u32 gi_enabled, u32 frame_seed, u32 debug_visualization,
                                   u32 y0, u32 y1)
{
    TGB_CUDA(cudaSetDevice(d->device));
    if (!tgbd_flush_objects(d)) return TG_FALSE;
    if (gi_enabled && debug_visualization == TG_DEBUG_SHOW_NONE && !d->svo.valid)
    {
        tgb_set_error("render_shading: GI is enabled but no SVO has been built or uploaded");
        return TG_FALSE;
    }
    if (y1 > d->height) y1 = d->height;
    if (y0 >= y1) return TG_TRUE;
    TGB_CUDA(cudaEventRecord(d->ev[7], d->stream));
    k_object_frames<<<(d->object_capacity + 127) / 128, 128, 0, d->stream>>>(d->d_objects, d->object_capacity,
                                                                            tgb_v3(p_cam->camera.x, p_cam->camera.y, p_cam->camera.z), d->d_frames_all);
    TGB_LAUNCH_CHECK(d);
    if (!tgbd__shade_launch(d, p_cam, false, n_local_pointers, gi_enabled, frame_seed, debug_visualization, y0, y1)) return TG_FALSE;
    TGB_CUDA(cudaEventRecord(d->ev[8], d->stream));
    d->ev_shade = TG_TRUE;
    return TG_TRUE;
}

extern "C" b32 tgbd_render_shading_sharded(struct tgb_device* d, const tg_camera_rays* p_cam, u32 n_local_pointers, u32 gi_enabled, u32 frame_seed, u32 debug_visualization)
{
    TGB_CUDA(cudaSetDevice(d->device));
    if (!tgbd_flush_objects(d)) return TG_FALSE;
    if (!d->p_comm || d->n_ranks < 2) { tgb_set_error("render_shading_sharded: no communicator"); return TG_FALSE; }
    if (gi_enabled && debug_visualization == TG_DEBUG_SHOW_NONE && !d->svo.valid)
    {
        tgb_set_error("render_shading: GI is enabled but no SVO has been built (call tgb200_svo_update on every rank)");
        return TG_FALSE;
    }
    const u32 cap = d->object_capacity, n_global = cap * d->n_ranks;
    const v3 camera = tgb_v3(p_cam->camera.x, p_cam->camera.y, p_cam->camera.z);
    const u64 tile_px = (u64)d->width * d->tile_rows, n_padded = tile_px * d->n_ranks, n_pixels = n_padded; /* padding rows hold the clear word */
    const bool fused = !d->vis_merged;
    if (fused && !d->p2p_ready)
    {
        tgb_set_error("render_shading: the visibility buffer is not merged (call tgb200_merge_visibility on every rank, or let tg_raytracer_render pick the peer-memory merge)");
        return TG_FALSE;
    }
    if (fused)
    {
        /* merge over peer memory (tgb_peer.cu). K1 flagged the tiles in which this rank has a hit: the material of this rank's LOCAL
         * winners is resolved for those tiles only; then publish "K1 + materials done", wait for the peers' counters, min + winner's
         * material for this rank's tile straight from the peers' buffers. Timed as the merge stage. Words that did not come from K1
         * (an uploaded buffer, the BLOCKS view's SVO pass) take the pass over the whole buffer and count every tile as hit; the object records travel
         * through the same peer-memory tail (tgbd_p2p_barrier). */
        tgbd_merge_begin(d);
        if (d->tiles_flagged)
        {
            const dim3 tiles((tgbd_tiles_x(d) + 7u) / 8u, d->tile_rows * d->n_ranks / TGB_BAND_ROWS);
            k_resolve_material_tiles<<<tiles, 256, 0, d->stream>>>(d->d_vis, d->width, tgbd_tiles_x(d), tgbd_mat_tile_flags(d, d->d_mat), d->d_cluster_pointers, d->d_c2o, d->d_objects,
                                                                   d->d_lut_idx, d->d_color_lut, d->global_pointer_base, n_local_pointers, d->rank * cap, d->d_mat);
            TGB_LAUNCH_CHECK(d);
        }
        else
        {
            k_resolve_material<<<(u32)((n_padded + 255) / 256), 256, 0, d->stream>>>(d->d_vis, n_pixels, n_padded, d->d_cluster_pointers, d->d_c2o, d->d_objects, d->d_lut_idx,
                                                                                     d->d_color_lut, d->global_pointer_base, n_local_pointers, d->rank * cap, d->d_mat);
            TGB_LAUNCH_CHECK(d);
            if (!tgbd_p2p_flag_all_tiles(d)) return TG_FALSE;
        }
        if (!tgbd_p2p_barrier(d)) return TG_FALSE; /* publishes this rank's object records + frame counter, records ev[11] (published) and ev[12] (every peer arrived), collects the peers' records */
        if (!tgbd_p2p_merge_tile(d)) return TG_FALSE;
        tgbd_merge_end(d);
        d->ev_merge_parts = TG_TRUE;
        TGB_CUDA(cudaEventRecord(d->ev[7], d->stream));
        k_object_frames<<<(n_global + 127) / 128, 128, 0, d->stream>>>(d->d_objects_global, n_global, camera, d->d_frames_global);
        TGB_LAUNCH_CHECK(d);
    }
    else
    {
        TGB_CUDA(cudaEventRecord(d->ev[7], d->stream));
        /* replicate the object records (96 B each) of every shard; pointers globalised by the owner */
        if (!d->objects_gathered && !tgbd_gather_objects(d)) return TG_FALSE;
        k_object_frames<<<(n_global + 127) / 128, 128, 0, d->stream>>>(d->d_objects_global, n_global, camera, d->d_frames_global);
        TGB_LAUNCH_CHECK(d);

        /* owner resolves the material of every pixel it won, max-reduce-scatter by screen tile */
        k_resolve_material<<<(u32)((n_padded + 255) / 256), 256, 0, d->stream>>>(d->d_vis, n_pixels, n_padded, d->d_cluster_pointers, d->d_c2o, d->d_objects, d->d_lut_idx,
                                                                                 d->d_color_lut, d->global_pointer_base, n_local_pointers, d->rank * cap, d->d_mat);
        TGB_LAUNCH_CHECK(d);
        if (!tgbn_reducescatter_max_u64(d->p_comm, d->d_mat, d->d_mat_tile, tile_px, d->stream)) return TG_FALSE;
    }

    /* GI + shading of this rank's rows */
    const u32 y0 = d->rank * d->tile_rows, y1 = y0 + d->tile_rows; /* virtual rows; k_shade skips the padding rows */
    if (y0 < y1 && !tgbd__shade_launch(d, p_cam, true, n_local_pointers, gi_enabled, frame_seed, debug_visualization, y0, y1)) return TG_FALSE;
    TGB_CUDA(cudaEventRecord(d->ev[8], d->stream));
    d->ev_shade = TG_TRUE;
    return TG_TRUE;
}

extern "C" b32 tgbd_gather_radiance(struct tgb_device* d)
{
    TGB_CUDA(cudaSetDevice(d->device));
    if (!d->p_comm || d->n_ranks < 2) return TG_TRUE;
    const u64 tile_bytes = (u64)d->width * d->tile_rows * sizeof(float4);
    return tgbn_allgather_bytes(d->p_comm, (const u8*)d->d_radiance + (u64)d->rank * tile_bytes, d->d_radiance, tile_bytes, d->stream);
}

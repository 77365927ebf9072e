"""CPU: a SECOND, independent statement of shading.frag:116-316 (G1, G2, G5) -- float64 numpy written straight from the GLSL text,
matrices instead of the oracle's hoisted chain, no shared code with oracle/tgo_shade.c -- against the oracle on a whole small frame.
shading.frag has no CPU twin in the reference and cannot run here, so this is what guards its transcription against formula and
sign mistakes: every hit pixel must agree to 1e-4 relative except the handful whose float32 face-normal choice sits on a tie."""
import numpy as np

from tg_b200 import scenes

CLEAR = 0xFFFFFFFFFFFFFFFF
PI = 3.14159265358979323846


def normalize(v):
    return v / np.sqrt(v @ v)


def ray_aabb(o, d, bmin, bmax):
    """collide.inc:3-24"""
    with np.errstate(divide="ignore", invalid="ignore"):
        v0 = np.where(d == 0.0, -3.402823466e38, (bmin - o) / np.where(d == 0.0, 1.0, d))
        v1 = np.where(d == 0.0, 3.402823466e38, (bmax - o) / np.where(d == 0.0, 1.0, d))
    enter, exit_ = np.minimum(v0, v1).max(), np.maximum(v0, v1).min()
    return (exit_ > 0.0 and enter <= exit_), enter


def tg_shade(n, v, l, diffuse_albedo, specular_albedo, metallic, roughness, radiance):
    """shading.frag:53-110"""
    h = normalize(v + l)
    h_dot_n, l_dot_n, n_dot_v = (np.clip(x, 0.0, 1.0) for x in (h @ n, n @ l, n @ v))
    a_sqr = (roughness * roughness) ** 2
    denom = h_dot_n * h_dot_n * (a_sqr - 1.0) + 1.0
    d = a_sqr / (PI * denom * denom)
    f = (1.0 - roughness) + roughness * (1.0 - n_dot_v) ** 5
    k = (roughness + 1.0) ** 2 / 8.0
    g = (l_dot_n / (l_dot_n * (1.0 - k) + k)) * (n_dot_v / (n_dot_v * (1.0 - k) + k))
    specular = specular_albedo * (d * f * g / max(4.0 * n_dot_v * l_dot_n, 0.001))
    diffuse = (1.0 - f) * (1.0 - metallic) * diffuse_albedo / PI
    return (diffuse + specular) * radiance * l_dot_n


def second_opinion(oracle, scene):
    view = oracle.SceneView.from_scene(scene, with_lut=True)
    rays = oracle.camera_rays(oracle.camera_from_spec(scene.camera))
    w, h = scene.width, scene.height
    vis, _ = oracle.visibility(view, rays, w, h, oracle.VIS_SCREEN_RECT)
    want = oracle.shade(view, rays, w, h, vis, None, gi=False)
    cam = np.array([rays.camera.x, rays.camera.y, rays.camera.z], dtype=np.float64)
    corner = {k: np.array([getattr(rays, k).x, getattr(rays, k).y, getattr(rays, k).z], dtype=np.float64) for k in ("ray_bl", "ray_br", "ray_tr", "ray_tl")}
    far = float(rays.far_plane)
    od = view.object_data
    got = np.zeros((h, w, 4), dtype=np.float64)
    for py in range(h):
        for px in range(w):
            word = int(vis[py, px])
            depth = (word >> 40) / 16777215.0
            if not depth < 1.0:
                got[py, px] = (1.0, 0.0, 1.0, 1.0)                                     # :335
                continue
            pointer, voxel = (word >> 9) & 0x7FFFFFFF, word & 511
            cluster = int(view.cluster_pointers[pointer])                              # :119-124
            obj = od[int(view.c2o[cluster])]
            packed = int(view.color_lut[int(obj["lut_idx"]) * 256 + int(view.lut_idx[cluster, voxel])])   # :128-131
            albedo = np.array([(packed >> 24) & 255, (packed >> 16) & 255, (packed >> 8) & 255], dtype=np.float64) / 255.0
            nx, ny, _ = (int(x) for x in obj["dims"])
            rel = pointer - int(obj["first_cluster_pointer"])
            offset = 8.0 * np.array([rel % nx, (rel // nx) % ny, rel // (nx * ny)], dtype=np.float64)
            half = 4.0 * obj["dims"].astype(np.float64)
            rot = obj["rotation"].astype(np.float64).reshape(4, 4).T[:3, :3]           # column-major mat4 -> 3x3
            inv_rot = np.linalg.inv(rot)
            translation = obj["translation"].astype(np.float64)
            fx, fy = (px + 0.5) / w, 1.0 - (py + 0.5) / h                               # :144-145
            mix = lambda a, b, t: a * (1.0 - t) + b * t
            dir_ws = mix(mix(corner["ray_bl"], corner["ray_tl"], fy), mix(corner["ray_br"], corner["ray_tr"], fy), fx)
            o_ms = inv_rot @ (cam - translation) + half - offset                       # :166-178
            d_ms = normalize(inv_rot @ dir_ws)
            vmin = np.array([voxel % 8, (voxel // 8) % 8, voxel // 64], dtype=np.float64)
            normal_ws = np.zeros(3)
            hit, enter = ray_aabb(o_ms, d_ms, vmin, vmin + 1.0)
            if hit:                                                                    # :196-229
                n = (o_ms + enter * d_ms if enter > 0.0 else o_ms) - (vmin + 0.5)
                if abs(n[0]) > abs(n[1]):
                    n = np.array([np.sign(n[0]), 0.0, 0.0]) if abs(n[0]) > abs(n[2]) else np.array([0.0, 0.0, np.sign(n[2])])
                else:
                    n = np.array([0.0, np.sign(n[1]), 0.0]) if abs(n[1]) > abs(n[2]) else np.array([0.0, 0.0, np.sign(n[2])])
                normal_ws = normalize(rot @ n)
            hit_ws = cam + depth * far * dir_ws                                        # :231, un-normalised direction (Q3)
            v = normalize(cam - hit_ws)                                                # :301-316
            l = normalize(np.array([0.0, 0.8, 0.3]))
            specular_albedo = 0.04 * (1.0 - 0.1) + albedo * 0.1
            lo = tg_shade(normal_ws, v, l, albedo, specular_albedo, 0.1, 0.8, np.array([3.0, 3.0, 3.0]))
            got[py, px, :3] = 0.1 * albedo + lo
            got[py, px, 3] = 1.0
    return got, want, vis


def test_shading_transcription_against_an_independent_float64_statement(oracle):
    s = scenes.small_grid(grid=3, width=96, height=54)
    for i, o in enumerate(s.objects):
        o.lut_indices = scenes.random_lut_indices(o.seed, o.n_clusters)
    got, want, vis = second_opinion(oracle, s)
    hit = vis != np.uint64(CLEAR)
    assert hit.sum() > 1500
    assert np.array_equal(got[~hit], want[~hit].astype(np.float64))
    close = np.isclose(got, want, rtol=1e-4, atol=1e-6).all(axis=-1)
    # float32 vs float64 may pick a different face on a voxel edge (|n.x| == |n.y| to within rounding): a handful of pixels at most
    assert (~close[hit]).sum() <= max(3, hit.sum() // 500), f"{int((~close[hit]).sum())} of {int(hit.sum())} hit pixels disagree"

"""CPU: a SECOND, independent statement of what visibility.frag computes (V3-V7), against the oracle's transcription on a frame of
ROTATED objects (the reference-run cross-check in test_reference_pins.py covers axis-aligned ones only). No DDA here: in float64,
for every pixel and every cluster, the slab test of the ray against EVERY solid voxel; the visible voxel of the cluster is the one
entered first, its depth max(0, enter / far), the pixel's word the minimum over the clusters. The Amanatides-Woo march of the
shader must find exactly that voxel except where float32 stepping and exact geometry part ways on a voxel edge or corner."""
import numpy as np

from tg_b200 import scenes

CLEAR = np.uint64(0xFFFFFFFFFFFFFFFF)


def brute_force_visibility(oracle, scene):
    view = oracle.SceneView.from_scene(scene, with_lut=False)
    rays = oracle.camera_rays(oracle.camera_from_spec(scene.camera))
    w, h = scene.width, scene.height
    cam = np.array([rays.camera.x, rays.camera.y, rays.camera.z], dtype=np.float64)
    c = {k: np.array([getattr(rays, k).x, getattr(rays, k).y, getattr(rays, k).z], dtype=np.float64) for k in ("ray_bl", "ray_br", "ray_tr", "ray_tl")}
    far = float(rays.far_plane)
    px, py = np.meshgrid(np.arange(w), np.arange(h))
    fx, fy = ((px + 0.5) / w).reshape(-1, 1), (1.0 - (py + 0.5) / h).reshape(-1, 1)
    mix = lambda a, b, t: a * (1.0 - t) + b * t
    dir_ws = mix(mix(c["ray_bl"], c["ray_tl"], fy), mix(c["ray_br"], c["ray_tr"], fy), fx)          # [P, 3], un-normalised
    best = np.full(w * h, CLEAR, dtype=np.uint64)
    best_depth = np.full(w * h, np.inf)
    for obj in view.object_data:
        nx, ny, nz = (int(x) for x in obj["dims"])
        rot = obj["rotation"].astype(np.float64).reshape(4, 4).T[:3, :3]
        inv_rot = np.linalg.inv(rot)
        d_ms = dir_ws @ inv_rot.T
        d_ms /= np.sqrt((d_ms * d_ms).sum(axis=1, keepdims=True))
        o_obj = inv_rot @ (cam - obj["translation"].astype(np.float64)) + 4.0 * obj["dims"].astype(np.float64)
        for rel in range(nx * ny * nz):
            pointer = int(obj["first_cluster_pointer"]) + rel
            mask = view.masks[int(view.cluster_pointers[pointer])]
            bits = np.unpackbits(mask.view(np.uint8), bitorder="little")                         # bit 64 z + 8 y + x
            solid = np.flatnonzero(bits)
            if solid.size == 0:
                continue
            vmin = np.stack([solid % 8, (solid // 8) % 8, solid // 64], axis=1).astype(np.float64)   # [K, 3]
            o_ms = o_obj - 8.0 * np.array([rel % nx, (rel // nx) % ny, rel // (nx * ny)], dtype=np.float64)
            with np.errstate(divide="ignore", invalid="ignore"):
                t0 = (vmin[None] - o_ms) / d_ms[:, None, :]                                          # [P, K, 3]
                t1 = (vmin[None] + 1.0 - o_ms) / d_ms[:, None, :]
            enter = np.minimum(t0, t1).max(axis=2)
            exit_ = np.maximum(t0, t1).min(axis=2)
            crossed = (exit_ > 0.0) & (enter < exit_)                                                # the ray passes THROUGH the voxel
            enter = np.where(crossed, enter, np.inf)
            k = enter.argmin(axis=1)
            e = enter[np.arange(len(k)), k]
            depth = np.maximum(0.0, e / far)
            ok = np.isfinite(e) & (depth <= 1.0)
            word = (np.floor(np.where(ok, depth, 0.0) * 16777215.0).astype(np.uint64) << np.uint64(40)) | np.uint64(pointer << 9) | solid[k].astype(np.uint64)
            better = ok & (word < best)
            best = np.where(better, word, best)
            best_depth = np.where(better, depth, best_depth)
    return best.reshape(h, w), view, rays


def test_visibility_transcription_against_brute_force_geometry(oracle):
    s = scenes.small_grid(grid=2, width=96, height=54, dims=(3, 2, 3))
    s.camera = scenes.CameraSpec(position=(0.0, 34.0, 26.0), pitch=float(scenes.deg2rad(-42.0)), yaw=float(scenes.deg2rad(9.0)), roll=0.0, aspect=96 / 54)
    got, view, rays = brute_force_visibility(oracle, s)
    want, _ = oracle.visibility(view, rays, s.width, s.height, oracle.VIS_BRUTE_FORCE)
    hit = want != CLEAR
    assert hit.sum() > 1500, int(hit.sum())
    assert np.array_equal(got == CLEAR, ~hit) or ((got == CLEAR) != ~hit).sum() <= 3        # silhouette pixels may graze
    both = hit & (got != CLEAR)
    same_voxel = (got & np.uint64((1 << 40) - 1)) == (want & np.uint64((1 << 40) - 1))      # cluster pointer and voxel
    depth_close = np.abs((got >> np.uint64(40)).astype(np.int64) - (want >> np.uint64(40)).astype(np.int64)) <= 1
    bad = both & ~(same_voxel & depth_close)
    assert bad.sum() <= max(3, both.sum() // 200), f"{int(bad.sum())} of {int(both.sum())} hit pixels see another voxel than brute-force geometry"

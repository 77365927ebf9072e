"""Shared helpers of the parity tests: run one scene through the CUDA path (via the C ABI) and the oracle."""
import numpy as np

CLEAR = np.uint64(0xFFFFFFFFFFFFFFFF)


def oracle_visibility(O, scene, mode=None, y0=0, y1=None, ystep=1, base=0):
    cam = O.camera_from_spec(scene.camera)
    rays = O.camera_rays(cam)
    view = O.SceneView.from_scene(scene, base, with_lut=False)
    vis, n = O.visibility(view, rays, scene.width, scene.height, O.VIS_SCREEN_RECT if mode is None else mode, y0, y1, ystep)
    return vis


def gpu_visibility(scene, base=0):
    from tg_b200.raytracer import from_scene
    rt = from_scene(scene)
    try:
        if base:
            rt.set_shard(0, 1, base)
        rt.clear()
        rt.render_visibility()
        rt.synchronize()
        return rt.read_visibility(), rt.timings()
    finally:
        rt.destroy()


def describe_mismatch(a, b, limit=5):
    bad = np.argwhere(a != b)
    lines = [f"{len(bad)} of {a.size} pixels differ"]
    for y, x in bad[:limit]:
        lines.append(f"  ({x},{y}): got {int(a[y, x]):#018x} want {int(b[y, x]):#018x}")
    return "\n".join(lines)

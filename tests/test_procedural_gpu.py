"""GPU: the reference's UNMODIFIED object constructor, tg_raytracer_create_object(center, extent)
(graphics/vulkan/tgvk_raytracer.c:805-992): simplex-noise terrain generated on the device (tg_b200/csrc/tgb_procedural.cu)
against the oracle's restatement (oracle/tgo_procedural.c, itself pinned bit for bit to the reference's own
tgm_simplex_noise by tests/test_reference_pins.py), and the reference application's sample scene
(tg_application.c:49-98) rendered through exactly the calls the application makes."""
import ctypes as C

import numpy as np
import pytest

from tg_b200 import ctypes_defs as T
from tg_b200 import scenes
from tg_b200.raytracer import Raytracer, from_scene

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("object_idx,dims", [(0, (16, 4, 16)), (1, (4, 4, 4)), (7, (3, 5, 2)), (123, (2, 9, 1))])
def test_procedural_bits_bit_exact(gpu, oracle, object_idx, dims):
    import tg_b200
    n = dims[0] * dims[1] * dims[2]
    got = np.zeros((n, 16), dtype=np.uint32)
    tg_b200.lib().tgb200_procedural_solid_bits(object_idx, T.v3u(*dims), T.ptr(got, T.u32))
    tg_b200.check()
    want = oracle.procedural_solid_bits(object_idx, dims)
    assert np.array_equal(got, want), f"{int((got != want).sum())} of {got.size} words differ"
    dens = np.unpackbits(want.view(np.uint8)).mean()
    assert 0.02 < dens < 0.98  # terrain, not a constant


def test_reference_application_scene_through_the_reference_calls(gpu, oracle):
    """tg_raytracer_create(&camera, 1 << 12, 1 << 21, &rt); 10 x tg_raytracer_create_object; 256 x tg_raytracer_color_lut_set;
    clear(); render() -- then every product of the frame against the oracle on the same scene."""
    w, h = 480, 270
    s = scenes.reference_app_scene(w, h, oracle.procedural_solid_bits)
    cam = T.make_camera(s.camera.position, s.camera.pitch, s.camera.yaw, s.camera.roll, s.camera.fov_y_deg, s.camera.aspect, s.camera.near, s.camera.far)
    rt = Raytracer(cam, 1 << 12, 1 << 21, w, h)
    try:
        for center, extent in scenes.reference_app_objects():
            rt.create_object(center, extent)
        for i, (r, g, b) in enumerate(scenes.reference_lut_ramp(256)):
            rt.color_lut_set(i, r, g, b)
        sc = rt.scene
        assert sc.n_objects == 10 and sc.n_cluster_pointers == 1024 + 9 * 64  # 1600 clusters
        # the CPU mirror the reference keeps (tg_scene.p_voxel_cluster_data, tgvk_raytracer.c:934) holds the generated bits
        mirror = np.ctypeslib.as_array(sc.p_voxel_cluster_data, shape=(sc.n_cluster_pointers, 16))
        want_bits = np.concatenate([o.bits for o in s.objects])
        assert np.array_equal(mirror, want_bits)
        angles = [sc.p_objects[i].angle_in_radians for i in range(10)]
        assert angles[0] == np.float32(scenes.deg2rad(15.0)) and angles[3] == np.float32(scenes.deg2rad(21.0))  # tgvk_raytracer.c:830-831

        rt.set_gi(True, 1)
        rt.clear(); rt.render(); rt.synchronize()
        vis, rad = rt.read_visibility(), rt.read_radiance()
        svo, nodes, leaf, vox = rt.svo_download()
        rt.svo_free(svo)
        hit, depth, cluster, voxel = rt.get_hovered_voxel(w // 2, h // 2)
    finally:
        rt.destroy()
    rays = oracle.camera_rays(oracle.camera_from_spec(s.camera))
    view = oracle.SceneView.from_scene(s, with_lut=True)
    want_vis, _ = oracle.visibility(view, rays, w, h, oracle.VIS_SCREEN_RECT)
    assert np.array_equal(vis, want_vis), f"{int((vis != want_vis).sum())} visibility words differ"
    assert (vis != np.uint64(0xFFFFFFFFFFFFFFFF)).mean() > 0.2
    osvo = oracle.svo_create(view)
    wn, wl, wv = oracle.svo_arrays(osvo)
    assert np.array_equal(nodes, wn) and np.array_equal(leaf, wl) and np.array_equal(vox, wv)
    want_rad = oracle.shade(view, rays, w, h, want_vis, osvo, gi=True, frame_seed=1)
    oracle.svo_destroy(osvo)
    assert np.allclose(rad, want_rad, rtol=1e-3, atol=1e-6)
    # tgvk_raytracer.c:1639-1654 on the centre pixel
    word = int(want_vis[h // 2, w // 2])
    assert hit == ((word >> 40) / 16777215.0 < 1.0)
    if hit:
        assert cluster == (word >> 9) & 0x7FFFFFFF and voxel == word & 511 and depth == np.float32(np.float32(word >> 40) / np.float32(16777215.0))


def test_synthetic_objects_generated_on_device_equal_uploaded_ones(gpu, oracle):
    """tg_raytracer_create_object_synthetic (k_synthetic_fill, BASELINE configs[4]'s device-side generation): the CPU mirror the
    library reads back holds exactly scenes.random_solid_bits, and the frame of a device-generated scene equals both the frame of
    the same scene uploaded from host arrays and the oracle's."""
    import ctypes as C
    s_host = scenes.small_grid(grid=4, dims=(5, 3, 4))
    s_dev = scenes.grid_scene("small_dev", 4, 4, s_host.width, s_host.height, k=3, pitch_units=8.0 * 5 * 1.5, dims=(5, 3, 4), with_bits=False)
    s_dev.camera = s_host.camera
    frames = []
    for s in (s_host, s_dev):
        rt = from_scene(s)
        try:
            n = s.n_clusters
            mirror = np.ctypeslib.as_array(rt.scene.p_voxel_cluster_data, shape=(n * 16,)).reshape(n, 16).copy()
            want = np.concatenate([scenes.random_solid_bits(o.seed, o.n_clusters, o.k) for o in s.objects])
            assert np.array_equal(mirror, want)
            rt.set_gi(True, 1)
            rt.clear(); rt.render(); rt.synchronize()
            frames.append((rt.read_visibility(), rt.read_radiance()))
        finally:
            rt.destroy()
    assert np.array_equal(frames[0][0], frames[1][0]) and np.array_equal(frames[0][1], frames[1][1])
    rays = oracle.camera_rays(oracle.camera_from_spec(s_host.camera))
    want_vis, _ = oracle.visibility(oracle.SceneView.from_scene(s_host, with_lut=True), rays, s_host.width, s_host.height, oracle.VIS_SCREEN_RECT)
    assert np.array_equal(frames[1][0], want_vis)

"""CPU: K1's per-ray walk in its resumable form (tg_b200/csrc/tgb_k1_walk.cuh: tgb_k1_setup / tgb_k1_next_candidate /
tgb_cluster_march -- what k_visibility_pool runs per ray slot) compiled for the host (tests/cpu_sim) and held against the oracle's
visibility buffer: bit-exact, every pixel, on rotated multi-object scenes, a single big object and a sharded pointer base."""
import numpy as np
import pytest

from tg_b200 import scenes
from tests import cpu_sim
from tests.helpers import describe_mismatch


@pytest.mark.parametrize("make,base", [(lambda: scenes.small_grid(), 0), (lambda: scenes.small_grid(grid=4, width=333, height=177, dims=(3, 5, 2)), 4096),
                                       (lambda: scenes.config1(k=3, width=240, height=135), 0), (lambda: scenes.config1(k=1, width=160, height=90, dims=(4, 4, 4)), 0)])
def test_resumable_walk_equals_the_oracle(oracle, make, base):
    s = make()
    rays = oracle.camera_rays(oracle.camera_from_spec(s.camera))
    view = oracle.SceneView.from_scene(s, base, with_lut=False)
    want, _ = oracle.visibility(view, rays, s.width, s.height, oracle.VIS_SCREEN_RECT)
    got, work = cpu_sim.visibility(view, rays, s.width, s.height)
    assert np.array_equal(got, want), describe_mismatch(got, want)
    # k_visibility's deferred word: the walk continues on an upper bound of t_skip, the exact word is computed at the end of the object
    deferred, work_d = cpu_sim.visibility(view, rays, s.width, s.height, defer=True)
    assert np.array_equal(deferred, want), describe_mismatch(deferred, want)
    assert work_d[0] >= work[0] and work_d[0] < 1.05 * work[0] + 50, (work, work_d)   # the looser bound admits few extra candidates
    assert work[0] > 0 and (want != np.uint64(0xFFFFFFFFFFFFFFFF)).sum() > 100


def test_camera_inside_an_object_and_grazing_views(oracle):
    """enter <= 0 (camera inside the grid), axis-parallel view directions (zero direction components after the rotation is undone)."""
    s = scenes.config1(k=3, width=96, height=54, dims=(4, 4, 4))
    s.objects[0].angle = 0.0
    for pos, pitch in (((0.3, 0.2, 0.1), 0.0), ((0.0, 0.0, 40.0), 0.0), ((0.0, 40.0, 0.0), -1.5707963), ((3.5, 2.5, 1.5), 0.4)):
        s.camera.position, s.camera.pitch = pos, pitch
        rays = oracle.camera_rays(oracle.camera_from_spec(s.camera))
        view = oracle.SceneView.from_scene(s, with_lut=False)
        want, _ = oracle.visibility(view, rays, s.width, s.height, oracle.VIS_BRUTE_FORCE)
        for defer in (False, True):
            got, _ = cpu_sim.visibility(view, rays, s.width, s.height, defer=defer)
            assert np.array_equal(got, want), f"camera {pos} defer={defer}: " + describe_mismatch(got, want)

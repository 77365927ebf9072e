"""Writes tests/golden/*.npz from the oracle (run here, commit the output). The GPU parity tests compare the
CUDA path with these fixtures as well as with the live oracle, so a drifting oracle cannot silently move the target.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tg_b200 import scenes  # noqa: E402

CASES = {
    "config1_k3_320x180": lambda: scenes.config1(k=3, width=320, height=180),
    "config1_k1_320x180": lambda: scenes.config1(k=1, width=320, height=180),
    "small_grid3_320x180": lambda: scenes.small_grid(),
}


def compute(O, name):
    s = CASES[name]()
    cam = O.camera_from_spec(s.camera)
    rays = O.camera_rays(cam)
    view = O.SceneView.from_scene(s, with_lut=True)
    vis, _ = O.visibility(view, rays, s.width, s.height, O.VIS_SCREEN_RECT)
    svo = O.svo_create(view)
    nodes, leaf, vox = O.svo_arrays(svo)
    rad = O.shade(view, rays, s.width, s.height, vis, svo, gi=True, frame_seed=1)
    O.svo_destroy(svo)
    # the radiance is kept at reduced resolution (every 4th pixel) to keep the fixture small
    return dict(vis=vis, svo_nodes=nodes, svo_leaf=leaf, svo_voxels_nonzero_idx=np.nonzero(vox)[0].astype(np.uint32),
                svo_voxels_nonzero=vox[np.nonzero(vox)[0]], radiance_4=rad[::4, ::4].copy())


if __name__ == "__main__":
    from oracle import oracle as O
    for name in CASES:
        out = compute(O, name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})

"""Writes tests/golden/ref_*.npz from THE REFERENCE'S OWN CODE (oracle/_ref/libtg_ref.so = the reference's portable C files
compiled from /root/reference by oracle/Makefile; run in the container that has the reference tree, commit the output):

  ref_svo_<case>.npz      tg_svo_create (graphics/tg_sparse_voxel_octree.c:466-542) on the synthetic scenes of tests/test_reference_pins.py
  ref_simplex_noise.npz   tgm_simplex_noise (math/tg_math.c:182-302) on the arguments the procedural fill uses + random points

The oracle (CPU suite) and the CUDA builder (GPU suite) are compared with these files, so the chain reference -> oracle ->
kernels also holds on a box where /root/reference does not exist.

    python tests/golden/make_reference_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tg_b200 import ctypes_defs as T  # noqa: E402
from tg_b200 import scenes  # noqa: E402

SVO_CASES = {
    "small_grid": lambda: scenes.small_grid(),
    "grid4_tall": lambda: scenes.small_grid(grid=4, dims=(3, 5, 2)),
    "config1_small": lambda: scenes.config1(k=3, width=64, height=36, dims=(6, 4, 6)),
    "dense_k1": lambda: scenes.config1(k=1, width=64, height=36, dims=(4, 4, 4)),
    "sparse_k5": lambda: scenes.small_grid(grid=3, k=5),
}


def noise_points():
    rng = np.random.default_rng(5)
    pts = [rng.uniform(-300, 300, 3).astype(np.float32) for _ in range(4000)]
    for obj in (0, 1, 7):
        for _ in range(1500):
            x, y, z = rng.integers(0, 128), rng.integers(0, 128), rng.integers(0, 128)
            xf = np.float32(x) + np.float32(obj) * np.float32(1024.0)
            pts.append(np.array([xf * np.float32(0.008), 0.0, np.float32(z) * np.float32(0.008)], dtype=np.float32))
            pts.append(np.array([xf * np.float32(0.2), 0.0, np.float32(z) * np.float32(0.2)], dtype=np.float32))
            pts.append(np.array([np.float32(0.06) * xf, np.float32(0.06) * np.float32(y), np.float32(0.06) * np.float32(z)], dtype=np.float32))
    return np.stack(pts)


if __name__ == "__main__":
    from oracle import oracle as O
    from tests.test_reference_pins import svo_arrays_of
    R = O.ref()
    assert R is not None, "oracle/_ref/libtg_ref.so is missing: run `make -C oracle` where /root/reference exists"
    for name, make in SVO_CASES.items():
        view = O.SceneView.from_scene(make(), with_lut=False)
        svo = T.tg_svo()
        scene = O.ref_scene(view)
        R.tg_svo_create(T.v3(-512, -512, -512), T.v3(512, 512, 512), C.byref(scene), C.byref(svo))
        nodes, leaf, vox = svo_arrays_of(svo)
        R.tg_svo_destroy(C.byref(svo))
        nz = np.nonzero(vox)[0].astype(np.uint32)
        np.savez_compressed(os.path.join(HERE, f"ref_svo_{name}.npz"), nodes=nodes, leaf=leaf, n_voxel_words=np.uint32(vox.size), voxels_nonzero_idx=nz, voxels_nonzero=vox[nz])
        print(name, nodes.shape, leaf.shape, vox.size, int(nz.size))
    pts = noise_points()
    vals = np.array([R.tgm_simplex_noise(*map(np.float32, p)) for p in pts], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "ref_simplex_noise.npz"), points=pts, values=vals)
    print("noise", pts.shape)

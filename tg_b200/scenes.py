"""Seeded synthetic scenes for BASELINE.json's configs (the L5 "application" of SURVEY.md section 1).

The reference can only fill objects procedurally (tgvk_raytracer.c:871-978) and ships one demo scene
(tg_application.c:49-98); BASELINE's configs ask for random solid bits, so the inputs are defined here,
once, for the CUDA path, the oracle and the bench alike:

  * RNG = the reference's xorshift32 (math/tg_math.c:328-338) and murmur3 finaliser (util.inc:47-56).
  * Every cluster owns a stream: state0 = hash_u32(object_seed ^ hash_u32(rel_cluster)) | 1. A mask word is
    the AND of `k` successive draws (k=1 -> density 1/2, k=3 -> density 1/8); 16 words per cluster.
    (Per-cluster instead of per-object streams so that generation vectorises over 10^6..10^8 clusters.)
  * Material (LUT) index of voxel v of cluster c: hash_u32(hash_u32(object_seed + 0x9E3779B9) ^ (c*512+v)) & 7.
  * Object rotation = the reference's rule: idx*7 degrees about +Y, object 0 -> 15 degrees (tgvk_raytracer.c:829-832).
  * LUT = the reference's ramp (tg_application.c:89-98).
"""
from dataclasses import dataclass, field

import numpy as np

PI_F32 = np.float32(3.14159274)


def deg2rad(deg):
    """TG_DEG2RAD, math/tg_math.h:27, in float32."""
    return np.float32(deg) * ((PI_F32 * np.float32(2.0)) / np.float32(360.0))


def hash_u32(v):
    v = np.asarray(v, dtype=np.uint32).copy()
    v ^= v >> np.uint32(16)
    v *= np.uint32(0x85EBCA6B)
    v ^= v >> np.uint32(13)
    v *= np.uint32(0xC2B2AE35)
    v ^= v >> np.uint32(16)
    return v


def xorshift32_step(state):
    state ^= state << np.uint32(13)
    state ^= state >> np.uint32(17)
    state ^= state << np.uint32(5)
    return state


def random_solid_bits(object_seed, n_clusters, k):
    """[n_clusters, 16] uint32 masks, see module docstring."""
    with np.errstate(over="ignore"):
        rel = np.arange(n_clusters, dtype=np.uint32)
        state = hash_u32(np.uint32(object_seed) ^ hash_u32(rel)) | np.uint32(1)
        out = np.empty((n_clusters, 16), dtype=np.uint32)
        for wi in range(16):
            word = np.full(n_clusters, 0xFFFFFFFF, dtype=np.uint32)
            for _ in range(k):
                state = xorshift32_step(state)
                word &= state
            out[:, wi] = word
    return out


def random_lut_indices(object_seed, n_clusters, n_entries=8):
    """[n_clusters, 512] uint8 material indices in [0, n_entries)."""
    with np.errstate(over="ignore"):
        base = hash_u32(np.uint32((int(object_seed) + 0x9E3779B9) & 0xFFFFFFFF))
        idx = np.arange(n_clusters * 512, dtype=np.uint32)
        h = hash_u32(base ^ idx)
    return (h % np.uint32(n_entries)).astype(np.uint8).reshape(n_clusters, 512)


def reference_lut_ramp(n=256):
    """tg_application.c:89-98 -> list of (r, g, b) float32 triples."""
    lut = [(1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)]
    for i in range(3, 256):
        b = np.float32(i - 3) / np.float32(252.0)
        r = np.float32(0.5) - np.float32(0.5) * b
        lut.append((float(r), 0.0, float(b)))
    return lut[:n]


def reference_object_angle(object_idx):
    """tgvk_raytracer.c:830-831."""
    return float(deg2rad(15.0)) if object_idx == 0 else float(deg2rad(np.float32(object_idx * 7)))


@dataclass
class ObjectSpec:
    center: tuple
    extent: tuple                 # voxels, multiples of 8
    angle: float
    axis: tuple = (0.0, 1.0, 0.0)
    lut_idx: int = 0
    bits: np.ndarray = None       # [n_clusters, 16] uint32, pointer order (x fastest, then y, then z)
    lut_indices: np.ndarray = None  # [n_clusters, 512] uint8 or None (reference rule)
    seed: int = None              # bits is None and seed given: random_solid_bits(seed, n, k), generated on the device
    k: int = 3

    @property
    def dims(self):
        return (self.extent[0] // 8, self.extent[1] // 8, self.extent[2] // 8)

    @property
    def n_clusters(self):
        d = self.dims
        return d[0] * d[1] * d[2]


@dataclass
class CameraSpec:
    position: tuple
    pitch: float
    yaw: float
    roll: float
    fov_y_deg: float = 70.0
    aspect: float = 16.0 / 9.0
    near: float = 0.1
    far: float = 1000.0


@dataclass
class SceneSpec:
    name: str
    width: int
    height: int
    camera: CameraSpec
    objects: list = field(default_factory=list)
    lut: list = field(default_factory=lambda: reference_lut_ramp(8))
    n_luts: int = 1

    @property
    def n_clusters(self):
        return sum(o.n_clusters for o in self.objects)


def default_lut_indices(dims):
    """The reference's material rule (8*rel_x + vx) % 256 (tgvk_raytracer.c:947-978) -> [n_clusters, 512] uint8."""
    nx, ny, nz = dims
    rel_x = np.tile(np.arange(nx, dtype=np.uint32), ny * nz)
    vx = np.tile(np.arange(8, dtype=np.uint32), 64)
    return ((8 * rel_x[:, None] + vx[None, :]) % 256).astype(np.uint8)


def config1(k=3, width=1280, height=720, dims=(16, 16, 16), with_materials=True):
    """BASELINE configs[0]: one object of 16^3 clusters, random bits, 8-entry LUT, 1280x720."""
    n = dims[0] * dims[1] * dims[2]
    obj = ObjectSpec(center=(0.0, 0.0, 0.0), extent=(dims[0] * 8, dims[1] * 8, dims[2] * 8), angle=reference_object_angle(0),
                     bits=random_solid_bits(1, n, k), lut_indices=random_lut_indices(1, n) if with_materials else None)
    cam = CameraSpec(position=(0.0, 0.0, 160.0), pitch=0.0, yaw=0.0, roll=0.0, aspect=width / height)
    return SceneSpec(name=f"config1_k{k}", width=width, height=height, camera=cam, objects=[obj])


def deal_by_distance(grid_x, grid_z, n_ranks, pitch_units=192.0, eye_xz=(0.0, 0.0)):
    """Which rank holds lattice cell idx = j * grid_x + i: the cells in order of their distance from the camera's ground position,
    dealt out boustrophedon (0 1 .. n-1 n-1 .. 1 0 ...). Every rank gets the same number of objects and an equal share of the near
    (large on screen) and of the far ones -- what an engine that balances its GPUs by screen coverage would do. With the plain
    interleave (i + j) % n the few nearest objects, which cover a quarter of the frame each, made one rank's K1 twice as long as the
    mean (profiles/r03g_*)."""
    idx = np.arange(grid_x * grid_z)
    cx = (idx % grid_x - (grid_x - 1) / 2.0) * pitch_units - eye_xz[0]
    cz = (idx // grid_x - (grid_z - 1) / 2.0) * pitch_units - eye_xz[1]
    order = np.argsort(cx * cx + cz * cz, kind="stable")
    pos = np.arange(len(idx)) % n_ranks
    snake = np.where((np.arange(len(idx)) // n_ranks) % 2 == 0, pos, n_ranks - 1 - pos)
    owner = np.empty(len(idx), dtype=np.int64)
    owner[order] = snake
    return owner


def grid_scene(name, grid_x, grid_z, width, height, k=3, pitch_units=192.0, dims=(16, 8, 16), first_object=0, n_objects=None,
               with_bits=True, owner=None):
    """Objects on a grid_x x grid_z XZ lattice centred on the origin, y = 0 (configs 2-5). with_bits=False leaves the masks to
    the device generator (tg_raytracer_create_object_synthetic with the same seed idx + 1 and k): same bits, no host array.
    owner=(rank, n_ranks) keeps the objects of one rank of a multi-GPU partition (deal_by_distance; seed, angle and position still
    follow the global lattice index)."""
    total = grid_x * grid_z
    if n_objects is None:
        n_objects = total - first_object
    n = dims[0] * dims[1] * dims[2]
    objs = []
    owner_of = deal_by_distance(grid_x, grid_z, owner[1], pitch_units) if owner is not None else None
    for idx in range(first_object, first_object + n_objects):
        i, j = idx % grid_x, idx // grid_x
        if owner is not None and owner_of[idx] != owner[0]:
            continue
        cx = (i - (grid_x - 1) / 2.0) * pitch_units
        cz = (j - (grid_z - 1) / 2.0) * pitch_units
        objs.append(ObjectSpec(center=(cx, 0.0, cz), extent=(dims[0] * 8, dims[1] * 8, dims[2] * 8), angle=reference_object_angle(idx),
                               bits=random_solid_bits(idx + 1, n, k) if with_bits else None, lut_indices=None, seed=idx + 1, k=k))
    cam = CameraSpec(position=(0.0, 200.0, 0.0), pitch=float(deg2rad(-30.0)), yaw=0.0, roll=0.0, aspect=width / height)
    return SceneSpec(name=name, width=width, height=height, camera=cam, objects=objs)


def config2(width=3840, height=2160, grid=32, k=3):
    """BASELINE configs[1]: 1,024 rotated/translated objects (2^21 clusters, ~10^9 voxels), 4K."""
    return grid_scene(f"config2_{grid}x{grid}", grid, grid, width, height, k=k)


def config5_shard(rank, n_ranks, width=3840, height=2160, k=3):
    """BASELINE configs[4]: rank's share of a 384 x (32 n_ranks) lattice of 16x8x16-cluster objects -- 12,288 objects
    = 25,165,824 clusters = 1.29e10 voxels per rank; 8 ranks = 98,304 objects, 201,326,592 clusters, 1.03e11 voxels (SURVEY.md
    section 8d). The lattice cells are dealt out over the ranks in order of their distance from the camera (deal_by_distance), so every
    rank holds an equal part of whatever the camera sees; a rank's objects are a contiguous range of GLOBAL cluster pointers (rank-major pointer space). The masks are
    generated on the device (seed = global lattice index + 1), nothing of that size exists on the host."""
    return grid_scene(f"config5_x{n_ranks}", 384, 32 * n_ranks, width, height, k=k, with_bits=False, owner=(rank, n_ranks))


def small_grid(grid=3, width=320, height=180, k=3, dims=(4, 2, 4)):
    """Miniature of config 2 for CPU-sized parity tests."""
    s = grid_scene(f"small_grid{grid}", grid, grid, width, height, k=k, pitch_units=8.0 * dims[0] * 1.5, dims=dims)
    s.camera = CameraSpec(position=(0.0, 60.0, 40.0), pitch=float(deg2rad(-40.0)), yaw=0.0, roll=0.0, aspect=width / height)
    return s


def flat_arrays(scene, global_pointer_base=0, with_lut=True):
    """Fresh-scene flat arrays (cluster idx == cluster pointer, tgvk_raytracer.c:704-712): dict of numpy arrays
    objects[VOXEL_OBJECT_DTYPE], cluster_pointers, c2o, masks [n,16], lut_idx [n,512], color_lut [n_luts*256]."""
    from .ctypes_defs import VOXEL_OBJECT_DTYPE
    n_obj = len(scene.objects)
    objects = np.zeros(n_obj, dtype=VOXEL_OBJECT_DTYPE)
    first = 0
    masks, luts, c2o = [], [], []
    for oi, o in enumerate(scene.objects):
        objects[oi]["dims"] = o.dims
        objects[oi]["first_cluster_pointer"] = first
        objects[oi]["translation"] = o.center
        objects[oi]["angle_in_radians"] = o.angle
        objects[oi]["axis"] = o.axis
        masks.append(o.bits)
        if with_lut:
            luts.append(o.lut_indices if o.lut_indices is not None else default_lut_indices(o.dims))
        c2o.append(np.full(o.n_clusters, oi, dtype=np.uint32))
        first += o.n_clusters
    from .ctypes_defs import TG_U32_MAX  # noqa: F401
    color_lut = np.zeros(256 * scene.n_luts, dtype=np.uint32)
    for i, (r, g, b) in enumerate(scene.lut):
        color_lut[i] = pack_color(r, g, b)
    return dict(objects=objects, object_lut_idx=np.array([o.lut_idx for o in scene.objects], dtype=np.uint32),
                cluster_pointers=np.arange(first, dtype=np.uint32), c2o=np.concatenate(c2o),
                masks=np.ascontiguousarray(np.concatenate(masks)), lut_idx=np.ascontiguousarray(np.concatenate(luts)) if with_lut else None,
                color_lut=color_lut, global_pointer_base=global_pointer_base)


def pack_color(r, g, b):
    """tgvk_raytracer.c:1130-1134."""
    f = np.float32
    return (int(f(r) * f(255.0)) << 24) | (int(f(g) * f(255.0)) << 16) | (int(f(b) * f(255.0)) << 8) | 255


def reference_app_objects():
    """The ten tg_raytracer_create_object(center, extent) calls of the reference's sample scene (tg_application.c:65-87)."""
    calls = [((0.0, -64.0, 0.0), (128, 32, 128))]
    for depth_idx in range(3):
        offset_z = -float(depth_idx) * 128.0
        x = -256.0
        calls.append(((x, -16.0, -64.0 + offset_z), (32, 32, 32)))
        calls.append(((x, 9.0, -96.0 + offset_z), (32, 32, 32)))
        calls.append(((x - 6.0, 100.0, -70.0 + offset_z), (32, 32, 32)))
    return calls


def reference_app_scene(width, height, solid_bits_of, camera=None):
    """The reference application's scene (tg_application.c:49-98) as a SceneSpec: camera, ten procedural objects, the 256-entry
    LUT ramp. `solid_bits_of(object_idx, dims) -> [n_clusters, 16] uint32` supplies the terrain bits (tests pass the oracle's
    restatement of tgvk_raytracer.c:871-943); materials follow the reference rule (8 x + vx) % 256."""
    objs = []
    for idx, (center, extent) in enumerate(reference_app_objects()):
        dims = (extent[0] // 8, extent[1] // 8, extent[2] // 8)
        objs.append(ObjectSpec(center=center, extent=extent, angle=reference_object_angle(idx), bits=solid_bits_of(idx, dims), lut_indices=None))
    cam = camera or CameraSpec(position=(65.1368790, -30.7384720, 73.0285263), pitch=-0.173136666, yaw=0.710419059, roll=0.0, aspect=width / height)
    return SceneSpec(name="reference_app", width=width, height=height, camera=cam, objects=objs, lut=reference_lut_ramp(256))

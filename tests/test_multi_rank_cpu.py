"""CPU, world_size 2 over gloo: the host-side plan of the multi-GPU path (tg_b200/sharding.py). Each rank evaluates the
oracle on ITS shard with global pointers, the buffers are merged with a 64-bit min, and the result must equal the
oracle on the whole scene -- the same contract the NCCL path has on GPUs (tests/test_multi_gpu.py)."""
import os
import subprocess
import sys

import numpy as np

from tg_b200 import scenes, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["TG_ROOT"])
from tg_b200 import scenes, sharding
from oracle import oracle as O

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
s = scenes.small_grid(grid=3, width=160, height=90)
sub, base, first = sharding.shard_scene(s, world, rank)
rays = O.camera_rays(O.camera_from_spec(s.camera))
vis, _ = O.visibility(O.SceneView.from_scene(sub, base, with_lut=False), rays, s.width, s.height, O.VIS_SCREEN_RECT)
merged = sharding.allreduce_min_u64(vis.ravel()).reshape(vis.shape)
whole, _ = O.visibility(O.SceneView.from_scene(s, 0, with_lut=False), rays, s.width, s.height, O.VIS_SCREEN_RECT)
assert np.array_equal(merged, whole), f"rank {rank}: {int((merged != whole).sum())} words differ after the min merge"
hit = vis != np.uint64(0xFFFFFFFFFFFFFFFF)
ptr = (vis[hit] >> np.uint64(9)) & np.uint64(0x7FFFFFFF)
assert ptr.size and ptr.min() >= base and ptr.max() < base + sub.n_clusters, "a shard must only write its own global pointers"
# the peer-memory merge (tgb_peer.cu): every rank takes min + winner over all ranks' LOCAL buffers for ITS tile only
import torch
gathered = [torch.zeros(vis.size, dtype=torch.int64) for _ in range(world)]
dist.all_gather(gathered, torch.from_numpy(vis.ravel().view(np.int64).copy()))
all_vis = np.stack([g.numpy().view(np.uint64).reshape(vis.shape) for g in gathered])
tile, who = sharding.merge_tile_from_peers(all_vis, rank, s.height, s.width)
my_rows = sharding.tile_physical_rows(s.height, world, rank)
my_rows = my_rows[my_rows >= 0]
assert np.array_equal(tile, whole[my_rows]), "the tile merged from the peers' buffers must equal the all-reduced rows"
# ... and with K1's per-tile hit flags: a peer's tile without a hit is not read at all, the merged words stay the same
flags = [sharding.tile_hit_flags(v) for v in all_vis]
assert all(f.shape == ((s.height + 15) // 16, (s.width + 15) // 16) for f in flags) and flags[rank].any() and not all(f.all() for f in flags)
tile_f, skipped = sharding.merge_tile_from_flagged_peers(all_vis, flags, rank, s.height, s.width)
assert np.array_equal(tile_f, tile) and skipped > 0.0, "skipping the peers' empty tiles must not change the merged tile"
bases = [sharding.shard_scene(s, world, r)[1] for r in range(world)] + [s.n_clusters]
tile_ptr = ((tile >> np.uint64(9)) & np.uint64(0x7FFFFFFF)).astype(np.int64)
owner = np.searchsorted(np.asarray(bases[1:]), tile_ptr, side="right")
assert np.array_equal(who[who >= 0], owner[who >= 0]), "the rank holding the minimum is the owner of the winning pointer (it supplies the material)"
rows = np.zeros(s.height, dtype=np.int64); rows[my_rows] = 1
import torch
t = torch.from_numpy(rows); dist.all_reduce(t)
assert (t.numpy() == 1).all(), "screen tiles must partition the rows"
dist.barrier()
if rank == 0:
    print("MULTI_RANK_CPU_OK")
dist.destroy_process_group()
'''


def test_object_ranges_partition_the_scene():
    for n_obj in (1, 7, 9, 1024):
        for n in (1, 2, 4, 8):
            ranges = [sharding.object_range(n_obj, n, r) for r in range(n)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n_obj
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    s = scenes.small_grid()
    bases = [sharding.shard_scene(s, 4, r)[1] for r in range(4)]
    assert bases == sorted(bases) and bases[0] == 0
    assert sum(sharding.shard_scene(s, 4, r)[0].n_clusters for r in range(4)) == s.n_clusters
    # screen tiles: 16-row bands dealt out to the ranks; tiles are whole bands, equal for every rank, and partition the rows
    assert sharding.tile_row_count(2160, 8) == 272 and [sharding.tile_rows(2160, 8, r) for r in (0, 7)] == [(0, 272), (1904, 2176)]
    assert sharding.tile_rows(2160, 1, 0) == (0, 2160) and np.array_equal(sharding.tile_physical_rows(7, 1, 0), np.arange(7))
    assert sharding.tile_physical_rows(2160, 8, 3)[:18].tolist() == list(range(48, 64)) + [176, 177]
    for h, n in ((2160, 8), (2160, 4), (177, 2), (10, 4), (33, 8)):
        owned = np.concatenate([sharding.tile_physical_rows(h, n, r) for r in range(n)])
        assert len(owned) == n * sharding.tile_row_count(h, n) and sorted(owned[owned >= 0].tolist()) == list(range(h))


def test_min_merge_of_sharded_oracle_frames_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, TG_ROOT=ROOT, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29613", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "MULTI_RANK_CPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]

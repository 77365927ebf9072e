"""Worker of tests/test_multi_gpu.py (one process per GPU, launched by torch.distributed.run): the sharded frame --
local K1, ncclAllReduce(u64, min), collective SVO build, owner-resolved materials, tile-split GI -- must equal, bit for
bit, the frame one GPU renders from the union of the shards (which the single-GPU tests pin against the oracle)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from tg_b200 import scenes, sharding
    from tg_b200.raytracer import comm_unique_id, from_scene

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = {
        "grid5": lambda: scenes.small_grid(grid=5, width=333, height=177, dims=(3, 5, 2)),   # ragged resolution: uneven last tile
        "grid4": lambda: scenes.small_grid(grid=4, width=320, height=180),
    }
    for name, make in cases.items():
        s = make()
        for i, o in enumerate(s.objects):
            o.lut_idx = i % 2
        s.n_luts = 2
        sub, base, first = sharding.shard_scene(s, world, rank)
        cap = max(sharding.object_range(len(s.objects), world, r)[1] - sharding.object_range(len(s.objects), world, r)[0] for r in range(world))
        results = {}
        # the exchanges: 0 = one kernel over peer memory (CUDA IPC / NVLink), 1 = NCCL all-reduce + reduce-scatter,
        # 2 = peer memory requested but one rank cannot map it: every rank must agree on the NCCL path
        for kind in (0, 1, 2):
            if kind == 2:
                os.environ["TGB200_FAIL_P2P_ON_RANK"] = str(world - 1)
            rt = from_scene(sub, device=local, max_n_objects=cap, max_n_clusters=max(sub.n_clusters, 1))
            for i in range(8):
                rt.color_lut_set(i, 0.1 * i, 1.0 - 0.1 * i, 0.5, lut_idx=1)
            rt.set_shard(rank, world, base)
            ids = [comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            rt.comm_init(ids[0], rank, world)
            rt.set_merge_kind(0 if kind == 2 else kind)
            rt.set_gi(True, 5)
            for _ in range(3):       # three frames: the peer-memory path alternates between its two buffer pairs
                rt.clear()
                rt.render()          # K1 -> merge -> collective K2 -> materials -> K3 on this rank's rows
            rt.synchronize()
            dist.barrier()
            y0, y1 = rt.tile_rows()
            assert (y0, y1) == sharding.tile_rows(s.height, world, rank)
            assert np.array_equal(rt.tile_physical_rows(), sharding.tile_physical_rows(s.height, world, rank))
            vis = rt.read_visibility()
            svo, nodes, leaf, vox = rt.svo_download()
            rt.svo_free(svo)
            rt.gather_radiance()
            rt.synchronize()
            rad = rt.read_radiance()
            t = rt.timings()
            assert t["merge_ms"] > 0
            assert (t["merge_kernel_ms"] > 0) == (kind == 0), "kind 0 must run k_merge_tile, the others the NCCL collectives"
            # frame sink on a sharded frame: this rank's tile rows (16-row bands, tile order) land in host memory, HDR and presented
            rows = rt.tile_physical_rows()
            import torch
            hdr = torch.zeros(len(rows) * s.width * 4, dtype=torch.float32).pin_memory().numpy().reshape(len(rows), s.width, 4)
            rt.set_frame_sink(hdr, 3)
            rt.clear(); rt.render(); rt.wait_frame(rt.frame_ticket())
            assert np.array_equal(hdr[rows >= 0], rad[rows[rows >= 0]]), f"{name} rank {rank} kind {kind}: frame sink rows differ from the gathered frame"
            bgra = torch.zeros(len(rows) * s.width, dtype=torch.int32).pin_memory().numpy().view(np.uint32).reshape(len(rows), s.width)
            rt.set_frame_sink(bgra, 2)
            rt.clear(); rt.render(); rt.wait_frame(rt.frame_ticket())
            rt.set_frame_sink(None)
            rt.gather_radiance(); rt.synchronize()
            assert np.array_equal(bgra[rows >= 0], rt.read_present()[rows[rows >= 0]]), f"{name} rank {rank} kind {kind}: presented sink rows differ"
            dist.barrier()
            rt.comm_destroy()
            rt.destroy()
            results[kind] = (vis, nodes, leaf, vox, rad, t)
            os.environ.pop("TGB200_FAIL_P2P_ON_RANK", None)
        for other in (1, 2):
            assert all(np.array_equal(a, b) for a, b in zip(results[0][:5], results[other][:5])), f"{name} rank {rank}: the peer-memory merge and exchange {other} disagree"
        vis, nodes, leaf, vox, rad, t = results[0]

        # reference: the whole scene on this GPU alone
        ref = from_scene(s, device=local)
        for i in range(8):
            ref.color_lut_set(i, 0.1 * i, 1.0 - 0.1 * i, 0.5, lut_idx=1)
        ref.set_gi(True, 5)
        ref.clear(); ref.render(); ref.synchronize()
        want_vis, want_rad = ref.read_visibility(), ref.read_radiance()
        rsvo, rnodes, rleaf, rvox = ref.svo_download()
        ref.svo_free(rsvo)
        ref.destroy()
        assert np.array_equal(vis, want_vis), f"{name} rank {rank}: merged visibility differs in {int((vis != want_vis).sum())} words"
        assert np.array_equal(nodes, rnodes) and np.array_equal(vox, rvox), f"{name} rank {rank}: sharded SVO nodes / voxels differ"
        assert np.array_equal(leaf, rleaf), f"{name} rank {rank}: sharded SVO leaf records differ"
        bad = np.argwhere((rad != want_rad).any(axis=-1))
        assert bad.size == 0, (f"{name} rank {rank}: radiance differs in {len(bad)} pixels; rows {np.unique(bad[:, 0])[:12].tolist()}.. first "
                               f"{[(int(y), int(x), rad[y, x].tolist(), want_rad[y, x].tolist()) for y, x in bad[:3]]}")
        dist.barrier()
        if rank == 0:
            print(f"MGPU_OK {name} world={world} tile_rows={y1 - y0} merge_ms peer={results[0][5]['merge_ms']:.3f} nccl={results[1][5]['merge_ms']:.3f}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

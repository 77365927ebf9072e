"""Host-side plan of the multi-GPU path: which objects a rank owns, the global cluster-pointer
base of its shard, the screen tile it shades, and the 64-bit min merge expressed for a CPU (gloo) process group.

One process per GPU. Clusters are sharded BY OBJECT (an object's clusters are one contiguous pointer range,
tgvk_raytracer.c:826-827), ranks own contiguous object ranges, so global pointers = local pointer + base. The per-GPU
visibility buffers are merged with an element-wise 64-bit min (ncclAllReduce(ncclUint64, ncclMin) on GPUs); GI rays are
split by screen tile: the frame is cut into 16-row bands and rank r shades the bands b with b mod N == r (an equal share of
the hit pixels for every rank, whatever the view)."""
import numpy as np


def object_range(n_objects, n_ranks, rank):
    """Contiguous, balanced: the first n_objects % n_ranks ranks own one object more."""
    q, r = divmod(n_objects, n_ranks)
    first = rank * q + min(rank, r)
    return first, first + q + (1 if rank < r else 0)


def shard_scene(scene, n_ranks, rank):
    """(SceneSpec holding this rank's objects, global pointer base, global object base)."""
    from .scenes import SceneSpec
    first, last = object_range(len(scene.objects), n_ranks, rank)
    base = sum(o.n_clusters for o in scene.objects[:first])
    sub = SceneSpec(name=f"{scene.name}_shard{rank}of{n_ranks}", width=scene.width, height=scene.height, camera=scene.camera,
                    objects=scene.objects[first:last], lut=scene.lut, n_luts=scene.n_luts)
    return sub, base, first


BAND_ROWS = 16  # tg_b200/csrc/tgb_rows.h: the frame is cut into 16-row bands, band b belongs to rank b mod n_ranks


def tile_row_count(height, n_ranks):
    """Rows of a rank's tile, padding included: whole bands, the same for every rank."""
    n_bands = (height + BAND_ROWS - 1) // BAND_ROWS
    return ((n_bands + n_ranks - 1) // n_ranks) * BAND_ROWS


def tile_rows(height, n_ranks, rank):
    """A rank's rows in tile ("virtual") order, [first, one_past_last): the whole frame in frame order on one GPU."""
    if n_ranks == 1:
        return 0, height
    t = tile_row_count(height, n_ranks)
    return rank * t, (rank + 1) * t


def tile_physical_rows(height, n_ranks, rank):
    """Frame row of every row of the rank's tile, in tile order; -1 = padding (a partial last band, a band the rank does not have)."""
    if n_ranks == 1:
        return np.arange(height, dtype=np.int64)
    t = tile_row_count(height, n_ranks)
    i = np.arange(t, dtype=np.int64)
    rows = ((i // BAND_ROWS) * n_ranks + rank) * BAND_ROWS + i % BAND_ROWS
    return np.where(rows < height, rows, -1)


def allreduce_min_u64(vis, group=None):
    """Element-wise unsigned 64-bit min over a torch.distributed group that lacks uint64 reductions (gloo): flipping the
    sign bit maps unsigned order onto signed order. `vis` is a numpy uint64 array; returns the merged array."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy((vis ^ np.uint64(1 << 63)).view(np.int64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return t.numpy().view(np.uint64) ^ np.uint64(1 << 63)


def merge_tile_from_peers(all_vis, rank, height, width):
    """The peer-memory merge of tg_b200/csrc/tgb_peer.cu stated on the host: `all_vis` = every rank's LOCAL buffer
    [n_ranks, h, w] (what k_merge_tile reads over NVLink); returns (merged words of this rank's tile rows, winner rank per pixel
    or -1). min is associative, so the tile equals the same rows of the all-reduced frame."""
    n_ranks = all_vis.shape[0]
    rows = tile_physical_rows(height, n_ranks, rank)
    tile = all_vis[:, rows[rows >= 0]]
    who = np.argmin(tile, axis=0)          # first rank holding the minimum, like the kernel's strict `<`
    best = np.take_along_axis(tile, who[None], axis=0)[0]
    who = np.where(best == np.uint64(0xFFFFFFFFFFFFFFFF), -1, who)
    return best, who


def tile_hit_flags(vis):
    """K1's epilogue on the peer-memory path, stated on the host: one flag per 16x16 pixel tile of a rank's LOCAL buffer, set when
    the rank has any hit there ([ceil(h/16), ceil(w/16)] bool). k_merge_tile does not read a peer's tile whose flag is clear."""
    h, w = vis.shape
    th, tw = (h + BAND_ROWS - 1) // BAND_ROWS, (w + 15) // 16
    padded = np.full((th * BAND_ROWS, tw * 16), np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
    padded[:h, :w] = vis
    return (padded.reshape(th, BAND_ROWS, tw, 16) != np.uint64(0xFFFFFFFFFFFFFFFF)).any(axis=(1, 3))


def merge_tile_from_flagged_peers(all_vis, all_flags, rank, height, width):
    """merge_tile_from_peers reading only the tiles a peer flagged (what crosses NVLink): (merged words of the rank's tile rows,
    fraction of the (rank, pixel) reads that were skipped). Skipped tiles hold only the clear value, so the minimum is unchanged."""
    n_ranks = all_vis.shape[0]
    rows = tile_physical_rows(height, n_ranks, rank)
    rows = rows[rows >= 0]
    cols = np.arange(width)
    flagged = np.stack([f[np.ix_(rows // BAND_ROWS, cols // 16)] for f in all_flags])     # [n_ranks, rows, width]
    words = np.where(flagged, all_vis[:, rows], np.uint64(0xFFFFFFFFFFFFFFFF))
    return words.min(axis=0), 1.0 - float(flagged.mean())

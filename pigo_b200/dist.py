"""Frame sharding + the single detection gather of the multi-GPU path.

Frames are independent (core/pigo.go:226-256 holds no cross-window state): rank g of G takes frames
[g*ceil(N/G), (g+1)*ceil(N/G)); cascades are replicated; the only exchange is one gather of the per-frame counts and
one gather of the padded detection slices to rank 0, which restores frame order so the result equals the 1-GPU one.
torch.distributed is plumbing (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(nframes: int, rank: int, world: int) -> Tuple[int, int]:
    per = -(-nframes // world) if world > 0 else nframes
    lo = min(nframes, rank * per)
    return lo, min(nframes, lo + per)


def gather_detections(dets: torch.Tensor, counts: torch.Tensor, dst: int = 0,
                      group=None) -> Optional[Tuple[List[torch.Tensor], List[torch.Tensor]]]:
    """dets: [n_local, cap, 4] int32 (row, col, scale, q-bits); counts: [n_local] int32, both on this rank's device.
    Every rank must hold the same n_local (pad the last shard).  Returns, on `dst`, the per-rank lists in rank (= frame)
    order; None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return [dets], [counts]
    if rank == dst:
        cl = [torch.empty_like(counts) for _ in range(world)]
        dl = [torch.empty_like(dets) for _ in range(world)]
    else:
        cl = dl = None
    dist.gather(counts, cl, dst=dst, group=group)
    dist.gather(dets, dl, dst=dst, group=group)
    return (dl, cl) if rank == dst else None


def merge_gathered(dets_list: List[torch.Tensor], counts_list: List[torch.Tensor], nframes: int):
    """Concatenates the rank slices in frame order and drops the padding frames of the last shard."""
    d = torch.cat(dets_list, dim=0)[:nframes]
    c = torch.cat(counts_list, dim=0)[:nframes]
    return d, c


def gather_pipeline(faces: torch.Tensor, nfaces: torch.Tensor, points: torch.Tensor, dst: int = 0, group=None):
    """The gather of the full pipeline (BASELINE configs[4]): faces [n_local, cap, 4] int32, nfaces [n_local] int32 and
    points [n_local, cap, 2 + ncalls, 4] int32 (eyes + landmark calls) of every rank's frame shard, to `dst`, in rank
    (= frame) order.  Returns (faces_list, nfaces_list, points_list) on dst, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return [faces], [nfaces], [points]
    outs = []
    for t in (nfaces, faces, points):
        lst = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, lst, dst=dst, group=group)
        outs.append(lst)
    return (outs[1], outs[0], outs[2]) if rank == dst else None


def merge_pipeline(faces_list, nfaces_list, points_list, nframes: int):
    """Frame-ordered concatenation of the gathered shards, padding frames of the last shard dropped."""
    return (torch.cat(faces_list, dim=0)[:nframes], torch.cat(nfaces_list, dim=0)[:nframes], torch.cat(points_list, dim=0)[:nframes])

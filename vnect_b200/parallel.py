"""Multi-GPU plumbing: one process per GPU, video streams sharded round-robin, results gathered once per step.

The hot path has no data-path collective: frames of different streams share nothing and the
14.6 M weights are replicated.  The only exchange is the gather of the per-stream results (21 x 5 numbers per frame),
done with ``torch.distributed`` (NCCL over NVLink on the GPUs, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist

JOINTS = 21


def owned_streams(n_streams, rank, world):
    """Stream i lives on rank i % world (its OneEuroFilter state never moves)."""
    return list(range(rank, n_streams, world))


def pack_results(j2, j3):
    """[n,21,2] float64 + [n,21,3] float32 -> [n,21,5] float64 (lossless)."""
    n = j2.shape[0]
    out = np.empty((n, JOINTS, 5), np.float64)
    out[:, :, :2] = j2
    out[:, :, 2:] = j3
    return out


def unpack_results(packed):
    return packed[:, :, :2].copy(), packed[:, :, 2:].astype(np.float32)


def gather_results(packed_local, n_streams, device=None, group=None):
    """All-gather the per-rank results into global stream order.  Every rank must own ceil/floor(n_streams/world)
    streams as given by owned_streams(); ranks with one stream fewer pad with a dummy row."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-n_streams // world)
    buf = torch.zeros((per, JOINTS, 5), dtype=torch.float64, device=device)
    loc = torch.as_tensor(packed_local, dtype=torch.float64, device=device)
    buf[:loc.shape[0]] = loc
    out = torch.empty((world, per, JOINTS, 5), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out.view(world * per, JOINTS, 5), buf, group=group)
    res = np.empty((n_streams, JOINTS, 5), np.float64)
    out_h = out.cpu().numpy()
    for r in range(world):
        ids = owned_streams(n_streams, r, world)
        res[ids] = out_h[r, :len(ids)]
    return res

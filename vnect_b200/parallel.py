"""Multi-GPU plumbing: one process per GPU, video streams sharded round-robin, results gathered once per step.

The hot path has no data-path collective: frames of different streams share nothing and the
14.6 M weights are replicated.  The only exchange is the gather of the per-stream results (21 x 5 numbers per frame),
done with ``torch.distributed`` (NCCL over NVLink on the GPUs, gloo in the CPU tests).

Layout of the exchange: every rank owns ``per = ceil(n_streams / world)`` slots of float64 ``[per, 21, 5]``
(row, col, x, y, z -- written in that form by the post-process kernel, see vnect_set_packed_results); slot k of rank r is
global stream ``r + k * world``.  ``all_gather_packed`` moves the slots, ``unshard_index`` gives the permutation back to
stream order.  Both bench.py and gather_results() go through them.
"""
import numpy as np
import torch
import torch.distributed as dist

JOINTS = 21


def owned_streams(n_streams, rank, world):
    """Stream i lives on rank i % world (its OneEuroFilter state never moves)."""
    return list(range(rank, n_streams, world))


def slots_per_rank(n_streams, world):
    return -(-n_streams // world)


def pack_results(j2, j3):
    """[n,21,2] float64 + [n,21,3] float32 -> [n,21,5] float64 (lossless)."""
    n = j2.shape[0]
    out = np.empty((n, JOINTS, 5), np.float64)
    out[:, :, :2] = j2
    out[:, :, 2:] = j3
    return out


def unpack_results(packed):
    return packed[:, :, :2].copy(), packed[:, :, 2:].astype(np.float32)


def all_gather_packed(local, out=None, group=None, async_op=False):
    """local: tensor float64 [per, 21, 5] on this rank's device (padded to `per` slots).  Returns (out, work) with out
    float64 [world * per, 21, 5] in RANK-MAJOR order; work is None unless async_op."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world * local.shape[0], JOINTS, 5), dtype=torch.float64, device=local.device)
    work = dist.all_gather_into_tensor(out, local, group=group, async_op=async_op)
    return out, work


def unshard_index(n_streams, world):
    """index[s] = row of stream s in the rank-major gathered tensor."""
    per = slots_per_rank(n_streams, world)
    s = np.arange(n_streams)
    return (s % world) * per + s // world


def gather_results(packed_local, n_streams, device=None, group=None):
    """All-gather the per-rank results into global stream order (host array float64 [n_streams, 21, 5]).  Every rank
    must own the streams given by owned_streams(); ranks with one stream fewer pad with a dummy slot."""
    world = dist.get_world_size(group)
    per = slots_per_rank(n_streams, world)
    buf = torch.zeros((per, JOINTS, 5), dtype=torch.float64, device=device)
    loc = torch.as_tensor(packed_local, dtype=torch.float64, device=device)
    buf[:loc.shape[0]] = loc
    out, _ = all_gather_packed(buf, group=group)
    return out.cpu().numpy()[unshard_index(n_streams, world)]

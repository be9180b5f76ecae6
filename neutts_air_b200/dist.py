"""Multi-GPU plumbing: one process per GPU, full weight replica each, utterances sharded across
ranks, ONE all-gather of the finished waveforms (plus an 8-byte all_reduce when the
caller cannot bound the waveform length).  The path has no other exchange step.  Works on NCCL (GPU box) and gloo (CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as td


def world() -> tuple:
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


def shard_plan(n_items: int, lengths, world_size: int) -> list:
    """Static assignment of utterance indices to ranks: longest-first round robin over prompt lengths
    (better balance than contiguous blocks for mixed lengths); deterministic on every rank."""
    order = sorted(range(n_items), key=lambda i: (-int(lengths[i]), i))
    plan = [[] for _ in range(world_size)]
    for j, i in enumerate(order):
        r = j % world_size if (j // world_size) % 2 == 0 else world_size - 1 - (j % world_size)
        plan[r].append(i)
    return [sorted(p) for p in plan]


def shard_indices(n_items: int, lengths) -> list:
    rank, ws = world()
    return shard_plan(n_items, lengths, ws)[rank]


def all_gather_waveforms(local_wavs, local_idx, n_items: int, device="cuda", lengths=None, t_max: int | None = None) -> list:
    """local_wavs[j] is the waveform (numpy array or tensor, host or device) of global item local_idx[j].
    Returns all n_items waveforms, on every rank, through ONE all_gather of a padded [slots, 2 + T_max] float32
    buffer: column 0 carries the global index + 1, column 1 the sample count, so indices and lengths ride in the same
    collective.  ``slots`` = ceil(n_items / world) is static (the shard plan deals the items round robin).  When the
    caller knows an upper bound on the sample count (``t_max``: bench.py, fixed-length jobs) that all-gather is the
    only collective; otherwise one 8-byte all_reduce(MAX) agrees on T_max first."""
    rank, ws = world()
    if ws == 1:
        out = [None] * n_items
        for i, w in zip(local_idx, local_wavs):
            out[i] = _to_numpy(w)
        return out
    dev = torch.device(device)
    slots = (n_items + ws - 1) // ws
    if len(local_wavs) > slots:
        raise ValueError("more local items than the round-robin shard plan allows")
    tmax = t_max
    if tmax is None:
        meta = torch.tensor([max([len(w) for w in local_wavs], default=0)], dtype=torch.int64, device=dev)
        td.all_reduce(meta, op=td.ReduceOp.MAX)
        tmax = int(meta[0])
    buf = torch.zeros(slots, 2 + tmax, dtype=torch.float32, device=dev)
    for j, (i, w) in enumerate(zip(local_idx, local_wavs)):
        wt = w if isinstance(w, torch.Tensor) else torch.as_tensor(np.asarray(w, dtype=np.float32))
        wt = wt.flatten()
        if len(wt) > tmax:
            raise ValueError(f"waveform of {len(wt)} samples exceeds t_max {tmax}")
        buf[j, 0] = float(i + 1)             # global index + 1 (0 = empty slot); exact in fp32 below 2^24
        buf[j, 1] = float(len(wt))
        buf[j, 2: 2 + len(wt)] = wt.to(dev, torch.float32)
    gathered = torch.empty(ws * slots, 2 + tmax, dtype=torch.float32, device=dev)
    td.all_gather_into_tensor(gathered, buf)
    g = gathered.cpu().numpy()
    out = [None] * n_items
    for row in g:
        if row[0] > 0:
            out[int(row[0]) - 1] = row[2: 2 + int(row[1])].copy()
    return out


def _to_numpy(w):
    if isinstance(w, torch.Tensor):
        return w.detach().flatten().float().cpu().numpy()
    return np.asarray(w, dtype=np.float32)

"""Frame sharding across the GPUs of one box.

Frames are independent units: rank r of W processes the contiguous range shard_range(n, r, W)
with sampler streams keyed by the GLOBAL frame index, so results are identical for any W.
There is no data-path collective; gather_rows is only used to assemble small per-frame
results (poses, errors) for reporting.
"""
import torch
import torch.distributed as dist


def shard_range(n_frames, rank, world):
    """Contiguous split; the first n %% world ranks get one extra frame."""
    q, r = divmod(n_frames, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def gather_rows(local, n_total, world):
    """all_gather of per-frame rows ([n_local, k] tensors of unequal n_local) in frame order."""
    if world == 1:
        return local
    k = local.shape[1]
    q, r = divmod(n_total, world)
    pad = q + (1 if r else 0)
    buf = torch.zeros((pad, k), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    outs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    rows = []
    for rank in range(world):
        lo, hi = shard_range(n_total, rank, world)
        rows.append(outs[rank][: hi - lo])
    return torch.cat(rows, 0)

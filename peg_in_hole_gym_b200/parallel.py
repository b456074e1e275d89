"""Multi-GPU plumbing: one process per GPU, envs sharded with NO data-path collective.

`mp_num` of the reference (one pybullet client process each, envs/base_env_mp.py:27-35) maps to ranks: every rank
owns a contiguous block of clients, hence a contiguous range of global env ids.  RNG is keyed by global env id, so
results do not depend on the number of ranks.  The only collective is the OPTIONAL gather of the packed
(obs5, reward, done) rows into a rank-0 learner buffer (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import os

import torch


def world():
    """(rank, world_size) from torch.distributed if initialised, else from the torchrun env, else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_clients(mp_num, rank, world_size):
    """Contiguous block of reference clients owned by `rank`: [lo, hi).  Remainder goes to the low ranks."""
    if world_size <= 1:
        return 0, mp_num
    base, rem = divmod(mp_num, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_rows(obs, reward, done):
    """[n,5] obs + [n] reward + [n] done -> one [n,7] float32 block (28 B/env)."""
    return torch.cat([obs, reward.unsqueeze(1), done.to(obs.dtype).unsqueeze(1)], dim=1).contiguous()


def unpack_rows(rows):
    return rows[:, :5], rows[:, 5], rows[:, 6] != 0


def gather_rows_to_rank0(rows, counts=None):
    """Gather every rank's [n_r,7] block on rank 0, in rank order.  Returns the [sum n_r,7] tensor on rank 0 and
    None elsewhere.  Uses all_gather on padded blocks when shard sizes differ (NCCL has no ragged gather)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rows
    ws, rank = dist.get_world_size(), dist.get_rank()
    if counts is None:
        c = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
        cl = [torch.zeros_like(c) for _ in range(ws)]
        dist.all_gather(cl, c)
        counts = [int(v.item()) for v in cl]
    mx = max(counts)
    if rows.shape[0] < mx:
        rows = torch.cat([rows, rows.new_zeros(mx - rows.shape[0], rows.shape[1])], 0)
    if all(v == mx for v in counts):
        out = rows.new_empty(ws * mx, rows.shape[1]) if rank == 0 else None
        gl = list(out.chunk(ws, 0)) if rank == 0 else None
        dist.gather(rows, gl, dst=0)
        return out
    buf = [torch.empty_like(rows) for _ in range(ws)]
    dist.all_gather(buf, rows)
    if rank != 0:
        return None
    return torch.cat([b[:n] for b, n in zip(buf, counts)], 0)

"""World-axis sharding for multi-GPU runs.

Worlds are independent units — the reference runs one OS process per Monte-Carlo world
(libs/monte-carlo/src/lib.rs:2083) — so ranks own contiguous world ranges and the data path
needs no collective.  `torch.distributed` is used for the end-of-run gather and counters
only; on GPUs that is NCCL over NVLink, in the CPU tests gloo.
"""

from __future__ import annotations

from typing import List, Tuple



def shard_worlds(n_worlds: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced world range [w0, w1) of `rank` (sizes differ by at most 1)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, rem = divmod(int(n_worlds), world_size)
    w0 = rank * base + min(rank, rem)
    return w0, w0 + base + (1 if rank < rem else 0)


def shard_sizes(n_worlds: int, world_size: int) -> List[int]:
    return [shard_worlds(n_worlds, r, world_size)[1] - shard_worlds(n_worlds, r, world_size)[0] for r in range(world_size)]


def gather_worlds(local, n_worlds: int, group=None):
    """End-of-run gather of a per-world tensor [w_local, ...] from every rank into the global
    world order [n_worlds, ...] (all ranks get the result).  Ragged shards are padded to the
    largest shard for the fixed-size all_gather and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    ws = dist.get_world_size(group)
    sizes = shard_sizes(n_worlds, ws)
    mx = max(sizes)
    if min(sizes) == mx and local.shape[0] == mx:
        # equal shards: one flat all-gather straight into the result (no per-rank staging tensors)
        out = torch.empty((ws * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)


def total_entity_steps(local_entity_steps: int, group=None) -> int:
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([int(local_entity_steps)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, group=group)
    return int(t.item())

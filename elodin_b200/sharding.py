"""World-axis sharding for multi-GPU runs.

Worlds are independent units — the reference runs one OS process per Monte-Carlo world
(libs/monte-carlo/src/lib.rs:2083) — so ranks own contiguous world ranges and the data path
needs no collective.  `torch.distributed` is used for the end-of-run gather and counters
only; on GPUs that is NCCL over NVLink, in the CPU tests gloo.
"""

from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np


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


class Comm:
    """One rank of the library's own NCCL communicator (include/b200_sixdof.h `b200_comm`): the C-ABI route of
    the end-of-run trajectory gather, usable by a host without torch.  Rank 0 creates the 128-byte id with
    `Comm.unique_id()` and hands it to every rank over whatever channel the host has (here: a
    `torch.distributed` broadcast; the reference's Monte-Carlo driver would write it into context.json)."""

    def __init__(self, unique_id: bytes, n_ranks: int, rank: int, device: int = -1):
        from . import _lib

        self._L = _lib.lib()
        buf = (C.c_uint8 * _lib.COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        h = C.c_void_p()
        _lib.check(self._L.b200_comm_create(buf, int(n_ranks), int(rank), int(device), C.byref(h)))
        self._h = h
        self.n_ranks, self.rank = int(n_ranks), int(rank)

    @staticmethod
    def available() -> bool:
        from . import _lib

        return bool(_lib.lib().b200_comm_available())

    @staticmethod
    def unique_id() -> bytes:
        from . import _lib

        buf = (C.c_uint8 * _lib.COMM_ID_BYTES)()
        _lib.check(_lib.lib().b200_comm_unique_id(buf, _lib.COMM_ID_BYTES))
        return bytes(buf)

    @property
    def last_ms(self) -> float:
        return float(self._L.b200_comm_last_ms(self._h))

    def trajectory_allgather(self, exec_, worlds_per_rank: Sequence[int], out: Optional[np.ndarray] = None,
                             out_ptr: Optional[int] = None):
        """World-sharded all-gather of `exec_`'s device trajectory ring: [sum(worlds), samples, n_entities, width]
        on every rank, in rank order.  `out_ptr` may be a device pointer (e.g. a torch CUDA tensor's data_ptr())."""
        from . import _lib

        wpr = (C.c_uint64 * self.n_ranks)(*[int(w) for w in worlds_per_rank])
        nbytes = int(self._L.b200_sixdof_trajectory_gather_bytes(exec_._h, wpr, self.n_ranks))
        if out_ptr is None:
            if out is None:
                out = np.empty((int(sum(worlds_per_rank)), exec_.trajectory_len(), exec_.n_entities, exec_.trajectory_width()))
            if out.nbytes != nbytes:
                raise _lib.B200ValueError(_lib.ERR_VALUE_SIZE_MISMATCH, f"gathered trajectory is {nbytes} bytes, buffer has {out.nbytes}")
            out_ptr = out.ctypes.data
        _lib.check(self._L.b200_sixdof_trajectory_allgather(exec_._h, self._h, wpr, C.c_void_p(out_ptr), nbytes))
        return out

    def step_row_sharded(self, exec_, n_ticks: int) -> None:
        """ONE world, source rows split over the ranks (b200_sixdof_step_row_sharded): every rank holds the whole
        world and integrates rows [rank * N / R, (rank + 1) * N / R); the rows' new position / velocity planes are
        all-gathered over NCCL after every tick."""
        from . import _lib

        _lib.check(self._L.b200_sixdof_step_row_sharded(exec_._h, self._h, int(n_ticks)))

    def peer_attach(self, exec_) -> None:
        """Map every rank's peer window (b200_comm_peer_attach; collective): step_row_sharded then exchanges the rows
        with direct NVLink stores and counter releases instead of a collective per tick.  Raises B200Error
        (ERR_UNSUPPORTED) on every rank when the processes cannot share memory over CUDA IPC."""
        from . import _lib

        _lib.check(self._L.b200_comm_peer_attach(self._h, exec_._h))

    @property
    def peer_attached(self) -> bool:
        return bool(self._L.b200_comm_peer_attached(self._h))

    def peer_detach(self) -> None:
        """Collective: unmap the peers' windows, then free the own one.  Call before closing the attached executor."""
        self._L.b200_comm_peer_detach(self._h)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.b200_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RowShardedWorld:
    """bench.py helper for SURVEY §8e's second case ("report both"): one n-body world on N GPUs as replicas vs row
    shards.  Not a product class: the product entry is Comm.step_row_sharded / b200_sixdof_step_row_sharded."""

    @staticmethod
    def _comm(torch, dist, world_size, rank, local):
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(Comm.unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        return Comm(bytes(uid.cpu().tolist()), world_size, rank, local)

    @staticmethod
    def _timed(torch, el, stream, local, comm, barrier, max_over_ranks, world, grav, warm, ticks, peer, math="fast"):
        """One route of the row-sharded world: `peer` = NVLink stores into the peers' windows, else ncclAllGather per tick."""
        from . import _lib
        from .executor import WORLD_POS

        p, v, I = world
        N = p.shape[1]
        ex = el.B200Exec(N, 1, 3600.0, None, [grav(N)], "rk4", math, device=local)
        ex.set_stream(stream.cuda_stream)
        ex.set_state(p, v, I)
        why = None
        if peer:
            try:
                comm.peer_attach(ex)
            except _lib.B200Error as e:  # every rank gets the same answer (the attach is agreed across ranks)
                why = str(e)[:200]
        if peer and why:
            ex.close()
            return None, None, why
        ev = lambda: torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            comm.step_row_sharded(ex, warm)
            barrier()
            a, b = ev(), ev()
            a.record(stream)
            comm.step_row_sharded(ex, ticks)
            b.record(stream)
            barrier()
        ms = max_over_ranks(a.elapsed_time(b))
        pos = ex.download(WORLD_POS)
        if peer:
            comm.peer_detach()
        ex.close()
        return ms, pos, None

    @staticmethod
    def _both(torch, dist, el, stream, local, rank, world_size, barrier, max_over_ranks, world, grav, ref_pos, warm, ticks):
        comm = RowShardedWorld._comm(torch, dist, world_size, rank, local)
        N = world[0].shape[1]
        scale = float(np.max(np.abs(ref_pos[..., 4:])))
        out = {"ranks": world_size, "rows_per_gpu": N // world_size, "unit": "entity-steps/s"}
        for name, peer in (("nccl", False), ("peer", True)):
            ms, pos, why = RowShardedWorld._timed(torch, el, stream, local, comm, barrier, max_over_ranks, world, grav, warm, ticks, peer)
            if why:
                out[name] = {"unavailable": why}
                continue
            out[name] = {"us_per_tick": ms * 1e3 / ticks, "value": N * ticks / (ms * 1e-3),
                         "max_rel_diff_vs_replica": float(np.max(np.abs(pos[..., 4:] - ref_pos[..., 4:])) / scale)}
        out["nccl"]["exchange"] = "6 planes x N/R f64 per rank, in-place ncclAllGather per plane, one NCCL group per tick (inside libb200_sixdof.so)"
        out["peer"]["exchange"] = ("rows stored straight into every rank's CUDA-IPC window over NVLink + one counter release per peer; "
                                   "the next tick's gravity waits on the counters (no collective in the tick loop)")
        best = min((o for o in (out["nccl"], out["peer"]) if "us_per_tick" in o), key=lambda o: o["us_per_tick"])
        out["us_per_tick"], out["value"] = best["us_per_tick"], best["value"]
        out["max_rel_diff_vs_replica"] = best["max_rel_diff_vs_replica"]
        comm.close()
        return out

    @staticmethod
    def bench(args, torch, dist, el, stream, local, rank, world_size, barrier, max_over_ranks, world, grav, ref_pos):
        """ref_pos: the replica's positions after 420 ticks."""
        return RowShardedWorld._both(torch, dist, el, stream, local, rank, world_size, barrier, max_over_ranks, world, grav, ref_pos, 20, 400)

    @staticmethod
    def bench_large(torch, dist, el, stream, local, rank, world_size, barrier, max_over_ranks, world, grav, ref_pos, warm, ticks):
        return RowShardedWorld._both(torch, dist, el, stream, local, rank, world_size, barrier, max_over_ranks, world, grav, ref_pos, warm, ticks)

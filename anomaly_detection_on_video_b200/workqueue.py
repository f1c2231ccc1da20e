"""Work queue that shards independent items (videos, 188-clip chunks, feature files) over one process
per GPU.  The hot path has NO collective: every clip-crop is an independent forward and every video
an independent output file (reference extract_features.py:85-89,104-156), so ranks only need to agree
on who takes which item.

Dynamic mode: a host-side atomic counter in the ``torch.distributed`` TCP store (``store.add``) -- each
rank claims the next unclaimed index, so a rank stuck on a long video simply claims fewer items.
Static mode (no store): item ``i`` belongs to rank ``i % world_size`` (callers order items longest
first, which makes round-robin a reasonable LPT schedule).  Both are deterministic in *what* is
computed: outputs do not depend on which rank produced them.

Shutdown: the TCP store lives inside one of the worker ranks (rank 0 when ``from_env`` creates it).  A rank that runs out
of items must not take the store down while others still claim, so every rank calls ``close()`` when it is done: it
checks in on a ``done`` key and the store's owner waits there until all ranks have checked in.
"""
from __future__ import annotations

import os
from typing import Iterator, Optional

import torch.distributed as dist


class WorkQueue:
    def __init__(self, rank: int = 0, world_size: int = 1, local_rank: int = 0, store=None, dynamic: bool = True,
                 owns_store: bool = False) -> None:
        self.rank, self.world_size, self.local_rank = rank, world_size, local_rank
        self.store = store if dynamic else None
        self.owns_store = bool(owns_store and self.store is not None)  # this process hosts the store's server
        self._epoch = 0
        self._closed = False

    @classmethod
    def from_env(cls, dynamic: bool = True) -> Optional["WorkQueue"]:
        """Build from torchrun's RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*; None when single-process."""
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world <= 1:
            return None
        rank = int(os.environ["RANK"])
        local_rank = int(os.environ.get("LOCAL_RANK", rank))
        store = None
        owns = False
        if dynamic:
            if dist.is_available() and dist.is_initialized():
                store = _default_store()
            else:
                host = os.environ.get("MASTER_ADDR", "127.0.0.1")
                port = int(os.environ.get("MASTER_PORT", "29500")) + 17  # beside, not on, the rendezvous port
                store = dist.TCPStore(host, port, world, is_master=(rank == 0), wait_for_workers=False)
                owns = rank == 0
        return cls(rank, world, local_rank, store, dynamic, owns_store=owns)

    @classmethod
    def from_process_group(cls, dynamic: bool = True) -> "WorkQueue":
        rank, world = dist.get_rank(), dist.get_world_size()
        return cls(rank, world, int(os.environ.get("LOCAL_RANK", rank)), _default_store() if dynamic else None, dynamic)

    def claim(self, n_items: int, tag: str = "") -> Iterator[int]:
        """Yield the indices in [0, n_items) this rank must process; every index is yielded on exactly
        one rank.  All ranks must call ``claim`` the same number of times with the same ``n_items``."""
        self._epoch += 1
        if self.world_size == 1:
            yield from range(n_items)
            return
        if self.store is None:
            yield from range(self.rank, n_items, self.world_size)
            return
        key = f"vad_wq/{self._epoch}/{tag}"
        while True:
            idx = self.store.add(key, 1) - 1  # atomic fetch-and-add on the host; no GPU traffic
            if idx >= n_items:
                return
            yield idx

    def barrier(self) -> None:
        """Host-side barrier through the store (used between 'extract' and 'segment' phases)."""
        if self.world_size == 1:
            return
        if self.store is None:
            if dist.is_available() and dist.is_initialized():
                dist.barrier()
            return
        self._epoch += 1
        key = f"vad_wq/barrier/{self._epoch}"
        self.store.add(key, 1)
        import time

        while int(self.store.add(key, 0)) < self.world_size:
            time.sleep(0.005)


    def close(self, timeout_s: float = 600.0) -> None:
        """Check out.  Every rank calls this after its last ``claim`` / ``barrier``; the rank hosting the store returns only
        when all ranks have checked out (or after ``timeout_s``), so nobody ever talks to a dead server.  Idempotent."""
        if self._closed or self.world_size == 1 or self.store is None:
            self._closed = True
            return
        self._closed = True
        import time

        key = "vad_wq/done"
        try:
            self.store.add(key, 1)
            if self.owns_store:
                t0 = time.monotonic()
                while int(self.store.add(key, 0)) < self.world_size and time.monotonic() - t0 < timeout_s:
                    time.sleep(0.005)
        except Exception:  # a peer that already lost the server has nothing left to coordinate
            pass

    def __enter__(self) -> "WorkQueue":
        return self

    def __exit__(self, *exc) -> None:
        self.close()


def _default_store():
    # the store behind the default process group (private accessor; stable across torch 2.x)
    from torch.distributed import distributed_c10d as c10d

    return c10d._get_default_store()


__all__ = ["WorkQueue"]

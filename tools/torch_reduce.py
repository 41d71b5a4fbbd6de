"""torch.distributed twins of the multi-GPU reduction -- NOT part of the product.

The product (fuzzysearch_b200/) reduces the per-shard match lists inside libfuzzb200.so (peer-memory push +
device merge, NCCL only as bootstrap / staged fallback) and has no torch dependency.  These helpers do the
same reduction through torch.distributed (NCCL or gloo) and are kept for the CPU (gloo, world_size > 1)
tests of the host-side merge logic and as a comparison arm (`bench.py --reduce torch`).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

__all__ = ["gather_rows", "GroupReducer", "gather_and_merge_groups", "broadcast_bytes"]


def broadcast_bytes(payload, group=None):
    """rank 0's bytes -> every rank, over torch.distributed (used to ship an NCCL id in a process group that
    already exists; the product's own rendezvous is fuzzysearch_b200.sharding.rendezvous_bytes)."""
    import torch.distributed as dist
    box = [payload if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


def gather_rows(rows, group=None, device=None):
    """All-gather a variable number of int64 rows ([n_i, C] per rank) -> [sum n_i, C] on every rank,
    rank-major.  Uses torch.distributed when initialised (NCCL: tensors on `device`; gloo: CPU),
    otherwise returns `rows` (single process)."""
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    try:
        import torch
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return rows
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rows
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cpu")
    if backend == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ncols = rows.shape[1]
    count = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    width = max(max(counts), 1)
    padded = torch.zeros((width, ncols), dtype=torch.int64, device=dev)
    if rows.shape[0]:
        padded[:rows.shape[0]] = torch.from_numpy(rows).to(dev)
    out = [torch.zeros((width, ncols), dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    parts = [out[r][:counts[r]].cpu().numpy() for r in range(world)]
    return np.concatenate(parts, axis=0) if parts else rows


class GroupReducer(object):
    """The multi-GPU reduction of one search, with its buffers allocated once.

    Every rank contributes its locally consolidated groups (rows (start, end, dist, hull_start,
    hull_end)); ONE fixed-size all-gather (slot = count row + padded rows; grown and retried if a
    rank overflows it) and a linear merge of the almost-ordered shard lists (fzb_merge_groups) yield
    the global final list on every rank."""

    def __init__(self, group=None, device=None, cap=4096):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group)
        self.nccl = dist.get_backend(group) == "nccl"
        self.dev = torch.device("cpu")
        if self.nccl:
            self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._alloc(cap)

    def _alloc(self, cap):
        torch = self.torch
        self.cap = cap
        self.send_host = torch.zeros((cap + 1, 5), dtype=torch.int64, pin_memory=self.nccl)
        self.recv_host = torch.zeros((self.world, cap + 1, 5), dtype=torch.int64, pin_memory=self.nccl)
        self.send_np = self.send_host.numpy()
        self.recv_np = self.recv_host.numpy()
        if self.nccl:
            self.send_dev = torch.empty_like(self.send_host, device=self.dev)
            self.recv_dev = torch.empty_like(self.recv_host, device=self.dev)

    def reduce(self, result=None, rows=None, as_arrays=False):
        """`result`: a _native.Result (rows are pulled straight into the pinned send buffer), or
        `rows`: an int64 [n,5] array.  Returns the global final list."""
        from fuzzysearch_b200 import _native
        while True:
            if result is not None:
                n = result.group_rows(out=self.send_np[1:])
            else:
                rows = np.asarray(rows, dtype=np.int64).reshape(-1, 5)
                n = rows.shape[0]
                self.send_np[1:1 + min(n, self.cap)] = rows[:self.cap]
            self.send_np[0, 0] = n
            if self.nccl:
                self.send_dev.copy_(self.send_host, non_blocking=True)
                self.dist.all_gather_into_tensor(self.recv_dev.view(-1), self.send_dev.view(-1), group=self.group)
                self.recv_host.copy_(self.recv_dev, non_blocking=True)
                self.torch.cuda.current_stream().synchronize()
            else:
                parts = [self.torch.empty_like(self.send_host) for _ in range(self.world)]
                self.dist.all_gather(parts, self.send_host, group=self.group)
                for r in range(self.world):
                    self.recv_host[r].copy_(parts[r])
            counts = self.recv_np[:, 0, 0]
            top = int(counts.max())
            if top <= self.cap:
                parts = [self.recv_np[r, 1:1 + int(counts[r])] for r in range(self.world)]
                return _native.merge_groups(np.concatenate(parts, axis=0), as_arrays=as_arrays)
            cap = self.cap
            while cap < top:
                cap *= 2
            self._alloc(cap)


_REDUCERS = {}


def gather_and_merge_groups(group_rows=None, group=None, device=None, result=None, as_arrays=False):
    """Global final list from every rank's local groups (see GroupReducer); single-process when
    torch.distributed is not initialised."""
    from fuzzysearch_b200 import _native
    try:
        import torch.distributed as dist
        ready = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    except ImportError:  # pragma: no cover
        ready = False
    if not ready:
        rows = result.group_rows() if result is not None else group_rows
        return _native.merge_groups(rows, as_arrays=as_arrays)
    key = (id(group), str(device))
    red = _REDUCERS.get(key)
    if red is None:
        red = _REDUCERS[key] = GroupReducer(group, device)
    return red.reduce(result=result, rows=group_rows, as_arrays=as_arrays)



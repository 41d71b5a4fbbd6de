"""Multi-GPU layout of one long sequence: contiguous shards with a halo, one process per GPU.

SURVEY.md section 8e.  Every raw match is a pure function of a bounded window around its *anchor*
(n-gram hit index / Hamming start / LP start), so rank g OWNS the anchors in ``[own_lo, own_hi)``,
loads ``H[own_lo - halo : own_hi + halo)`` with ``halo = len(pattern) + max_l_dist`` and emits only
matches it owns.  The reference's clipping rules are evaluated at the global ends only, so the union
of the per-rank raw streams IS the single-device raw stream -- no haystack byte ever crosses NVLink.
The only exchange is the (tiny) match list over ``torch.distributed`` (NCCL on GPUs, gloo in the CPU
tests).  Each rank consolidates its own raw stream on the device and contributes one row per GROUP of
overlapping matches -- the winner and the group's hull; groups can chain across seams, but the winner
of a union of groups is the better of their winners and hulls overlap iff members do, so the global
consolidation (common.py:185-189) is ONE fixed-size all-gather plus a linear merge
(``gather_and_merge_groups``).  ``gather_rows``/``merge_raw_streams`` gather the raw streams instead
(parity tests).

The reference's CPU analogue of this layout is the chunk + carry-over tail loop of
``_search_binary_file`` (fuzzysearch/__init__.py:129-171).
"""
import numpy as np

__all__ = ["shard_bounds", "gather_rows", "merge_raw_streams", "gather_and_merge_groups", "GroupReducer", "init_shard_comm"]

ALIGN = 16  # shard buffers start on 16-byte boundaries of the global sequence (uint4 loads)


def shard_bounds(global_len, world_size, rank, halo):
    """-> (buf_lo, buf_hi, own_lo, own_hi) for `rank` of `world_size`.

    Owned ranges partition [0, global_len); seams are multiples of ALIGN; buffers extend the owned
    range by `halo` on both sides, clipped to the sequence and aligned down at the start."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")

    def seam(i):
        if i >= world_size:
            return global_len
        return (global_len * i // world_size) // ALIGN * ALIGN

    own_lo, own_hi = seam(rank), seam(rank + 1)
    buf_lo = max(0, own_lo - halo) // ALIGN * ALIGN
    buf_hi = min(global_len, own_hi + halo)
    return buf_lo, buf_hi, own_lo, own_hi


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


def init_shard_comm(haystack, group=None):
    """Give a shard handle its own NCCL communicator so that searches flagged F_GLOBAL all-gather the
    per-shard groups inside the library, on the search's stream (no Python or torch in the timed path).
    The unique id travels over torch.distributed (plumbing only)."""
    import torch.distributed as dist
    from . import _native
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [_native.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    haystack.comm_init(box[0], rank, world)


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
        from . import _native
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
    from . import _native
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


def merge_raw_streams(rows, ngram_route=True):
    """Gathered rows (start, end, dist, ngram, idx) -> (raw rows in reference order, final triples).

    Raw order: n-gram major then hit index for the n-gram routes (the reference's generation order),
    (start, end, dist) otherwise.  Final = consolidate_overlapping_matches via the native sweep."""
    from . import _native
    rows = np.asarray(rows, dtype=np.int64).reshape(-1, 5)
    if ngram_route:
        order = np.lexsort((rows[:, 4], rows[:, 3]))
    else:
        order = np.lexsort((rows[:, 2], rows[:, 1], rows[:, 0]))
    rows = rows[order]
    s, e, d = _native.consolidate(rows[:, 0], rows[:, 1], rows[:, 2].astype(np.int32))
    return rows, list(zip(s.tolist(), e.tolist(), d.tolist()))

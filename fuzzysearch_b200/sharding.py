"""Multi-GPU layout of one long sequence: contiguous shards with a halo, one process per GPU.

SURVEY.md section 8e.  Every raw match is a pure function of a bounded window around its *anchor*
(n-gram hit index / Hamming start / LP start), so rank g OWNS the anchors in ``[own_lo, own_hi)``,
loads ``H[own_lo - halo : own_hi + halo)`` with ``halo = len(pattern) + max_l_dist`` and emits only
matches it owns.  The reference's clipping rules are evaluated at the global ends only, so the union
of the per-rank raw streams IS the single-device raw stream -- no haystack byte ever crosses NVLink.
The only exchange is the (tiny) match list.  Each rank consolidates its own raw stream on the device and
contributes one row per GROUP of overlapping matches -- the winner and the group's hull; groups can chain
across seams, but the winner of a union of groups is the better of their winners and hulls overlap iff
members do, so the global consolidation (common.py:185-189) is one exchange plus a seam-local merge.  With
``F_GLOBAL`` the library does both on the device, over NVLink peer memory, behind the search kernels
(csrc/p2p_kernels.cuh); this module only holds the geometry and the torch-free bootstrap:
``init_shard_comm`` ships rank 0's NCCL id over a plain TCP socket on MASTER_ADDR (no torch in the
product), ``init_local_world`` binds several shards of ONE process.

The reference's CPU analogue of this layout is the chunk + carry-over tail loop of
``_search_binary_file`` (fuzzysearch/__init__.py:129-171).
"""
import numpy as np

import os
import socket
import struct
import time

import numpy as np

__all__ = ["shard_bounds", "merge_raw_streams", "rendezvous_bytes", "allgather_bytes", "init_shard_comm",
           "init_shard_world_ipc", "init_local_world", "search_all"]

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


_ROUND = [0]  # collectives issued by this process: every rank issues the same sequence, so the number identifies one


def _exchange(payload, rank, world_size, addr, port, timeout, gather):
    """One collective over TCP.  gather = False: rank 0's payload -> every rank.  gather = True: every rank's payload
    (equal lengths) -> the rank-major concatenation on every rank.  Rank 0 listens, the others connect (retrying until
    it is up); every connection starts with (magic, round, rank) so that a fast rank that is already in the NEXT
    collective cannot be mistaken for a participant of this one (it is turned away and retries)."""
    _ROUND[0] += 1
    rnd = _ROUND[0]
    if world_size == 1:
        return bytes(payload)
    addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
    if port is None:
        port = int(os.environ.get("FZB_RDV_PORT", 0)) or int(os.environ.get("MASTER_PORT", "29500")) + 17
    deadline = time.time() + timeout
    n = len(payload)

    def recv_exact(conn, count):
        data = b""
        while len(data) < count:
            chunk = conn.recv(count - len(data))
            if not chunk:
                raise ConnectionError("rendezvous closed early")
            data += chunk
        return data

    if rank == 0:
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        while True:
            try:
                srv.bind((addr, port))
                break
            except OSError:  # the previous collective's listener is still closing
                if time.time() > deadline:
                    raise
                time.sleep(0.02)
        srv.listen(4 * world_size)
        srv.settimeout(timeout)
        parts, conns = {0: bytes(payload)}, []
        try:
            while len(conns) < world_size - 1:
                conn, _ = srv.accept()
                conn.settimeout(timeout)
                try:
                    magic, r_rnd, r = struct.unpack("<4sII", recv_exact(conn, 12))
                    if magic != b"FZBR" or r_rnd != rnd or not (0 < r < world_size) or r in parts:
                        conn.close()  # somebody else's round (or noise): the sender retries
                        continue
                    parts[r] = recv_exact(conn, n) if gather else b""
                except (ConnectionError, socket.timeout, OSError, struct.error):
                    conn.close()
                    continue
                conns.append(conn)
            blob = b"".join(parts[r] for r in range(world_size)) if gather else bytes(payload)
            for conn in conns:
                conn.sendall(struct.pack("<I", len(blob)) + blob)
        finally:
            for conn in conns:
                conn.close()
            srv.close()
        return blob
    while True:
        try:
            with socket.create_connection((addr, port), timeout=5.0) as conn:
                conn.settimeout(max(1.0, deadline - time.time()))
                conn.sendall(struct.pack("<4sII", b"FZBR", rnd, rank) + (bytes(payload) if gather else b""))
                (m,) = struct.unpack("<I", recv_exact(conn, 4))
                return recv_exact(conn, m)
        except (ConnectionRefusedError, ConnectionError, socket.timeout, OSError):
            if time.time() > deadline:
                raise
            time.sleep(0.05)


def rendezvous_bytes(payload, rank, world_size, addr=None, port=None, timeout=120.0):
    """rank 0's `payload` (bytes) -> every rank, over a plain TCP socket on (addr, port).  Defaults: MASTER_ADDR and
    FZB_RDV_PORT, else MASTER_PORT + 17 (torchrun's own store owns MASTER_PORT itself).  No torch involved."""
    return _exchange(payload, rank, world_size, addr, port, timeout, gather=False)


def allgather_bytes(payload, rank, world_size, addr=None, port=None, timeout=120.0):
    """Every rank's `payload` (equal lengths) -> the rank-major concatenation on every rank (same transport)."""
    return _exchange(payload, rank, world_size, addr, port, timeout, gather=True)


def init_shard_world_ipc(haystack, rank=None, world_size=None, addr=None, port=None):
    """Collective: the peer-memory world WITHOUT NCCL -- the CUDA IPC handles of the receive areas are all-gathered
    over the TCP rendezvous.  Also works for several processes sharing one GPU.  Returns True if every rank could map
    every other rank's receive area (else the world is disabled on all ranks)."""
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world_size = int(os.environ.get("WORLD_SIZE", "1")) if world_size is None else world_size
    handles = allgather_bytes(haystack.p2p_export(rank, world_size), rank, world_size, addr, port)
    ok = True
    try:
        haystack.p2p_connect(handles)
    except Exception:  # noqa: BLE001 -- reported to the other ranks below
        ok = False
    flags = allgather_bytes(b"\x01" if ok else b"\x00", rank, world_size, addr, port)
    if flags != b"\x01" * world_size:
        haystack.p2p_disable()
        return False
    return True


def init_shard_comm(haystack, rank=None, world_size=None, addr=None, port=None):
    """Collective: give a shard handle its world -- an NCCL communicator (bootstrap + staged fallback) and,
    where CUDA IPC / peer access allow, the peer-memory receive areas that F_GLOBAL searches reduce over.
    rank / world_size default to RANK / WORLD_SIZE (torchrun's environment).  The NCCL id travels over
    ``rendezvous_bytes``."""
    from . import _native
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world_size = int(os.environ.get("WORLD_SIZE", "1")) if world_size is None else world_size
    uid = rendezvous_bytes(_native.nccl_unique_id() if rank == 0 else b"", rank, world_size, addr, port)
    haystack.comm_init(uid, rank, world_size)


def init_local_world(haystacks):
    """Bind the shard handles of ONE process into a world (rank = list index); F_GLOBAL searches must then
    run concurrently, one thread per handle (see ``search_all``)."""
    from . import _native
    _native.comm_init_local(list(haystacks))


def search_all(haystacks, call):
    """Run ``call(haystack)`` on every handle of a local world concurrently (one thread each: ctypes drops the
    GIL during the native search) and return the results in rank order; the first exception is re-raised."""
    import threading
    out = [None] * len(haystacks)
    errs = [None] * len(haystacks)

    def run(i):
        try:
            out[i] = call(haystacks[i])
        except BaseException as e:  # noqa: BLE001 -- re-raised below
            errs[i] = e

    threads = [threading.Thread(target=run, args=(i,)) for i in range(len(haystacks))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errs:
        if e is not None:
            raise e
    return out


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

"""World-size-2/3 CPU tests (gloo) of the multi-GPU host logic: shard geometry, the variable-length
gather of match rows and the merge.  The per-rank raw stream is emulated by filtering the oracle's
raw stream by anchor ownership (which is what each GPU rank emits, see test_gpu_oracle.py's
sharded test for the device side)."""
import os
import socket

import numpy as np
import pytest

import oracle
from corpus import ASCII, make_corpus
from fuzzysearch_b200 import _native
from fuzzysearch_b200.sharding import allgather_bytes, merge_raw_streams, rendezvous_bytes, shard_bounds
from parity import tup

import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from torch_reduce import gather_and_merge_groups, gather_rows  # noqa: E402  (torch twins, not in the product)


def test_shard_bounds_partition_and_halo():
    for n in (0, 1, 15, 16, 1000, (1 << 20) + 7):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                blo, bhi, lo, hi = shard_bounds(n, world, r, halo=22)
                assert lo == prev and lo <= hi <= n
                assert blo % 16 == 0 and blo <= max(0, lo - 22) and bhi >= min(n, hi + 22)
                assert lo % 16 == 0 or lo == n
                prev = hi
            assert prev == n


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rdv_worker(rank, world, port, q):
    payload = bytes(range(128)) if rank == 0 else b""
    got = rendezvous_bytes(payload, rank, world, "127.0.0.1", port, timeout=60)
    ok = got == bytes(range(128))
    # all-gather of equal-sized records (what carries the CUDA IPC handles of the NCCL-free world), twice in a row
    for rep in range(2):
        mine = bytes([rank * 16 + rep]) * 64
        allg = allgather_bytes(mine, rank, world, "127.0.0.1", port, timeout=60)
        ok = ok and allg == b"".join(bytes([r * 16 + rep]) * 64 for r in range(world))
    q.put((rank, ok))


@pytest.mark.parametrize("world", [2, 3])
def test_torch_free_rendezvous(world):
    """The product's own bootstrap (NCCL id from rank 0 to every rank over a TCP socket): no torch."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    # start the clients FIRST: they must retry until rank 0 listens
    procs = [ctx.Process(target=_rdv_worker, args=(r, world, port, q)) for r in reversed(range(world))]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(r, True) for r in range(world)]


def test_product_sharding_module_is_torch_free():
    import fuzzysearch_b200.sharding as sh
    src = open(sh.__file__).read()
    assert "import torch" not in src and "torch.distributed" not in src.replace("torch.distributed``", "")


def _worker(rank, world, port, n, m, k, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pat, hay, _ = make_corpus(77, n, ASCII, m, 48, 3)
        for r in range(1, world):  # plant across every seam
            seam = shard_bounds(n, world, r, m + k)[2]
            for j, delta in enumerate((-m, -m + 1, -4, -2, -1, 0, 1)):
                pos = seam + delta + 40 * (j - 3)
                hay[pos:pos + m] = np.frombuffer(pat, dtype=np.uint8)
        raw, ng, ix = oracle.levenshtein_ngrams_raw(pat, hay, k, with_anchor=True)
        blo, bhi, lo, hi = shard_bounds(n, world, rank, m + k)
        mine = (ix >= lo) & (ix < hi)
        rows = np.column_stack([raw[mine], ng[mine], ix[mine]]).astype(np.int64).reshape(-1, 5)
        if rank == world - 1:
            rows = rows[::-1]  # order inside a rank must not matter
        allrows = gather_rows(rows)
        merged, final = merge_raw_streams(allrows)
        ok = (tup(merged[:, :3]) == tup(raw)) and (final == tup(oracle.consolidate(raw)))
        # the production reduction: per-rank local groups -> one all-gather -> linear merge
        groups = _native.consolidate_groups(rows[:, 0], rows[:, 1], rows[:, 2].astype(np.int32))
        ok = ok and (gather_and_merge_groups(groups) == final)
        q.put((rank, bool(ok), int(rows.shape[0]), len(final)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_and_merge_equals_single_stream(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 1 << 18, 20, 2, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in results), results
    assert sum(cnt for _, _, cnt, _ in results) > 48
    assert len({nf for _, _, _, nf in results}) == 1

"""FZB_F_GLOBAL: the multi-rank reduction of per-shard groups over peer memory (k_push / k_merge), the staged
NCCL path, and the torch-free bootstrap.  The world of 2-4 shards on ONE GPU (fzb_comm_init_local, one thread
per shard) runs the very kernels a multi-GPU job runs, so this file needs no second GPU to be meaningful."""
import os
import socket

import numpy as np
import pytest

import oracle
from corpus import ASCII, DNA, make_corpus
from fuzzysearch_b200 import _native as F
from fuzzysearch_b200.sharding import init_local_world, init_shard_comm, search_all, shard_bounds
from conftest import needs_real_gpu
from parity import tup

pytestmark = pytest.mark.gpu


def test_global_flag_world_of_one(cuda_device):
    """A world of one rank: through NCCL bootstrap (peer-memory path with itself) and as a local world."""
    needs_real_gpu("NCCL bootstrap")
    pat, hay, _ = make_corpus(4, 1 << 20, ASCII, 20, 64, 3)
    hs = F.Haystack.from_host(hay)
    with pytest.raises(ValueError):
        hs.search_levenshtein(pat, 2, F.F_GLOBAL)      # no world yet
    for how in ("nccl", "local"):
        if how == "nccl":
            hs.comm_init(F.nccl_unique_id(), 0, 1)
        else:
            init_local_world([hs])
        assert hs.p2p_enabled()
        local = hs.search_levenshtein(pat, 2).triples(F.FINAL)
        assert len(local) > 30
        assert hs.search_levenshtein(pat, 2, F.F_GLOBAL).triples(F.FINAL) == local
        assert hs.search_levenshtein(pat, 2, F.F_GLOBAL | F.F_FORCE_LP).triples(F.FINAL) == \
            hs.search_levenshtein(pat, 2, F.F_FORCE_LP).triples(F.FINAL)
        ham = hs.search_hamming(pat, 3).triples(F.FINAL)
        assert hs.search_hamming(pat, 3, F.F_GLOBAL).triples(F.FINAL) == ham
        assert hs.search_exact(pat, F.F_GLOBAL).triples(F.FINAL) == hs.search_exact(pat).triples(F.FINAL)
        assert hs.search_generic(pat, 1, 2, 1, 2, F.F_GLOBAL).triples(F.FINAL) == \
            hs.search_generic(pat, 1, 2, 1, 2).triples(F.FINAL)
    hs.close()


def _seamy_corpus(seed, n, alphabet, m, k, world, plants):
    pat, hay, _ = make_corpus(seed, n, alphabet, m, plants, 3)
    for r in range(1, world):  # matches straddling every seam + overlapping copies whose groups chain over it
        seam = shard_bounds(n, world, r, m + k)[2]
        for j, delta in enumerate((-m, -m + 1, -4, -2, -1, 0, 1)):  # tests/test_find_near_matches_in_file.py:84-86
            pos = seam + delta + 96 * (j - 3)
            if 0 <= pos and pos + m <= n:
                hay[pos:pos + m] = np.frombuffer(pat, dtype=np.uint8)
        blob = pat + pat[m // 2:] + pat + pat[3:]
        if seam + 400 + len(blob) <= n:
            hay[seam + 400 - 30:seam + 400 - 30 + len(blob)] = np.frombuffer(blob, dtype=np.uint8)
    return pat, hay


def _make_world(hay, world, halo, device=0):
    n = len(hay)
    shards = []
    for r in range(world):
        blo, bhi, lo, hi = shard_bounds(n, world, r, halo)
        shards.append(F.Haystack.from_host(hay[blo:bhi], device=device, buf_lo=blo, global_len=n, own_lo=lo, own_hi=hi))
    init_local_world(shards)
    return shards


@pytest.mark.parametrize("world,n", [(2, (1 << 22) + 40), (3, 1 << 20), (4, 4096), (8, 1 << 21)])
def test_multi_rank_world_on_one_gpu(cuda_device, world, n):
    """The unskippable multi-rank parity test: `world` shards of one sequence on device 0, reduced by the same
    k_push / k_merge kernels that run across GPUs; every rank must hold the oracle's global list."""
    m, k = 20, 2
    pat, hay = _seamy_corpus(31 + world, n, ASCII, m, k, world, 256 if n > 100000 else 8)
    shards = _make_world(hay, world, m + k)
    try:
        assert all(h.p2p_enabled() for h in shards)
        exp = oracle.find_near_matches(pat, hay, max_l_dist=k)
        assert len(exp) > (30 if n > 100000 else 3)
        for rep in range(3):  # epochs / parities are reused
            got = search_all(shards, lambda h: h.search_levenshtein(pat, k, F.F_GLOBAL).triples(F.FINAL))
            for r in range(world):
                assert got[r] == exp, (rep, r)
        # the union of the shards' raw streams is the single-device raw stream
        raws = search_all(shards, lambda h: h.search_levenshtein(pat, k, F.F_GLOBAL | 0).arrays(F.RAW, anchors=True))
        rows = sorted((int(g), int(i), int(s), int(e), int(d)) for s_, e_, d_, g_, i_ in raws
                      for s, e, d, g, i in zip(s_.tolist(), e_.tolist(), d_.tolist(), g_.tolist(), i_.tolist()))
        raw_o, ng_o, ix_o = oracle.levenshtein_ngrams_raw(pat, hay, k, with_anchor=True)
        assert rows == sorted((int(g), int(i), int(s), int(e), int(d))
                              for (s, e, d), g, i in zip(raw_o.tolist(), ng_o.tolist(), ix_o.tolist()))
        # unconsolidated routes: the global list is the sorted union
        ham = search_all(shards, lambda h: h.search_hamming(pat, 3, F.F_GLOBAL).triples(F.FINAL))
        exact = search_all(shards, lambda h: h.search_exact(pat, F.F_GLOBAL).triples(F.FINAL))
        for r in range(world):
            assert ham[r] == tup(oracle.substitutions(pat, hay, 3))
            assert exact[r] == [(i, i + m, 0) for i in oracle.search_exact(pat, hay)]
    finally:
        for h in shards:
            h.close()


def test_multi_rank_lp_and_dna_routes(cuda_device):
    world, n = 3, 1 << 16
    pat, hay = _seamy_corpus(5, n, DNA, 8, 3, world, 8)    # m // (k+1) = 2 -> LP route; many overlapping groups
    shards = _make_world(hay, world, 8 + 3)
    try:
        exp = oracle.find_near_matches(pat, hay, max_l_dist=3)
        got = search_all(shards, lambda h: h.search_levenshtein(pat, 3, F.F_GLOBAL).triples(F.FINAL))
        assert all(g == exp for g in got)
    finally:
        for h in shards:
            h.close()
    pat, hay = _seamy_corpus(6, 1 << 18, DNA, 20, 2, 2, 64)  # dense filter route, thousands of groups per shard
    shards = _make_world(hay, 2, 22)
    try:
        exp = oracle.find_near_matches(pat, hay, max_l_dist=2)
        got = search_all(shards, lambda h: h.search_levenshtein(pat, 2, F.F_GLOBAL).triples(F.FINAL))
        assert all(g == exp for g in got)
    finally:
        for h in shards:
            h.close()


def test_local_world_refuses_what_needs_the_staged_path(cuda_device):
    """More groups than a slot holds: a multi-process world falls back to the staged NCCL path; the in-process
    world has no NCCL and must say so (loudly, on every rank) instead of returning a partial list."""
    hay = np.frombuffer(b"ab" * 40000, dtype=np.uint8)
    shards = _make_world(hay, 2, 4)
    try:
        with pytest.raises(F.UnsupportedError):
            search_all(shards, lambda h: h.search_exact(b"ab", F.F_GLOBAL).count(F.FINAL))
    finally:
        for h in shards:
            h.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, devices, q):
    try:
        from fuzzysearch_b200.sharding import init_shard_world_ipc
        n, m, k = (1 << 22) + 40, 20, 2
        pat, hay = _seamy_corpus(31, n, ASCII, m, k, world, 256)
        blo, bhi, lo, hi = shard_bounds(n, world, rank, m + k)
        exp = oracle.find_near_matches(pat, hay, max_l_dist=k)
        exp_ham = tup(oracle.substitutions(pat, hay, 3))
        # (1) the NCCL-free world: CUDA IPC handles over the TCP rendezvous -- also between processes on ONE GPU
        hs = F.Haystack.from_host(hay[blo:bhi], device=devices[rank], buf_lo=blo, global_len=n, own_lo=lo, own_hi=hi)
        ok = init_shard_world_ipc(hs, rank, world, "127.0.0.1", port) and hs.p2p_enabled()
        for _ in range(3):
            ok = ok and hs.search_levenshtein(pat, k, F.F_GLOBAL).triples(F.FINAL) == exp
        ok = ok and hs.search_hamming(pat, 3, F.F_GLOBAL).triples(F.FINAL) == exp_ham
        try:  # more groups than a slot holds and no NCCL behind this world: must refuse, on every rank
            hs.search_exact(pat[:1], F.F_GLOBAL).count(F.FINAL)
            ok = False
        except F.UnsupportedError:
            pass
        hs.close()
        nccl = len(set(devices)) == world  # (2) the NCCL-bootstrapped world needs one GPU per rank
        if nccl:
            hs = F.Haystack.from_host(hay[blo:bhi], device=devices[rank], buf_lo=blo, global_len=n, own_lo=lo, own_hi=hi)
            init_shard_comm(hs, rank, world, "127.0.0.1", port)   # torch-free bootstrap of the NCCL id
            ok = ok and hs.p2p_enabled()
            ok = ok and hs.search_levenshtein(pat, k, F.F_GLOBAL).triples(F.FINAL) == exp
            ok = ok and hs.search_hamming(pat, 3, F.F_GLOBAL).triples(F.FINAL) == exp_ham
            # more groups than a slot holds -> every rank takes the staged NCCL path together
            many = hs.search_exact(pat[:1], F.F_GLOBAL).triples(F.FINAL)
            ok = ok and many == [(i, i + 1, 0) for i in oracle.search_exact(pat[:1], hay)] and len(many) > 4096
            hs.close()
        q.put((rank, bool(ok), len(exp), nccl))
    except BaseException as e:  # noqa: BLE001
        q.put((rank, False, repr(e), False))
        raise


def test_global_multi_process_world(cuda_device):
    """One process per rank (the production layout).  With two GPUs: rank r on GPU r, both the NCCL-free and the
    NCCL-bootstrapped world.  With ONE GPU: both processes on device 0 through the NCCL-free world (CUDA IPC between
    processes sharing a GPU) -- so this test never skips."""
    import multiprocessing as mp
    needs_real_gpu("CUDA IPC / NCCL between processes")
    devices = [0, 1] if F.device_count() >= 2 else [0, 0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, devices, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in results), results


@pytest.mark.parametrize("world", [2, 3, 5])
def test_seam_rows_interleave_between_runs(cuda_device, world):
    """k_merge's binary search: a row of rank q whose hull starts INSIDE the range of hull starts of another rank's
    run.  Ranks own matches by ANCHOR, so it takes an occurrence that starts before the seam but is found only
    through its last n-gram (anchored after the seam: rank q) and a second, later-starting occurrence whose first
    n-gram is anchored before the seam (rank q - 1), plus an unrelated earlier group on rank q - 1.  A pattern of
    period 10 lets the two occurrences overlap."""
    m, k, n = 20, 2, 1 << 14
    pat = b"abcdefghij" * 2
    rng = np.random.default_rng(3)
    hay = rng.integers(48, 58, size=n, dtype=np.uint8)  # digits: no accidental hits
    for r in range(1, world):
        seam = shard_bounds(n, world, r, m + k)[2]
        hay[seam - 300:seam - 280] = np.frombuffer(pat, dtype=np.uint8)      # an earlier, isolated group of rank r - 1
        run = bytearray(b"abcdefghij" * 3)                                   # occurrences at seam - 12 and seam - 2
        run[2] = ord("X")                                                    # ... the first one keeps only its third
        run[8] = ord("Y")                                                    #     n-gram, anchored at `seam`
        hay[seam - 12:seam + 18] = np.frombuffer(bytes(run), dtype=np.uint8)
    shards = _make_world(hay, world, m + k)
    try:
        exp = oracle.find_near_matches(pat, hay, max_l_dist=k)
        assert len(exp) == 2 * (world - 1)
        got = search_all(shards, lambda h: h.search_levenshtein(pat, k, F.F_GLOBAL).triples(F.FINAL))
        for r in range(world):
            assert got[r] == exp, r
        # the rows really interleave: rank r's own list ends with a group starting after rank r + 1's first one
        own = search_all(shards, lambda h: h.search_levenshtein(pat, k).group_rows())
        for r in range(world - 1):
            assert own[r][-1][3] > own[r + 1][0][3] > own[r][0][3], (r, own[r][:, 3], own[r + 1][:, 3])
    finally:
        for h in shards:
            h.close()

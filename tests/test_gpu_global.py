"""FZB_F_GLOBAL: the in-library NCCL reduction of per-shard groups (fused behind the kernels, or staged)."""
import os
import socket

import numpy as np
import pytest

import oracle
from corpus import ASCII, DNA, make_corpus
from fuzzysearch_b200 import _native as F
from fuzzysearch_b200.sharding import shard_bounds
from parity import tup

pytestmark = pytest.mark.gpu


def test_global_flag_world_of_one(cuda_device):
    """A communicator of one rank exercises the fused and the staged paths on a single GPU."""
    pat, hay, _ = make_corpus(4, 1 << 20, ASCII, 20, 64, 3)
    hs = F.Haystack.from_host(hay)
    with pytest.raises(ValueError):
        hs.search_levenshtein(pat, 2, F.F_GLOBAL)      # no communicator yet
    hs.comm_init(F.nccl_unique_id(), 0, 1)
    local = hs.search_levenshtein(pat, 2).triples(F.FINAL)
    assert hs.search_levenshtein(pat, 2, F.F_GLOBAL).triples(F.FINAL) == local                 # fused
    assert hs.search_levenshtein(pat, 2, F.F_GLOBAL | F.F_FORCE_LP).triples(F.FINAL) == \
        hs.search_levenshtein(pat, 2, F.F_FORCE_LP).triples(F.FINAL)                           # staged
    ham = hs.search_hamming(pat, 3).triples(F.FINAL)
    assert hs.search_hamming(pat, 3, F.F_GLOBAL).triples(F.FINAL) == ham
    assert hs.search_exact(pat, F.F_GLOBAL).triples(F.FINAL) == hs.search_exact(pat).triples(F.FINAL)
    hs.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from fuzzysearch_b200.sharding import init_shard_comm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        n, m, k = (1 << 22) + 40, 20, 2
        pat, hay, _ = make_corpus(31, n, ASCII, m, 256, 3)
        for r in range(1, world):  # overlapping copies across every seam: groups that chain over it
            seam = shard_bounds(n, world, r, m + k)[2]
            blob = pat + pat[m // 2:] + pat + pat[3:]
            hay[seam - 30:seam - 30 + len(blob)] = np.frombuffer(blob, dtype=np.uint8)
        blo, bhi, lo, hi = shard_bounds(n, world, rank, m + k)
        hs = F.Haystack.from_host(hay[blo:bhi], device=rank, buf_lo=blo, global_len=n, own_lo=lo, own_hi=hi)
        init_shard_comm(hs)
        got = hs.search_levenshtein(pat, k, F.F_GLOBAL).triples(F.FINAL)
        exp = oracle.find_near_matches(pat, hay, max_l_dist=k)
        ok = got == exp
        ham = hs.search_hamming(pat, 3, F.F_GLOBAL).triples(F.FINAL)
        ok = ok and ham == tup(oracle.substitutions(pat, hay, 3))
        q.put((rank, bool(ok), len(got)))
        hs.close()
    finally:
        dist.destroy_process_group()


def test_global_two_gpus(cuda_device):
    if F.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results), results

"""Seeded random sweep over pattern length (1..255), k, alphabets and all routes: GPU vs CPU oracle."""
import numpy as np
import pytest

import oracle
from corpus import ASCII, DNA, make_corpus
from fuzzysearch_b200 import _native as F
from parity import tup

pytestmark = pytest.mark.gpu

ALPHABETS = [b"ab", DNA, b"abcdefgh", ASCII, bytes(range(256))]


def test_levenshtein_random_sweep(cuda_device):
    rng = np.random.default_rng(2026)
    for trial in range(60):
        alphabet = ALPHABETS[int(rng.integers(len(ALPHABETS)))]
        m = int(rng.choice([1, 2, 3, 5, 8, 13, 20, 33, 64, 100, 200, 255]))
        kmax = max(0, min(m - 1, 40))
        k = int(rng.integers(0, kmax + 1)) if trial % 3 else int(rng.integers(0, min(kmax, 4) + 1))
        n = int(rng.choice([0, 1, m - 1 if m > 1 else 1, m, 1000, 40000, 150001]))
        pat, hay, _ = make_corpus(1000 + trial, n, alphabet, m, 12, min(k + 1, 6), clusters=1) if n > 4 * m else \
            (bytes(rng.integers(0, 256, size=m, dtype=np.uint8)), rng.integers(0, 256, size=n, dtype=np.uint8), [])
        if k > 0 and m // (k + 1) < 3:
            # LP route: the candidate NFA is exponential in k (in the reference as well) -- keep it small
            k = min(k, 3)
            if m // (k + 1) < 3:
                hay = hay[:4000 if len(alphabet) <= 8 else 40000]
        if len(alphabet) <= 8 and k > 4:
            hay = hay[:6000]  # almost every position is an n-gram hit with an O(m^2) expansion behind it
        cpu = oracle.levenshtein_raw(pat, hay, k)
        assert len(cpu) < 2_000_000
        hs = F.Haystack.from_host(hay)
        for flags in (0, F.F_FORCE_DENSE, F.F_FORCE_SAMPLED):
            res = hs.search_levenshtein(pat, k, flags)
            got = res.triples(F.RAW)
            ctx = (trial, m, k, len(hay), len(alphabet), flags, res.stats()["route"])
            if res.stats()["route"] == "lp":
                assert sorted(got) == sorted(tup(cpu)), ctx
            else:
                assert got == tup(cpu), ctx
            assert res.triples(F.FINAL) == tup(oracle.consolidate(cpu)), ctx
            res.close()
        hs.close()


def test_hamming_random_sweep(cuda_device):
    rng = np.random.default_rng(77)
    for trial in range(40):
        alphabet = ALPHABETS[int(rng.integers(len(ALPHABETS)))]
        m = int(rng.choice([1, 2, 4, 7, 11, 16, 32, 39, 64, 128, 255]))
        k = int(rng.integers(0, 9))
        n = int(rng.choice([0, m, 999, 70000, 200003]))
        if n > 4 * m:
            pat, hay, _ = make_corpus(500 + trial, n, alphabet, m, 16, k + 1, subs_only=True, clusters=1)
        else:
            pat = bytes(rng.integers(0, 256, size=m, dtype=np.uint8))
            hay = rng.integers(0, 256, size=n, dtype=np.uint8)
        cpu = tup(oracle.substitutions(pat, hay, k))
        hs = F.Haystack.from_host(hay)
        for flags in (0, F.F_FORCE_DENSE):
            res = hs.search_hamming(pat, k, flags)
            assert res.triples(F.RAW) == cpu, (trial, m, k, n, len(alphabet), flags)
            res.close()
        hs.close()


def test_generic_random_sweep(cuda_device):
    rng = np.random.default_rng(5150)
    for trial in range(40):
        alphabet = ALPHABETS[int(rng.integers(1, len(ALPHABETS)))]
        m = int(rng.choice([2, 4, 9, 16, 24, 40]))
        subs, ins, dels = (int(x) for x in rng.integers(0, 4, size=3))
        l = int(rng.integers(1, 5))
        subs, ins, dels, l = oracle.normalize_params(subs, ins, dels, l)
        if l == 0:
            continue
        n = int(rng.choice([0, 50, 3000, 30000]))
        if len(alphabet) <= 4:
            n = min(n, 3000)
        if n > 4 * m:
            pat, hay, _ = make_corpus(900 + trial, n, alphabet, m, 8, l + 1, clusters=1)
        else:
            pat = bytes(rng.integers(0, 256, size=m, dtype=np.uint8))
            hay = rng.integers(0, 256, size=n, dtype=np.uint8)
        cpu = oracle.generic_raw(pat, hay, subs, ins, dels, l)
        hs = F.Haystack.from_host(hay)
        res = hs.search_generic(pat, subs, ins, dels, l)
        assert sorted(res.triples(F.RAW)) == sorted(tup(cpu)), (trial, m, subs, ins, dels, l, n)
        assert res.triples(F.FINAL) == tup(oracle.consolidate(cpu))
        res.close()
        hs.close()

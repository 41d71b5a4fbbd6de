"""fzb_search_levenshtein_batch / find_near_matches_batch (BASELINE.json configs[4], scaled down):
many patterns of mixed length and k over one resident haystack, each checked against the oracle."""
import numpy as np
import pytest

import oracle
from corpus import ASCII, mutate
from fuzzysearch_b200 import _native as F, find_near_matches_batch
from parity import tup

pytestmark = pytest.mark.gpu


def test_batch_matches_oracle(cuda_device):
    rng = np.random.default_rng(404)
    n = 1 << 20
    alpha = np.frombuffer(ASCII, dtype=np.uint8)
    hay = alpha[rng.integers(0, len(alpha), size=n)].copy()
    pats, ks = [], []
    for i in range(48):
        m = int(rng.integers(8, 65))
        k = int(rng.integers(1, 5))
        pat = bytes(alpha[rng.integers(0, len(alpha), size=m)])
        pats.append(pat)
        ks.append(k)
        for _ in range(8):  # 8 planted near-matches per pattern
            pos = int(rng.integers(100, n - 200))
            v = mutate(rng, pat, ASCII, int(rng.integers(0, k + 2)))
            hay[pos:pos + len(v)] = np.frombuffer(v, dtype=np.uint8)
    hs = F.Haystack.from_host(hay)
    results, stats = hs.search_levenshtein_batch(pats, ks)
    assert stats["route"] == "batch" and stats["n_launches"] >= len(pats)
    routes = set()
    total = 0
    for pat, k, res in zip(pats, ks, results):
        raw = oracle.levenshtein_raw(pat, hay, k)
        routes.add(res.stats()["route"])
        if res.stats()["route"] == "lp":
            assert sorted(res.triples(F.RAW)) == sorted(tup(raw))
        else:
            assert res.triples(F.RAW) == tup(raw)
        assert res.triples(F.FINAL) == tup(oracle.consolidate(raw))
        total += res.count(F.FINAL)
        res.close()
    assert total >= 48 * 4
    assert {"lp", "ngrams/sampled-filter", "ngrams/dense-filter"} <= routes
    hs.close()
    got = find_near_matches_batch(pats[:5], hay.tobytes(), ks[:5])
    for pat, k, ms in zip(pats, ks, got):
        assert [(m.start, m.end, m.dist) for m in ms] == oracle.find_near_matches(pat, hay, max_l_dist=k)

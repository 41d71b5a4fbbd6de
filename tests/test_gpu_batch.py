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
    assert stats["route"] == "batch"
    # ONE scan for all the patterns the q-sample lemma covers, ONE for the other n-gram-route patterns (n-gram
    # prefixes at every position)
    per_route = {}
    for r in results:
        per_route[r.stats()["route"]] = per_route.get(r.stats()["route"], 0) + 1
    assert per_route.get("ngrams/sampled-filter", 0) >= 24 and per_route.get("ngrams/dense-filter", 0) >= 4
    assert stats["bytes_scanned"] == sum(r.stats()["bytes_scanned"] for r in results)
    # ... and the LP-route patterns share scans of 64 patterns: three or four passes over the haystack in all
    assert stats["bytes_scanned"] <= n * 4
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


def test_batch_shared_scan_edge_cases(cuda_device):
    """Patterns that share 4-grams (prefixes of one another, duplicates, repetitive patterns whose grams occur at
    many offsets), a k = 0 entry (exact route, returns every overlapping occurrence), matches at both ends of
    the sequence and clusters -- all through one call; every list must equal the single-pattern search."""
    rng = np.random.default_rng(7)
    n = 1 << 22
    alpha = np.frombuffer(ASCII, dtype=np.uint8)
    hay = alpha[rng.integers(0, len(alpha), size=n)].copy()
    base = bytes(alpha[rng.integers(0, len(alpha), size=64)])
    pats = [base, base[:40], base[:20], base[10:50], base, b"abcabcabcabcabcabcabcabcabcabcabcabc", b"a" * 30,
            base[3:23], b"needle-in-a-haystack", b"xy" * 12]
    ks = [4, 3, 2, 2, 1, 3, 2, 0, 1, 2]
    for i in range(120):
        m = int(rng.integers(12, 65))
        pats.append(bytes(alpha[rng.integers(0, len(alpha), size=m)]))
        ks.append(int(rng.integers(1, 5)))
    for pat, k in zip(pats, ks):
        for _ in range(4):
            pos = int(rng.integers(100, n - 200))
            v = mutate(rng, pat, ASCII, int(rng.integers(0, k + 2)))
            hay[pos:pos + len(v)] = np.frombuffer(v, dtype=np.uint8)
    hay[:64] = np.frombuffer(base, dtype=np.uint8)              # at the very start
    hay[n - 60:] = np.frombuffer(base[:60], dtype=np.uint8)     # truncated at the very end
    blob = (b"abc" * 40) + (b"a" * 70) + (b"xy" * 30)
    hay[5000:5000 + len(blob)] = np.frombuffer(blob, dtype=np.uint8)  # long chains of overlapping matches
    hs = F.Haystack.from_host(hay)
    results, stats = hs.search_levenshtein_batch(pats, ks)
    shared = 0
    for pat, k, res in zip(pats, ks, results):
        one = hs.search_levenshtein(pat, k)
        shared += res.stats()["route"] == "ngrams/sampled-filter" and one.stats()["route"] == "ngrams/sampled-filter"
        if one.stats()["route"] == "lp":
            assert sorted(res.triples(F.RAW)) == sorted(one.triples(F.RAW))
        else:
            assert res.arrays(F.RAW, anchors=True)[3].tolist() == one.arrays(F.RAW, anchors=True)[3].tolist()
            assert res.triples(F.RAW) == one.triples(F.RAW), (pat, k)
        assert res.triples(F.FINAL) == one.triples(F.FINAL), (pat, k)
        raw = oracle.levenshtein_raw(pat, hay, k)
        assert res.triples(F.FINAL) == tup(oracle.consolidate(raw)) if k else res.triples(F.RAW) == tup(raw)
        one.close()
        res.close()
    assert shared >= 100
    hs.close()


def test_batch_lp_pass_with_large_budgets(cuda_device):
    """LP-route patterns with max_l_dist 5..8 share a pass too (`k_lp_verify_multi<8>`: nine automaton masks per
    lane); some end inside the sequence's last bytes (end-of-sequence acceptance, levenshtein.py:145-148)."""
    rng = np.random.default_rng(58)
    for alphabet, n in ((ASCII, 20000), (b"abcdefgh", 3000)):
        alpha = np.frombuffer(alphabet, dtype=np.uint8)
        hay = alpha[rng.integers(0, len(alpha), size=n)].copy()
        pats, ks = [], []
        for k in (5, 6, 7, 8, 5, 8, 2, 3):
            m = int(rng.integers(k + 1, min(3 * (k + 1) - 1, 31 - k) + 1))  # m // (k + 1) < 3 and m + k <= 31
            pat = bytes(alpha[rng.integers(0, len(alpha), size=m)])
            for _ in range(3):
                pos = int(rng.integers(50, n - 100))
                v = mutate(rng, pat, alphabet, int(rng.integers(0, k + 1)))
                hay[pos:pos + len(v)] = np.frombuffer(v, dtype=np.uint8)
            pats.append(pat)
            ks.append(k)
        hay[n - 7:] = np.frombuffer(pats[0][:7], dtype=np.uint8)  # a prefix of pattern 0 runs into the end
        hs = F.Haystack.from_host(hay)
        results, _ = hs.search_levenshtein_batch(pats, ks)
        for pat, k, res in zip(pats, ks, results):
            raw = oracle.levenshtein_raw(pat, hay, k)
            assert res.stats()["route"] == "lp"
            assert sorted(res.triples(F.RAW)) == sorted(tup(raw)), (pat, k)
            assert res.triples(F.FINAL) == tup(oracle.consolidate(raw)), (pat, k)
            res.close()
        hs.close()


def test_batch_dense_pass_flushes_and_overflows_its_cta_buffer(cuda_device):
    """k_filter_mdense buffers hits per CTA (3 072 entries, flushed at 1 024, straight to the global list when full).
    Thousands of occurrences packed into ONE 64 KiB tile drive one CTA through all three paths; the lists must still
    equal the single-pattern searches and the oracle."""
    rng = np.random.default_rng(91)
    n = 1 << 20
    alpha = np.frombuffer(ASCII, dtype=np.uint8)
    hay = alpha[rng.integers(0, len(alpha), size=n)].copy()
    pats = [b"QWERTYUIOPAS", b"zxcvbnmlkjhg", b"0192837465ab"]   # m = 12, k = 3: L = 3, the lemma does not hold
    ks = [3, 3, 3]
    hay[100:100 + 12 * 1500] = np.frombuffer(pats[0] * 1500, dtype=np.uint8)            # 6 000 n-gram hits in tile 0
    off = 3 * (1 << 16) + 40
    hay[off:off + 13 * 600] = np.frombuffer((pats[1] + b"-") * 600, dtype=np.uint8)     # 2 400 in another tile
    for i in range(50):
        pos = int(rng.integers(1 << 18, n - 100))
        hay[pos:pos + 12] = np.frombuffer(mutate(rng, pats[2], ASCII, int(rng.integers(0, 4)))[:12].ljust(12, b"#"),
                                          dtype=np.uint8)
    hs = F.Haystack.from_host(hay)
    results, _ = hs.search_levenshtein_batch(pats, ks)
    for pat, k, res in zip(pats, ks, results):
        assert res.stats()["route"] == "ngrams/dense-filter"
        one = hs.search_levenshtein(pat, k)
        assert res.triples(F.RAW) == one.triples(F.RAW), pat
        raw = oracle.levenshtein_raw(pat, hay, k)
        assert res.triples(F.RAW) == tup(raw), pat
        assert res.triples(F.FINAL) == tup(oracle.consolidate(raw)), pat
        one.close()
        res.close()
    hs.close()

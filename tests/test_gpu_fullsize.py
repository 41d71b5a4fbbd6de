"""BASELINE.json's full sizes (4 GiB per GPU) through size-independent properties:

* two independent candidate filters (sampled vs dense; counting vs brute force) must produce the
  identical raw stream over the whole 4 GiB;
* around every planted near-match the device's raw matches equal the CPU oracle's on a local window;
* sharding invariance: the same device buffer searched as 3 shards (adopted, with halo) == whole;
* idempotence, sortedness, and every reported match is a true near-match (edit distance re-checked).
"""
import numpy as np
import pytest

import oracle
from fuzzysearch_b200 import _native as F
from conftest import needs_real_gpu
from parity import tup

pytestmark = pytest.mark.gpu

GiB = 1 << 30
ASCII = bytes(range(32, 127))
DNA = b"ACGT"


def _edit_distance(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def _plant(hs, n, pat, alphabet, nplants, max_edits, subs_only, seed):
    rng = np.random.default_rng(seed)
    m = len(pat)
    positions = []
    span = (n - 8 * m) // nplants
    for i in range(nplants):
        pos = 4 * m + i * span + int(rng.integers(0, span - 4 * m))
        v = bytearray(pat)
        for _ in range(int(rng.integers(0, max_edits + 1))):
            op = 0 if subs_only else int(rng.integers(3))
            if op == 0:
                v[int(rng.integers(len(v)))] = alphabet[int(rng.integers(len(alphabet)))]
            elif op == 1:
                v.insert(int(rng.integers(len(v) + 1)), alphabet[int(rng.integers(len(alphabet)))])
            else:
                del v[int(rng.integers(len(v)))]
        hs.write(pos, bytes(v))
        positions.append(pos)
    # both global ends and a cluster of overlapping copies
    hs.write(0, pat[1:])
    hs.write(n - m + 1, pat[:m - 1])
    hs.write(n // 3, pat + pat[m // 2:] + pat)
    return positions


def _raw_with_anchor(res):
    s, e, d, ng, ix = res.arrays(F.RAW, anchors=True)
    return list(zip(ng.tolist(), ix.tolist(), s.tolist(), e.tolist(), d.tolist()))


def test_levenshtein_4gib_ascii(cuda_device):
    needs_real_gpu("4 GiB input")
    n, m, k = 4 * GiB, 20, 2
    hs = F.Haystack.alloc(n)
    hs.fill_synthetic(ASCII, 20260923)
    pat = bytes(np.random.default_rng(3).integers(32, 127, size=m, dtype=np.uint8))
    positions = _plant(hs, n, pat, ASCII, 512, k + 1, False, 11)

    res = hs.search_levenshtein(pat, k)
    assert res.stats()["route"] == "ngrams/sampled-filter"
    raw = _raw_with_anchor(res)
    final = res.triples(F.FINAL)
    assert len(raw) > 512 and len(final) > 256
    # idempotence + independent filter
    res2 = hs.search_levenshtein(pat, k)
    assert _raw_with_anchor(res2) == raw and res2.triples(F.FINAL) == final
    dense = hs.search_levenshtein(pat, k, F.F_FORCE_DENSE)
    assert dense.stats()["route"] == "ngrams/dense-filter"
    assert _raw_with_anchor(dense) == raw
    # order, consolidation, truth of every match
    assert raw == sorted(raw, key=lambda r: (r[0], r[1]))
    assert final == sorted(final)
    assert final == tup(oracle.consolidate(np.array([r[2:] for r in raw], dtype=np.int64)))
    for _, _, s, e, d in raw[:2000]:
        assert d <= k and _edit_distance(pat, hs.read(s, e - s)) <= d
    # local windows around plants vs the CPU oracle (anchors with full context only)
    by_idx = {}
    for r in raw:
        by_idx.setdefault(r[1], []).append(r)
    margin = 64
    for pos in positions[::4]:
        lo, hi = pos - margin, pos + m + k + margin
        window = hs.read(lo, hi - lo)
        wraw, wng, wix = oracle.levenshtein_ngrams_raw(pat, window, k, with_anchor=True)
        exp = sorted((int(g), int(i) + lo, int(s) + lo, int(e) + lo, int(d))
                     for (s, e, d), g, i in zip(tup(wraw), wng, wix) if margin // 2 <= i < len(window) - margin // 2 - m)
        got = sorted(r for i in range(lo + margin // 2, hi - margin // 2 - m) for r in by_idx.get(i, []))
        assert got == exp, pos
    # global ends: compare with the oracle on the first / last 4 KiB
    head = hs.read(0, 4096)
    exp_head = [t for t in tup(oracle.levenshtein_ngrams_raw(pat, head, k)) if t[1] < 2048]
    assert sorted(r[2:] for r in raw if r[3] < 2048) == sorted(exp_head)
    tail = hs.read(n - 4096, 4096)
    exp_tail = [(s + n - 4096, e + n - 4096, d) for s, e, d in tup(oracle.levenshtein_ngrams_raw(pat, tail, k))
                if s >= 2048]
    assert sorted(r[2:] for r in raw if r[2] >= n - 2048) == sorted(exp_tail)
    # the WHOLE 4 GiB raw stream against the C oracle (SURVEY 8c "big-input oracle"): element for element, in
    # generation order, with anchors; and the final list against consolidate_overlapping_matches of it
    host = np.empty(n, dtype=np.uint8)
    step = 256 << 20
    for off in range(0, n, step):
        host[off:off + step] = np.frombuffer(hs.read(off, min(step, n - off)), dtype=np.uint8)
    raw_o, ng_o, ix_o = oracle.levenshtein_ngrams_raw(pat, host, k, with_anchor=True)
    assert raw == [(int(g), int(i), int(s), int(e), int(d)) for (s, e, d), g, i in zip(raw_o.tolist(), ng_o.tolist(), ix_o.tolist())]
    assert final == tup(oracle.consolidate(raw_o))
    del host
    # sharding invariance on the same device memory
    halo = m + k
    got = []
    bounds = [0, (n // 3) // 16 * 16 + 16, (2 * n // 3) // 16 * 16, n]  # the middle seam cuts the cluster
    for i in range(3):
        lo, hi = bounds[i], bounds[i + 1]
        blo, bhi = max(0, lo - halo) // 16 * 16, min(n, hi + halo)
        sh = F.Haystack.adopt(hs.dev_ptr + blo, bhi - blo, buf_lo=blo, global_len=n, own_lo=lo, own_hi=hi)
        r = sh.search_levenshtein(pat, k, F.F_NO_FINAL)
        got += _raw_with_anchor(r)
        r.close()
        sh.close()
    assert sorted(got) == sorted(raw)
    hs.close()


def test_hamming_4gib_dna(cuda_device):
    needs_real_gpu("4 GiB input")
    n, m, k = 4 * GiB, 32, 3
    hs = F.Haystack.alloc(n)
    hs.fill_synthetic(DNA, 7)
    pat = bytes(np.frombuffer(DNA, dtype=np.uint8)[np.random.default_rng(5).integers(0, 4, size=m)])
    positions = _plant(hs, n, pat, DNA, 512, k + 1, True, 13)
    res = hs.search_hamming(pat, k)
    got = res.triples(F.RAW)
    assert len(got) > 256 and got == sorted(got) and res.triples(F.FINAL) == got
    brute = hs.search_hamming(pat, k, F.F_FORCE_DENSE)   # independent kernel: every position, exact count
    assert brute.triples(F.RAW) == got
    for s, e, d in got[:2000]:
        w = hs.read(s, m)
        assert e == s + m and d == sum(a != b for a, b in zip(w, pat)) <= k
    found = {s for s, _, _ in got}
    for pos in positions[::4]:
        w = hs.read(pos - 64, m + 128)
        exp = [(s + pos - 64, e + pos - 64, d) for s, e, d in tup(oracle.substitutions(pat, w, k))]
        assert all(s in found for s, _, _ in exp)
    # shards
    parts = []
    bounds = [0, (n // 2) // 16 * 16, n]
    for i in range(2):
        lo, hi = bounds[i], bounds[i + 1]
        blo, bhi = max(0, lo - m) // 128 * 128, min(n, hi + m)
        sh = F.Haystack.adopt(hs.dev_ptr + blo, bhi - blo, buf_lo=blo, global_len=n, own_lo=lo, own_hi=hi)
        parts += sh.search_hamming(pat, k).triples(F.RAW)
        sh.close()
    assert sorted(parts) == got
    hs.close()

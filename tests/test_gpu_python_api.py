"""The Python drop-in surface under threads and with large pageable inputs (staged upload ring)."""
import threading

import numpy as np
import pytest

import oracle
from corpus import ASCII, make_corpus
from fuzzysearch_b200 import find_near_matches, has_near_match

pytestmark = pytest.mark.gpu


def test_find_near_matches_is_thread_safe(cuda_device):
    """Four threads share the cached device workspace: every call must return ITS OWN haystack's matches
    (the workspace was shared without a lock while ctypes drops the GIL)."""
    jobs = []
    for t in range(4):
        pat, hay, _ = make_corpus(100 + t, (1 << 20) + 4096 * t, ASCII, 20, 32, 3)
        jobs.append((pat, hay.tobytes(), oracle.find_near_matches(pat, hay, max_l_dist=2)))
    errors = []

    def run(t):
        pat, hay, exp = jobs[t]
        try:
            for _ in range(6):
                got = [(m.start, m.end, m.dist) for m in find_near_matches(pat, hay, max_l_dist=2)]
                assert got == exp and len(exp) > 10
                assert has_near_match(pat, hay, max_l_dist=2)
        except BaseException as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=run, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_threads_share_one_resident_sequence(cuda_device):
    """Four threads search ONE DeviceSequence with different patterns and limits at the same time (the natural way
    to use a resident corpus): every call must return its own pattern's matches.  Byte sequence: the library
    serialises the calls on the handle; wide str: the Python lock also covers the re-reduction of the sequence to
    each pattern's alphabet."""
    from fuzzysearch_b200 import DeviceSequence, find_near_matches_batch, search_exact
    from symbols import reduce_to_bytes
    rng = np.random.default_rng(77)
    n = 1 << 21
    alpha = np.frombuffer(ASCII, dtype=np.uint8)
    hay = alpha[rng.integers(0, len(alpha), size=n)].copy()
    pats = [bytes(alpha[rng.integers(0, len(alpha), size=m)]) for m in (20, 9, 32, 14)]
    ks = [2, 3, 3, 1]
    for i, p in enumerate(pats):
        for j in range(6):
            pos = 5000 + 30000 * (4 * j + i)
            v = bytearray(p)
            if j % 2:
                v[len(v) // 2] ^= 1
            hay[pos:pos + len(v)] = np.frombuffer(bytes(v), dtype=np.uint8)
    hay_b = hay.tobytes()
    letters = [chr(c) for c in range(0x3B1, 0x3C9)] + ["\U0001F642", "x"]
    text = "".join(letters[i] for i in rng.integers(0, len(letters), size=200000))
    wpats = []
    for i in range(4):  # four patterns over four DIFFERENT alphabets (each thread forces a re-reduction)
        sub = letters[6 * i:6 * i + 6] + ["x"]
        w = "".join(sub[j] for j in rng.integers(0, len(sub), size=16))
        wpats.append(w)
        text = text[:7000 * (i + 1)] + w + text[7000 * (i + 1) + 16:]
    exp_b = [oracle.find_near_matches(p, hay_b, max_l_dist=k) for p, k in zip(pats, ks)]
    exp_h = [oracle.find_near_matches(p, hay_b, 2, 0, 0) for p in pats]
    exp_w = [oracle.find_near_matches(*reduce_to_bytes(w, text), max_l_dist=1) for w in wpats]
    assert all(len(e) >= 6 for e in exp_b) and all(len(e) >= 1 for e in exp_w)
    ds, dw = DeviceSequence(hay_b), DeviceSequence(text)
    errors = []

    def run(t):
        try:
            for _ in range(5):
                got = find_near_matches(pats[t], ds, max_l_dist=ks[t])
                assert [(m.start, m.end, m.dist) for m in got] == exp_b[t]
                assert all(m.matched == hay_b[m.start:m.end] for m in got)
                got = find_near_matches(pats[t], ds, max_substitutions=2, max_insertions=0, max_deletions=0)
                assert [(m.start, m.end, m.dist) for m in got] == exp_h[t]
                assert has_near_match(pats[t], ds, max_l_dist=ks[t])
                assert search_exact(pats[t][:6], ds, 100, n - 100) == oracle.search_exact(pats[t][:6], hay_b, 100, n - 100)
                got = find_near_matches(wpats[t], dw, max_l_dist=1)
                assert [(m.start, m.end, m.dist) for m in got] == exp_w[t]
                assert all(m.matched == text[m.start:m.end] for m in got)
                if t == 0:
                    b = find_near_matches_batch(pats, ds, ks)
                    assert [[(m.start, m.end, m.dist) for m in x] for x in b] == exp_b
        except BaseException as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=run, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    ds.close()
    dw.close()
    assert not errors, errors


def test_large_pageable_input_goes_through_the_upload_ring(cuda_device):
    """200 MiB of pageable memory (more than the three 64 MiB ring buffers): the staged, multi-threaded upload must
    deliver every byte in order -- matches planted in every ring slice and across slice boundaries are found."""
    n = 200 << 20
    rng = np.random.default_rng(5)
    hay = rng.integers(32, 127, size=n, dtype=np.uint8)
    pat = bytes(rng.integers(32, 127, size=24, dtype=np.uint8))
    slice_bytes = 64 << 20
    spots = [0, 12345, slice_bytes - 10, slice_bytes + 100, 2 * slice_bytes - 24, 2 * slice_bytes + 70, 3 * slice_bytes - 1,
             n - 24, (2 << 20) - 12, 150 << 20]
    for pos in spots:
        hay[pos:pos + 24] = np.frombuffer(pat, dtype=np.uint8)
    data = bytearray(hay.tobytes())
    got = [(m.start, m.end, m.dist, bytes(m.matched)) for m in find_near_matches(pat, data, max_l_dist=2)]
    assert got == [(p, p + 24, 0, pat) for p in sorted(spots)]


def test_has_near_match_all_routes_vs_oracle(cuda_device):
    """has_near_match == bool(find_near_matches) for every search class, on sequences with / without matches."""
    rng = np.random.default_rng(21)
    cases = [dict(max_l_dist=0), dict(max_l_dist=2), dict(max_substitutions=2, max_insertions=0, max_deletions=0),
             dict(max_substitutions=1, max_insertions=2, max_deletions=1), dict(max_l_dist=3),
             dict(max_substitutions=0, max_insertions=1, max_deletions=1, max_l_dist=2)]
    for trial in range(30):
        m = int(rng.integers(4, 30))
        n = int(rng.integers(0, 5000))
        alpha = ASCII if trial % 2 else b"ACGT"
        a = np.frombuffer(alpha, dtype=np.uint8)
        hay = a[rng.integers(0, len(a), size=n)].copy()
        pat = bytes(a[rng.integers(0, len(a), size=m)])
        if trial % 3 == 0 and n > 2 * m:
            pos = int(rng.integers(0, n - m))
            hay[pos:pos + m] = np.frombuffer(pat, dtype=np.uint8)
            hay[pos + m // 2] = a[0]
        for kw in cases:
            exp = len(oracle.find_near_matches(pat, hay, **kw)) > 0
            assert has_near_match(pat, hay.tobytes(), **kw) == exp, (trial, kw, m, n)
    with pytest.raises(ValueError):
        has_near_match(b"", b"abc", max_l_dist=1)
    assert has_near_match(b"ab", b"", max_l_dist=2)  # k >= len(pattern): the empty match at index 0


def test_has_near_match_stops_early_on_a_long_sequence(cuda_device):
    """Several chunks (64 MiB, then the rest): a match in the first chunk, only in the last one, at the chunk
    seam, nowhere."""
    import time
    from fuzzysearch_b200 import DeviceSequence
    n = 300 << 20
    rng = np.random.default_rng(8)
    hay = rng.integers(32, 127, size=n, dtype=np.uint8)
    pat = bytes(rng.integers(32, 127, size=20, dtype=np.uint8))
    seq = DeviceSequence(hay)
    kw = dict(max_l_dist=2)
    assert not has_near_match(pat, seq, **kw)
    assert not has_near_match(pat, seq, max_substitutions=3, max_insertions=0, max_deletions=0)
    t0 = time.perf_counter()
    for _ in range(5):
        has_near_match(pat, seq, **kw)
    t_none = (time.perf_counter() - t0) / 5
    seam = 64 << 20
    for pos in (n - 20, seam - 7, 1 << 20):   # last chunk / straddling the first seam / first chunk
        seq.haystack.write(pos, pat)
        assert has_near_match(pat, seq, **kw) and has_near_match(pat, seq, max_l_dist=0)
        assert has_near_match(pat, seq, max_substitutions=3, max_insertions=0, max_deletions=0)
    t0 = time.perf_counter()
    for _ in range(5):
        assert has_near_match(pat, seq, **kw)
    t_early = (time.perf_counter() - t0) / 5
    assert t_early < t_none, (t_early, t_none)  # the first chunk answers: the other 236 MiB are never read
    seq.close()


def test_has_near_match_chunk_seams_at_small_scale(cuda_device, monkeypatch):
    """The chunked early termination with the chunk size shrunk by the test hook (FZB_HAS_CHUNK_BYTES): chunks of
    256, 1024, 4096, ... bytes, the ONLY near-match planted at every delta around several chunk seams, every
    search class -- each chunk is a view of the resident buffer with its own halo, so nothing may be lost or
    invented at a seam."""
    monkeypatch.setenv("FZB_HAS_CHUNK_BYTES", "256")
    rng = np.random.default_rng(77)
    m = 12
    pat = b"needle-hay-x"
    cases = [dict(max_l_dist=0), dict(max_l_dist=2), dict(max_substitutions=2, max_insertions=0, max_deletions=0),
             dict(max_substitutions=1, max_insertions=1, max_deletions=1), dict(max_l_dist=4)]
    n = 9000
    base = rng.integers(48, 58, size=n, dtype=np.uint8)  # digits: no accidental matches
    for kw in cases:
        assert has_near_match(pat, base.tobytes(), **kw) is (kw.get("max_l_dist", 0) >= m)
    seams = [256, 256 + 1024, 256 + 1024 + 4096]
    for seam in seams:
        for delta in range(-m - 4, 5):
            hay = base.copy()
            v = bytearray(pat)
            v[5] = ord("#")  # one substitution: found by every class but the exact one
            hay[seam + delta:seam + delta + m] = np.frombuffer(bytes(v), dtype=np.uint8)
            for kw in cases:
                exp = len(oracle.find_near_matches(pat, hay, **kw)) > 0
                assert has_near_match(pat, hay.tobytes(), **kw) is exp, (seam, delta, kw)
            hay[seam + delta + 5] = pat[5]
            assert has_near_match(pat, hay.tobytes(), max_l_dist=0), (seam, delta)

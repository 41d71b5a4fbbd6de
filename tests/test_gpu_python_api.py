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

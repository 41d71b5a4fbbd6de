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

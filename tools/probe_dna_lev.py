"""Levenshtein n-gram search on a DNA corpus (small alphabet: the sampled filter is not selective)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fuzzysearch_b200 import _native as F
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256 << 20
hs = F.Haystack.alloc(n)
hs.fill_synthetic(b"ACGT", 1)
pat = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(1).integers(0, 4, size=20)])
for flags, name in ((0, "sampled"), (F.F_FORCE_DENSE, "dense")):
    for it in range(3):
        t0 = time.perf_counter()
        r = hs.search_levenshtein(pat, 2, flags)
        dt = time.perf_counter() - t0
        st = r.stats()
        print(name, "call %.2f ms  gpu %.2f  filter %.3f  raw %d final %d cand %d  -> %.1f GB/s" %
              (dt * 1e3, st["gpu_ms"], st["filter_ms"], r.count(F.RAW), r.count(F.FINAL), st["n_candidates"], n / dt / 1e9))
        r.close()

"""Break one resident 4 GiB search (the bench workload) into its parts: device events, host clock and
the phase times of k_post's last CTA (fzb_debug_counters)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from fuzzysearch_b200 import _native as F

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4 << 30
alphabet, m, k, seed = bench.ASCII, 20, 2, 20260923
hs = F.Haystack.alloc(n)
hs.fill_synthetic(alphabet, seed)
rng = np.random.default_rng(seed)
pat = bytes(np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), size=m)])
for pos, b in bench.make_plants(seed + 1, 0, n, m, k, pat, alphabet, 4096, False):
    hs.write(pos, b)
for flags, name in ((0, "final"), (F.F_NO_FINAL, "nofinal")):
    for it in range(6):
        t0 = time.perf_counter()
        r = hs.search_levenshtein(pat, k, flags)
        t1 = time.perf_counter()
        st = r.stats()
        c = r.count(F.RAW)
        nf = r.count(F.FINAL)
        r.close()
        t2 = time.perf_counter()
        cn = hs.debug_counters()
        print(name, "call %.3f ms  total %.3f ms  gpu %.3f  filter %.3f  raw %d final %d cand %d | k_post last CTA ns: "
              "rank %d ticket %d sweep %d copy %d" %
              ((t1 - t0) * 1e3, (t2 - t0) * 1e3, st["gpu_ms"], st["filter_ms"], c, nf, st["n_candidates"],
               cn[10], cn[11], cn[12], cn[13]))

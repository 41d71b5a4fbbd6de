"""Where the time of the 1024-pattern batch (BASELINE configs[4]) goes: per-route sums of the device time."""
import os
import sys
import time
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from fuzzysearch_b200 import _native as F

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4 << 30
npat = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
alphabet, seed = bench.ASCII, 20260923
hs = F.Haystack.alloc(n)
hs.fill_synthetic(alphabet, seed)
brng = np.random.default_rng(seed + 99)
alpha = np.frombuffer(alphabet, dtype=np.uint8)
pats, ks = [], []
for i in range(npat):
    bm, bk = int(brng.integers(8, 65)), int(brng.integers(1, 5))
    bp = bytes(alpha[brng.integers(0, len(alpha), size=bm)])
    pats.append(bp)
    ks.append(bk)
    for _ in range(8):
        pos = 1000 + int(brng.integers(0, n - 2000))
        hs.write(pos, bench.mutate(brng, bp, alphabet, int(brng.integers(0, bk + 2)), False))
for it in range(2):
    t0 = time.perf_counter()
    results, st = hs.search_levenshtein_batch(pats, ks)
    dt = time.perf_counter() - t0
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0, 0])
    for r in results:
        s = r.stats()
        a = agg[s["route"]]
        a[0] += 1
        a[1] += s["gpu_ms"]
        a[2] += s["filter_ms"]
        a[3] += r.count(F.RAW)
        a[4] += r.count(F.FINAL)
        r.close()
    print("batch wall %.1f ms, device %.1f ms" % (dt * 1e3, st["gpu_ms"]))
    for route, a in sorted(agg.items()):
        print("  %-24s patterns %4d  gpu %.1f ms  filter %.1f ms  raw %d final %d" % (route, a[0], a[1], a[2], a[3], a[4]))

"""Break one resident 4 GiB search into its parts (device events + host clock)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fuzzysearch_b200 import _native as F
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4 << 30
hs = F.Haystack.alloc(n)
hs.fill_synthetic(bytes(range(32, 127)), 1)
pat = bytes(np.random.default_rng(1).integers(32, 127, size=20, dtype=np.uint8))
nplants = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
for i in range(nplants):
    hs.write(1000 + i * (n // nplants - 1) // 1 % (n - 100), pat)
for flags, name in ((0, "final"), (F.F_NO_FINAL, "nofinal")):
    for it in range(6):
        t0 = time.perf_counter()
        r = hs.search_levenshtein(pat, 2, flags)
        t1 = time.perf_counter()
        st = r.stats()
        c = r.count(F.RAW)
        r.close()
        t2 = time.perf_counter()
        print(name, "call %.3f ms  total %.3f ms  gpu %.3f  filter %.3f  raw %d cand %d" %
              ((t1 - t0) * 1e3, (t2 - t0) * 1e3, st["gpu_ms"], st["filter_ms"], c, st["n_candidates"]))

"""A/B of the counter layouts of k_hamming_count (FZB_HAM_COUNTERS=nibble|sliced3|auto) on a DNA corpus:
same plants, same pattern; results must be identical; prints per-variant filter / whole-search times.
usage: python tools/probe_ham.py [n_bytes] [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from fuzzysearch_b200 import _native as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4 << 30
hs = F.Haystack.alloc(n)
hs.fill_synthetic(b"ACGT", 8)
rng = np.random.default_rng(3)
a = np.frombuffer(b"ACGT", dtype=np.uint8)
out = {"n": n, "cases": []}
for m, k in ((32, 3), (20, 1), (64, 7), (48, 5)):
    pat = bytes(a[rng.integers(0, 4, size=m)])
    for t in range(2048):  # plants with 0..k+1 substitutions
        v = bytearray(pat)
        for _ in range(int(rng.integers(0, k + 2))):
            v[int(rng.integers(0, m))] = int(a[rng.integers(0, 4)])
        hs.write(int(rng.integers(0, n - m)), bytes(v))
    case = {"m": m, "k": k}
    ref = None
    for variant in ("nibble", "sliced3", "auto", "nibble", "sliced3", "auto"):  # auto: two slices when Wc - k <= 4
        os.environ["FZB_HAM_COUNTERS"] = variant
        for _ in range(3):
            hs.search_hamming(pat, k).close()
        filt = []
        hs.timer_start()
        for _ in range(10):
            r = hs.search_hamming(pat, k)
            st = r.stats()
            filt.append(st["filter_ms"])
            r.close()
        ms = hs.timer_stop() / 10
        r = hs.search_hamming(pat, k)
        got = r.triples(F.RAW)
        cand = r.stats()["n_candidates"]
        r.close()
        if ref is None:
            ref = got
        assert got == ref, (variant, len(got), len(ref))
        case.setdefault(variant, []).append({"ms": round(ms, 4), "filter_ms": round(float(np.mean(filt)), 4),
                                             "GBps_filter": round(n / np.mean(filt) / 1e6, 1), "matches": len(got),
                                             "candidates": int(cand)})
        print(m, k, variant, case[variant][-1], flush=True)
    out["cases"].append(case)
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
        json.dump(out, f, indent=1)

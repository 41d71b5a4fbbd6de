"""H2D upload speed through fzb_haystack_upload: pageable (bytes) vs pinned host memory."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fuzzysearch_b200 import _native as F
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
hs = F.Haystack.alloc(n)
page = np.random.default_rng(0).integers(32, 127, size=n, dtype=np.uint8)
pin = F.PinnedBuffer(n)
pin.array[:] = page
for name, src in (("pageable", page), ("pinned", pin.array)):
    for it in range(3):
        t0 = time.perf_counter()
        hs.upload(src)
        dt = time.perf_counter() - t0
        print(name, "%.1f ms  %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
back = hs.read(n - 1000, 1000)
assert back == page[n - 1000:].tobytes() and hs.read(12345678, 100) == page[12345678:12345778].tobytes()
hs.upload(page)
assert hs.read(0, 4096) == page[:4096].tobytes() and hs.read(n // 2, 4096) == page[n // 2:n // 2 + 4096].tobytes()
print("content OK")

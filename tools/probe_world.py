"""Per-step breakdown of the sharded search (run under torchrun): device time, wall time and the phase times the
kernels stamp (k_post's last CTA, k_merge's wait for the peers and its merge)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from fuzzysearch_b200 import _native as F
from fuzzysearch_b200.sharding import init_shard_comm, shard_bounds

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 4 << 30
alphabet, m, k, seed = bench.ASCII, 20, 2, 20260923
n = per_gpu * world
blo, bhi, lo, hi = shard_bounds(n, world, rank, m + k)
hs = F.Haystack.alloc(bhi - blo, device=local, buf_lo=blo, global_len=n, own_lo=lo, own_hi=hi)
hs.fill_synthetic(alphabet, seed)
rng = np.random.default_rng(seed)
pat = bytes(np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), size=m)])
for pos, b in bench.make_plants(seed + 1 + rank, lo, hi, m, k, pat, alphabet, 4096, False):
    if pos >= blo and pos + len(b) <= bhi:
        hs.write(pos, b)
if os.environ.get("PROBE_TORCH"):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier()
init_shard_comm(hs)
print("rank", rank, "p2p", hs.p2p_enabled(), flush=True)
for flags, name in ((0, "local"), (F.F_GLOBAL, "global")):
    for it in range(6):
        t0 = time.perf_counter()
        r = hs.search_levenshtein(pat, k, flags)
        t1 = time.perf_counter()
        st = r.stats()
        nf = r.count(F.FINAL)
        r.close()
        cn = hs.debug_counters()
        if it >= 4:
            print("rank %d %s call %.3f ms gpu %.3f filter %.3f final %d | post: rank %d tick %d sweep %d copy %d | merge: "
                  "wait %d ns, merge %d ns, nonheads %d status %d" % (rank, name, (t1 - t0) * 1e3, st["gpu_ms"], st["filter_ms"],
                                                                    nf, cn[10], cn[11], cn[12], cn[13], cn[20], cn[21], cn[22], cn[16]),
                  flush=True)

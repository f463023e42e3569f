#!/usr/bin/env python3
"""Timeline of F contexts in flight (the `value` loop of bench.py) from a -DZK_TIMELINE build: per step the phase
boundaries in ms since a common reference.  usage: ZKMSM_DEV=1 ZKMSM_LIB=.../libzkmsm_<tag>.so tools/timeline.py [F] [steps] [logn]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkvm_b200 as zk
F = int(sys.argv[1]) if len(sys.argv) > 1 else 4
K = int(sys.argv[2]) if len(sys.argv) > 2 else 16
logn = int(sys.argv[3]) if len(sys.argv) > 3 else 20
n = 1 << logn
dev = torch.device("cuda", 0)
ctxs = [zk.Context(0) for _ in range(F)]
rng = np.random.default_rng(1)
tabs, scal = [], []
for s in range(2):
    tabs.append(zk.PointTable(ctxs[0], n).append_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8)))
    scal.append(torch.from_numpy(rng.integers(0, 256, size=(n * 32,), dtype=np.uint8)).to(dev))
parts = [torch.empty(128, dtype=torch.uint8, device=dev) for _ in range(F)]
torch.cuda.synchronize()
def run(k):
    for i in range(k):
        f = i % F
        ctxs[f].msm_table_dev(scal[i % 2].data_ptr(), tabs[i % 2], 0, n, parts[f].data_ptr())
        if i >= F - 1: ctxs[(i - (F - 1)) % F].ext_sum_compress_dev(parts[(i - (F - 1)) % F].data_ptr(), 1)
    for i in range(max(0, k - (F - 1)), k): ctxs[i % F].ext_sum_compress_dev(parts[i % F].data_ptr(), 1)
run(2 * F)
for c in ctxs: c.set_profiling(True)
print("# ctx call sort_start sort_end accum_end tail_queued (ms)", file=sys.stderr)
run(K)

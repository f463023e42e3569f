#!/usr/bin/env python3
"""One-off: a 2^26-term MSM entirely from device-generated inputs; two window widths must agree."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zkvm_b200 as zk
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n = 1 << logn
ctx = zk.Context(0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
tab = zk.PointTable(ctx, n)
step = 1 << 22
for lo in range(0, n, step):
    u = torch.randint(0, 256, (min(step, n - lo), 64), dtype=torch.uint8, device="cuda", generator=g)
    torch.cuda.synchronize(); tab.append_uniform_dev(u.data_ptr(), u.shape[0]); del u
sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
out = torch.empty(128, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
res = {}
for c in (16, 15):
    ctx.set_window(c)
    t0 = time.perf_counter()
    ctx.msm_table_dev(sc.data_ptr(), tab, 0, n, out.data_ptr())
    r = ctx.ext_sum_compress_dev(out.data_ptr(), 1)
    res[c] = (bytes(r).hex(), time.perf_counter() - t0)
    print(c, res[c], flush=True)
assert res[16][0] == res[15][0]
print(f"ok: n=2^{logn}, {n / res[16][1] / 1e6:.0f} M points/s single call, mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB (torch side only)")

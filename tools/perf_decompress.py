#!/usr/bin/env python3
"""Time table construction from compressed points (device pointers, CUDA events via torch)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkvm_b200 as zk
ctx = zk.Context(0)
n = 1 << 20
u = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
tab = zk.PointTable(ctx, n).append_uniform_dev(u.data_ptr(), n)
comp = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
tab.compress_dev(comp.data_ptr()); ctx.sync()
st = torch.cuda.ExternalStream(ctx.stream)
best = 1e9
t2 = zk.PointTable(ctx, n)
for i in range(6):
    t2.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); t2.append_compressed_dev(comp.data_ptr(), n); e1.record(st); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"lib": os.path.basename(os.environ.get("ZKMSM_LIB", "libzkmsm.so")), "decompress_2e20_ms": round(best, 4)}))

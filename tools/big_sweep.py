#!/usr/bin/env python3
"""Window sweep for block-scale MSMs (device-generated inputs, device-pointer API, CUDA-event timing)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zkvm_b200 as zk
ctx = zk.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
g = torch.Generator(device="cuda"); g.manual_seed(1)
for logn in [int(a) for a in sys.argv[1:]] or [22, 24]:
    n = 1 << logn
    tab = zk.PointTable(ctx, n)
    step = 1 << 22
    for lo in range(0, n, step):
        u = torch.randint(0, 256, (min(step, n - lo), 64), dtype=torch.uint8, device="cuda", generator=g)
        torch.cuda.synchronize(); tab.append_uniform_dev(u.data_ptr(), u.shape[0]); del u
    sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    out = torch.empty(128, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    row, ref = {}, None
    for c in (15, 16, 17, 18, 19, 20):
        ctx.set_window(c)
        best = 1e9
        for i in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); ctx.msm_table_dev(sc.data_ptr(), tab, 0, n, out.data_ptr()); e1.record(st)
            r = ctx.ext_sum_compress_dev(out.data_ptr(), 1); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        ref = ref or bytes(r); assert bytes(r) == ref
        row[c] = round(best, 3)
    print(json.dumps({"logn": logn, "ms": row, "best_c": min(row, key=row.get), "Mpts_s": round(n / min(row.values()) / 1e3)}), flush=True)
    del tab, sc

#!/usr/bin/env python
"""Time the device BBHash build / lookup / dump at a few key counts (not a bench line)."""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

g.build()
from spacegraphcats_b200.device import DeviceMphf, get_context  # noqa: E402

ctx = get_context(0)
for lg in [int(a) for a in sys.argv[1:]] or [24, 26, 29]:
    n = 1 << lg
    gen = torch.Generator(device="cuda").manual_seed(lg)
    keys = torch.unique(torch.randint(-(2**63), 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=gen))
    torch.cuda.synchronize()
    DeviceMphf.build(ctx, keys).close()  # warm-up: first-touch of fresh device memory
    ctx.sync()
    t0 = time.perf_counter()
    m = DeviceMphf.build(ctx, keys)
    ctx.sync()
    t1 = time.perf_counter()
    idx = m.lookup(keys, to_host=False)
    ctx.sync()
    t2 = time.perf_counter()
    raw = m.dumps()
    t3 = time.perf_counter()
    print(f"n=2^{lg} ({keys.numel()}): build {1e3*(t1-t0):.1f} ms ({keys.numel()/(t1-t0)/1e6:.0f} M keys/s), "
          f"lookup {1e3*(t2-t1):.1f} ms ({keys.numel()/(t2-t1)/1e9:.2f} G/s), dump {len(raw)/1e6:.1f} MB in {1e3*(t3-t2):.0f} ms, "
          f"final {m.info()['n_final']}", flush=True)
    m.close()
    del keys, idx
    torch.cuda.empty_cache()

if os.environ.get("MPHF_PROF"):
    from torch.profiler import ProfilerActivity, profile

    n = 1 << int(os.environ["MPHF_PROF"])
    gen = torch.Generator(device="cuda").manual_seed(1)
    keys = torch.unique(torch.randint(-(2**63), 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=gen))
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        m = DeviceMphf.build(ctx, keys)
        ctx.sync()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))

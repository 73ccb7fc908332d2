#!/usr/bin/env python
"""Warm, in-graph kernel timeline of the bench step through torch.profiler (CUPTI activity tracing -- guidance only, a
number taken under a profiler is never a bench value).  Prints per-kernel totals over a few graph replays and the gaps
between consecutive kernels on the main stream.
    python tools/step_profile.py [--batch 64] [--steps 5]"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gif-gan_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--timeline", action="store_true", help="list every kernel of the middle replay: start, duration, stream, gap")
    a = ap.parse_args()
    from gifgan import ops
    from gifgan.model import DCGAN
    ops.set_precision("bf16")
    ops.reset_default_store(device="cuda", seed=7)
    B = a.batch
    m = DCGAN(None, batch_size=B, output_size=64, c_dim=3)
    img = torch.from_numpy(np.random.RandomState(1).uniform(-1, 1, (B, 64, 64, 3)).astype(np.float32)).pin_memory()
    z = torch.from_numpy(np.random.RandomState(2).uniform(-1, 1, (B, 100)).astype(np.float32)).pin_memory()
    for _ in range(3):
        m.train_step(img, z, use_graph=True)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(a.steps):
            m._graph["graph"].replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0]
    evs.sort(key=lambda e: e.time_range.start)
    agg = collections.OrderedDict()
    for e in evs:
        k = e.name.split("(")[0][-60:]
        s = agg.setdefault(k, [0, 0.0])
        s[0] += 1
        s[1] += e.device_time_total
    tot = sum(v[1] for v in agg.values())
    span = (evs[-1].time_range.end - evs[0].time_range.start) if evs else 0
    print(f"kernels {len(evs)} over {a.steps} replays; sum of kernel time {tot / a.steps:.1f} us/step; wall span {span / a.steps:.1f} us/step")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{k:62s} {n / a.steps:6.1f}/step {t / a.steps:9.1f} us/step {t / n:8.2f} us avg")
    if a.timeline and evs:
        per = len(evs) // a.steps
        mid = evs[per * (a.steps // 2): per * (a.steps // 2 + 1)]
        t0 = mid[0].time_range.start
        last_end = {}
        busy_any = 0.0
        cur_end = t0
        print("# timeline of one replay: start_us dur_us gap_same_stream_us stream name")
        for e in mid:
            st = getattr(e, "device_index", 0), getattr(e, "stream", None) if hasattr(e, "stream") else None
            sid = getattr(e, "device_resource_id", None)
            b, en = e.time_range.start, e.time_range.end
            gap = b - last_end.get(sid, b)
            last_end[sid] = en
            if en > cur_end:
                busy_any += en - max(b, cur_end)
                cur_end = en
            print(f"{b - t0:9.1f} {en - b:7.1f} {gap:7.1f} {str(sid):>4s} {e.name.split('(')[0][-70:]}")
        print(f"# replay span {mid[-1].time_range.end - t0:.1f} us; time with >= 1 kernel running {busy_any:.1f} us")


if __name__ == "__main__":
    main()

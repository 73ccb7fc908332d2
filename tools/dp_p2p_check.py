#!/usr/bin/env python
"""gg_dp_allreduce (csrc/dp_allreduce.cu) on N GPUs of one box: correctness against an explicit rank-order sum of all-gathered
buffers (bit-exact, identical on every rank), ragged ranges, repeated launches, CUDA-graph replays; then its device time next to
ncclAllReduce (fp32 and bf16) for the gradient ranges of the DCGAN step.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_p2p_check.py [--time]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gif-gan_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from gifgan import _cabi as c
    from gifgan.dp import DataParallel
    dp = DataParallel()
    torch.cuda.set_device(dp.local_rank)
    W, R = dp.world_size, dp.rank
    n_total = 6 * 1024 * 1024 + 3
    grads = torch.empty(n_total, dtype=torch.float32, device="cuda")
    tabs = dp._peer_tables(grads)
    assert tabs, "peer-memory exchange not available"
    L = c.lib()

    def fill(seed):
        g = torch.Generator(device="cuda").manual_seed(1000 * seed + R)
        grads.copy_(torch.randn(n_total, device="cuda", generator=g))

    def expected(lo, n):
        parts = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(W)]
        dist.all_gather(parts, grads[lo:lo + n].clone())
        s = parts[0].clone()
        for p in parts[1:]:
            s += p                       # rank order, fp32: what the kernel computes
        return s.to(torch.bfloat16).float() if wire[0] else s

    wire = [0]

    def run(lo, n):
        c.check(L.gg_dp_allreduce(tabs["grads"], tabs["stage"], tabs["sig"], R, W, lo, n, wire[0], c.stream()), "gg_dp_allreduce")

    worst = 0
    cases = [(0, n_total), (4, 1), (8, 3), (1024, 4096), (4096, 1 << 20), (12, (1 << 20) + 1), (0, 5 * 1024 * 1024 + 2), (64, 127)]
    for k, (lo, n) in enumerate(cases + cases):
        wire[0] = 1 if k >= len(cases) else 0          # second pass: bf16 on the wire in the all-gather phase
        fill(k)
        before = grads.clone()
        want = expected(lo, n)
        run(lo, n)
        torch.cuda.synchronize()
        got = grads[lo:lo + n]
        assert torch.equal(got, want), (k, lo, n, (got - want).abs().max().item())
        # untouched outside the range
        assert torch.equal(grads[:lo], before[:lo]) and torch.equal(grads[lo + n:], before[lo + n:]), (k, "outside")
        worst += 1
        dist.barrier()
    # captured: three exchanges per replay (like a train step), replayed many times with fresh data
    fill(99)
    src = grads.clone()
    g = torch.cuda.CUDAGraph()
    ranges = [(0, 2 * 1024 * 1024), (2 * 1024 * 1024, 3 * 1024 * 1024 + 1), (5 * 1024 * 1024 + 4, 1024 * 1024 - 4)]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for lo, n in ranges:
            run(lo, n)
        torch.cuda.synchronize()
        grads.copy_(src)
        with torch.cuda.graph(g, stream=s):
            for lo, n in ranges:
                run(lo, n)
    torch.cuda.synchronize()
    for it in range(20):
        fill(200 + it)
        wants = [expected(lo, n) for lo, n in ranges]
        g.replay()
        torch.cuda.synchronize()
        for (lo, n), w in zip(ranges, wants):
            assert torch.equal(grads[lo:lo + n], w), ("graph", it, lo, n)
    dist.barrier()
    if R == 0:
        print("dp_p2p_check: %d ranks, %d eager cases + 20 graph replays x 3 ranges: bit-exact and identical on every rank" % (W, worst), flush=True)
    if "--time" in sys.argv:
        out = {}
        for name, n in (("D group 17.3 MB", 4325000), ("G group 20.5 MB", 5135000), ("1 MB", 262144)):
            res = {}
            buf16 = torch.empty(n, dtype=torch.bfloat16, device="cuda")
            for kind in ("p2p fp32", "p2p bf16 wire", "nccl fp32", "nccl bf16"):
                wire[0] = 1 if kind == "p2p bf16 wire" else 0

                def one():
                    if kind.startswith("p2p"):
                        run(0, n)
                    elif kind == "nccl fp32":
                        dist.all_reduce(grads[:n])
                    else:
                        dist.all_reduce(buf16)
                for _ in range(5):
                    one()
                torch.cuda.synchronize(); dist.barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(20):
                    one()
                b.record(); b.synchronize()
                t = torch.tensor([a.elapsed_time(b) / 20 * 1e3], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                res[kind] = round(t.item(), 1)
                if kind.startswith("p2p"):      # block 0 of the last launch: cycles since kernel entry (DpSignals::prof)
                    sig = tabs["keep"][0]
                    pr = sig[sig.numel() - 64:].view(torch.int64)[:4].tolist()
                    res[kind + " cycles to [barrier0, phase1, barrier1, phase2]"] = pr
            out[name] = res
        if R == 0:
            print("TIMING_US " + json.dumps(dict(world=W, us_per_allreduce=out)), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Kernel-time sensitivity sweep for the tcgen05 kernels: true device time per launch (plan built once, 20
back-to-back launches, CUDA events, warm L2) for full-size DCGAN layer shapes under GG_TC_STAGES / GG_TC_BN overrides.
Each configuration runs in a subprocess (the overrides are read from the environment at plan time)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gif-gan_b200")):
    sys.path.insert(0, p)

SHAPES = {   # name: N, H(large), C(large), K(small)
    "d_h1": (128, 32, 64, 128), "d_h2": (128, 16, 128, 256), "d_h3": (128, 8, 256, 512),
    "g_h1": (64, 8, 256, 512), "g_h2": (64, 16, 128, 256), "g_h3": (64, 32, 64, 128),
    "b256_h1": (256, 8, 256, 512), "b256_h2": (256, 16, 128, 256), "b256_h3": (256, 32, 64, 128),
}


def one(shape, op, reps=20):
    import torch
    from collections import OrderedDict
    from gifgan import ops, _cabi
    N, H, C, K = SHAPES[shape]
    g = ops._Geom(N, (1, H, H), C, (1, H // 2, H // 2), K, (1, 5, 5), (1, 2, 2), (0, 1, 1))
    ops.set_precision("bf16", tensor_cores=True)
    st = ops.reset_default_store(device="cuda")
    wv = st.get_variable("w", (5, 5, C, K), lambda r, s: __import__("numpy").zeros(s, dtype="float32") + 0.01, filter_taps=25)
    st.finalize(OrderedDict(all=[wv]))
    large = torch.randn(N, H, H, C, device="cuda").to(torch.bfloat16)
    small = torch.randn(N, H // 2, H // 2, K, device="cuda").to(torch.bfloat16)
    out_l, out_s = torch.empty_like(large), torch.empty_like(small)
    if os.environ.get("GG_SWEEP_STATS"):      # the in-step form of a layer that feeds a batch norm: fp32 pre-norm output + fused statistics
        out_l, out_s = out_l.float(), out_s.float()
        st_l = torch.zeros(2 * 2 * C, dtype=torch.float64, device="cuda")
        st_s = torch.zeros(2 * 2 * K, dtype=torch.float64, device="cuda")
        fn = {"down": lambda: ops._run_down(g, large, wv, None, torch.float32, None, 0.0, 4, out=out_s, stats=st_s, groups=2),
              "up": lambda: ops._run_up(g, small, wv, None, torch.float32, None, 0.0, 4, out=out_l, stats=st_l, groups=2),
              "wgrad": lambda: ops._run_wgrad(g, large, small, wv)}[op]
    elif os.environ.get("GG_SWEEP_BNB"):      # dgrad launch with the consumer batch norm's backward reductions fused into its epilogue
        def info(shape, Cc):
            b = ops._BnInfo()
            b.pre = torch.randn(shape, device="cuda")
            b.mean, b.rstd = torch.zeros(1, Cc, device="cuda"), torch.ones(1, Cc, device="cuda")
            b.gamma, b.beta = torch.ones(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
            b.act, b.act_param, b.groups, b.Cc = "lrelu", 0.2, 1, Cc
            return b
        bl, bs = info(large.shape, C), info(small.shape, K)
        fn = {"down": lambda: ops._run_down(g, large, wv, None, torch.bfloat16, None, 0.0, 4, out=out_s, bnb=bs),
              "up": lambda: ops._run_up(g, small, wv, None, torch.bfloat16, None, 0.0, 4, out=out_l, bnb=bl),
              "wgrad": lambda: ops._run_wgrad(g, large, small, wv)}[op]
    else:
      fn = {"down": lambda: ops._run_down(g, large, wv, None, torch.bfloat16, None, 0.0, 4, out=out_s),
          "up": lambda: ops._run_up(g, small, wv, None, torch.bfloat16, None, 0.0, 4, out=out_l),
          "wgrad": lambda: ops._run_wgrad(g, large, small, wv)}[op]
    fn(); torch.cuda.synchronize()
    _cabi.lib().gg_debug_set_repeat(reps)
    best = 1e9
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record(); fn(); e.record(); e.synchronize()
        best = min(best, s.elapsed_time(e) / reps)
    _cabi.lib().gg_debug_set_repeat(1)
    if op != "wgrad" and os.environ.get("GG_PROF"):
        buf = torch.zeros(512 * 8, dtype=torch.int64, device="cuda")
        _cabi.lib().gg_debug_set_prof(buf.data_ptr())
        fn(); torch.cuda.synchronize()
        _cabi.lib().gg_debug_set_prof(None)
        b = buf.cpu().reshape(512, 8).double()
        b = b[b[:, 6] > 0]
        names = ["prodA_loop", "xchg", "mma_loop", "mma_wait_full", "acc_ready", "epi_done", "cta_total", "setup"]
        raw = buf.cpu().reshape(512, 8)
        raw = raw[raw[:, 6] > 0][:, 1]
        xc = [(((raw >> (16 * k)) & 0xFFFF).double().mean().item() * 16) for k in range(4)]      # split-K through L2: exchange phases
        print("PROF ctas=%d " % len(b) + " ".join("%s=%.0f" % (n, b[:, i].mean().item()) for i, n in enumerate(names) if n != "xchg") +
              " cta_total_max=%.0f acc_ready_max=%.0f" % (b[:, 6].max().item(), b[:, 4].max().item()) +
              " xchg_stage=%.0f xchg_store=%.0f xchg_wait=%.0f xchg_load=%.0f" % tuple(xc), flush=True)
    import bench
    fl = bench.conv_flops_per_image(H, C, K) * N
    print("RESULT " + json.dumps(dict(shape=shape, op=op, us=round(best * 1e3, 2), tflops=round(fl / best / 1e9, 1),
                                      stages=os.environ.get("GG_TC_STAGES", ""), bn=os.environ.get("GG_TC_BN", ""))), flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 2 and sys.argv[1] == "--inproc":      # default plan only, every shape, one process
        for shape in SHAPES:
            for op in ("down", "up", "wgrad"):
                one(shape, op)
        sys.exit(0)
    if len(sys.argv) > 2:
        one(sys.argv[1], sys.argv[2])
        sys.exit(0)
    configs = [dict()] + [dict(GG_TC_STAGES=str(s)) for s in (2, 4)] + [dict(GG_TC_BN=str(b)) for b in (64, 128, 256)]
    if os.environ.get("GG_PROF"):
        configs = [dict()]
    default_shapes = list(SHAPES) if os.environ.get("GG_PROF") else ["d_h2", "g_h1", "g_h3"]
    for shape in (default_shapes if len(sys.argv) < 2 else [sys.argv[1]]):
        for op in ("down", "up", "wgrad"):
            for cfg in (configs if op != "wgrad" else configs[:3]):
                env = dict(os.environ); env.update(cfg)
                r = subprocess.run([sys.executable, os.path.abspath(__file__), shape, op], capture_output=True, text=True, env=env, timeout=120)
                for l in r.stdout.splitlines():
                    if l.startswith("PROF "):
                        print(l, flush=True)
                lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
                print(lines[-1][7:] if lines else json.dumps(dict(shape=shape, op=op, cfg=cfg, err=(r.stderr or r.stdout)[-300:])), flush=True)

#!/usr/bin/env python
"""Hardware probe for the tcgen05 kernels: runs conv_down / conv_up / conv_wgrad on the tensor-core path
for a list of shapes, each in its own subprocess (a faulting kernel leaves a sticky CUDA error), compares
with the SIMT fp32-accumulate kernels of the same library (validated against the oracle by tests/) and with
the CPU oracle, and prints one line per case.  Usage: python tools/tc_probe.py [--out gpurun_out/tc_probe.log]"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gif-gan_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

# name, N, (D,H,W), C(large), K(small), k, s
CASES = [
    ("c2d_16_64_64", 2, (1, 16, 16), 64, 64, (1, 5, 5), (1, 2, 2)),
    ("c2d_16_64_128", 3, (1, 16, 16), 64, 128, (1, 5, 5), (1, 2, 2)),
    ("d_h1", 8, (1, 32, 32), 64, 128, (1, 5, 5), (1, 2, 2)),
    ("d_h2", 8, (1, 16, 16), 128, 256, (1, 5, 5), (1, 2, 2)),
    ("d_h3", 16, (1, 8, 8), 256, 512, (1, 5, 5), (1, 2, 2)),
    ("odd_14x10", 3, (1, 14, 10), 64, 64, (1, 5, 5), (1, 2, 2)),
    ("c3d_h1", 2, (16, 8, 8), 256, 256, (3, 3, 3), (2, 2, 2)),
    ("c3d_h3", 4, (4, 2, 2), 256, 256, (3, 3, 3), (2, 2, 2)),
    ("d_h1_full", 128, (1, 32, 32), 64, 128, (1, 5, 5), (1, 2, 2)),
    ("d_h2_full", 128, (1, 16, 16), 128, 256, (1, 5, 5), (1, 2, 2)),
    ("g_h1_full", 64, (1, 8, 8), 256, 512, (1, 5, 5), (1, 2, 2)),
]
OPS = ["down", "up", "wgrad"]


def same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2


def run_case(case_idx, op):
    import numpy as np
    import torch
    from gifgan import ops
    from oracle import tf_ops as T
    name, N, (D, H, W), C, K, k, s = CASES[case_idx]
    Do, pd = same_pad(D, k[0], s[0]); Ho, ph = same_pad(H, k[1], s[1]); Wo, pw = same_pad(W, k[2], s[2])
    g = ops._Geom(N, (D, H, W), C, (Do, Ho, Wo), K, k, s, (pd, ph, pw))
    rs = np.random.RandomState(case_idx)
    ndim = 5 if D > 1 or k[0] > 1 else 4
    lshape, sshape = g.large_shape(ndim), g.small_shape(ndim)
    large = torch.tensor(rs.randn(*lshape), dtype=torch.float32, device="cuda").to(torch.bfloat16)
    small = torch.tensor(rs.randn(*sshape), dtype=torch.float32, device="cuda").to(torch.bfloat16)
    wshape = tuple(k[(3 - (ndim - 2)):]) + (C, K)
    w = (rs.randn(*wshape) * 0.05).astype(np.float32)
    w = torch.tensor(w).to(torch.bfloat16).float().numpy()     # bf16-representable weights: both paths see identical values
    ops.set_precision("bf16", tensor_cores=True)
    st = ops.reset_default_store(device="cuda")
    wv = st.get_variable("w", wshape, lambda r, sh: w, filter_taps=k[0] * k[1] * k[2])
    from collections import OrderedDict
    st.finalize(OrderedDict(all=[wv]))
    bias = torch.tensor(rs.randn(K if op == "down" else C) * 0.1, dtype=torch.float32, device="cuda")

    def run(tc):
        ops._USE_TC = tc
        if op == "down":
            return ops._run_down(g, large, wv, bias, torch.bfloat16, None, 0.0, ndim).float()
        if op == "up":
            return ops._run_up(g, small, wv, bias, torch.bfloat16, None, 0.0, ndim).float()
        wv.grad.zero_()
        ops._run_wgrad(g, large, small, wv)
        return wv.grad.clone()

    ref = run(False)
    torch.cuda.synchronize()
    got = run(True)
    torch.cuda.synchronize()
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    l2 = ((got - ref).norm() / ref.norm()).item()
    res = dict(case=name, op=op, max_rel=err, l2_rel=l2, ref_absmax=ref.abs().max().item(), got_absmax=got.abs().max().item(),
               nan=bool(torch.isnan(got).any()), shape=list(got.shape))
    if err > 2e-2:
        # diagnostics: where are the errors?
        e = (got - ref).abs() / ref.abs().max()
        flat = e.reshape(-1, e.shape[-1])
        res["bad_frac"] = (e > 2e-2).float().mean().item()
        res["err_by_col64"] = [round(flat[:, i:i + 64].max().item(), 3) for i in range(0, flat.shape[1], 64)][:8]
        res["err_by_row8"] = [round(flat[i::8].max().item(), 3) for i in range(8)]
        res["first_rows_err"] = [round(flat[i].max().item(), 3) for i in range(min(16, flat.shape[0]))]
    # timing
    for _ in range(3):
        run(True)
    torch.cuda.synchronize()
    s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_.record()
    for _ in range(10):
        run(True)
    e_.record(); e_.synchronize()
    res["tc_ms"] = s_.elapsed_time(e_) / 10
    s_.record()
    for _ in range(10):
        run(False)
    e_.record(); e_.synchronize()
    res["simt_ms"] = s_.elapsed_time(e_) / 10
    print("RESULT " + json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=-1)
    ap.add_argument("--op", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "tc_probe.log"))
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    if a.case >= 0:
        run_case(a.case, a.op)
        return
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        for i, c in enumerate(CASES):
            if a.only and a.only not in c[0]:
                continue
            for op in OPS:
                try:
                    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", str(i), "--op", op], capture_output=True,
                                       text=True, timeout=120)
                    lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
                    msg = lines[-1][7:] if lines else json.dumps(dict(case=c[0], op=op, rc=r.returncode, err=(r.stderr or r.stdout)[-600:]))
                except subprocess.TimeoutExpired:
                    msg = json.dumps(dict(case=c[0], op=op, err="timeout"))
                f.write(msg + "\n"); f.flush()
                print(msg, flush=True)


if __name__ == "__main__":
    main()

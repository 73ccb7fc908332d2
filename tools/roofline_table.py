#!/usr/bin/env python
"""Per-kernel roofline table of the conv layers from a layer-kernel log (tools/layer_kernels.py output, one JSON line
per kernel: device time timed alone with CUDA events, exact valid-tap FLOPs):
    python tools/roofline_table.py profiles/r02g_layer_kernels.log > profiles/r02g_roofline_table.md
Columns: time per launch, launches per train step, achieved TFLOP/s on the exact FLOPs, fraction of the measured bf16
burst peak (MEASURED_PEAKS.json, 1644.8 TFLOP/s fallback), the time the launch would take AT that peak, and the step
time that kernel class would give back if it ran at 50 % of peak (north_star's target)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path = sys.argv[1]
    peak = 1644.8
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        try:
            d = json.load(open(pk))
            peak = float(d.get("bf16_tflops_burst", d.get("tf_burst", peak)))
        except Exception:
            pass
    rows = [json.loads(l) for l in open(path) if l.startswith("{")]
    rows.sort(key=lambda r: -r["share_ms"])
    tot = sum(r["share_ms"] for r in rows)
    print("# Conv-layer kernels of one DCGAN-64 train step (batch 64): measured vs the tensor-core roofline\n")
    print("Source: `%s`; peak = %.1f TFLOP/s (measured bf16 burst).  `in step` = time per launch x launches per step.\n" % (os.path.basename(path), peak))
    print("| kernel | path | us / launch | launches | in step (us) | TFLOP/s | of peak | us at peak | saved at 50 % of peak (us) |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|")
    saved = 0.0
    for r in rows:
        us = r["ms"] * 1e3
        at_peak = r["flops"] / (peak * 1e12) * 1e6
        gain = max(0.0, us - 2 * at_peak) * r["uses"] if r["path"] == "tcgen05" else 0.0
        saved += gain
        print("| `%s` | %s | %.1f | %d | %.1f | %.0f | %.3f | %.2f | %.1f |" % (
            r["kernel"], r["path"], us, r["uses"], r["share_ms"] * 1e3, r["tflops"], r["tflops"] / peak, at_peak, gain))
    print("\nconv kernels in the step: %.0f us; the tcgen05 kernels at 50 %% of peak would give back %.0f us." % (tot * 1e3, saved))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- GAN train frames/sec for gif-gan's conv-GAN training step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]

A "step" is one iteration of the reference's train loop body (models/recurrent_z/model.py:226-239):
1 discriminator update + 2 generator updates on one batch of synthetic 64x64x3 frames (BASELINE.json
config 2: per-frame DCGAN, batch 64 per GPU; weak scaling for N > 1).  `value` = frames/s with the batch
already resident in HBM (CUDA-graph replay); `e2e` = the same through the public API
`DCGAN.train_step(host_images, host_z)` with the host->device copies and the device->host loss read
inside the timed region.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gif-gan_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "GAN train frames/sec"
UNIT = "frames/s"


# ---------------------------------------------------------------------------------------------
# work accounting: exact valid-tap FLOPs
# ---------------------------------------------------------------------------------------------
def valid_taps(n, k=5, s=2):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    lo = total // 2
    return sum(1 for o in range(out) for t in range(k) if 0 <= s * o + t - lo < n)


def dcgan_layers(size=64, c=3, gf=64, df=64, z_dim=100):
    """(name, large_hw, C_large, K_small) for the eight 5x5 stride-2 layers + the two linears."""
    convs = []
    chans = [c, df, 2 * df, 4 * df, 8 * df]
    hw = size
    for i in range(4):
        convs.append((f"d_h{i}_conv", hw, chans[i], chans[i + 1]))
        hw //= 2
    g = [8 * gf, 4 * gf, 2 * gf, gf, c]
    hw = size // 16
    for i in range(4):
        hw *= 2
        convs.append((f"g_h{i + 1}", hw, g[i + 1], g[i]))
    return convs


def conv_flops_per_image(hw, C, K):
    v = valid_taps(hw)
    return 2.0 * v * v * C * K


def step_flops_per_image(size=64, c=3, gf=64, df=64, z_dim=100):
    """1 D-update + 2 G-updates, GEMM FLOPs only (SURVEY App. B accounting)."""
    L = {n: conv_flops_per_image(hw, C, K) for n, hw, C, K in dcgan_layers(size, c, gf, df)}
    s16 = size // 16
    lin_g = 2.0 * z_dim * 8 * gf * s16 * s16
    lin_d = 2.0 * 8 * df * s16 * s16
    fwd_d = sum(L[f"d_h{i}_conv"] for i in range(4)) + lin_d
    fwd_g = sum(L[f"g_h{i}"] for i in range(1, 5)) + lin_g
    d_dgrad_1_3 = sum(L[f"d_h{i}_conv"] for i in range(1, 4)) + lin_d
    d_upd = fwd_g + 2 * fwd_d + 2 * (fwd_d + d_dgrad_1_3)            # wgrad all layers + dgrad layers 1-3, both halves
    g_upd = fwd_g + fwd_d + (fwd_d) + fwd_g + (fwd_g - lin_g)        # D dgrad (all), G wgrad (all), G dgrad (deconvs)
    return d_upd + 2 * g_upd


# ---------------------------------------------------------------------------------------------
def ncu_by_layer():
    """profiles/*_ncu_by_layer.json (tools/summarize_profiles.py): per-launch DRAM traffic and tensor-pipe activity of each
    layer kernel from the committed `ncu --set full` capture of tools/layer_kernels.py (newest file wins)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_by_layer.json")), key=os.path.getmtime)
    if not files:
        return {}, None
    with open(files[-1]) as f:
        return json.load(f), os.path.basename(files[-1])


def workload_name(batch):
    return f"DCGAN 64x64x3 per-frame GAN (BASELINE config 2: models/recurrent_z model.py), batch {batch}/GPU, 1 D + 2 G updates per step"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def cpu_reference_step_rate(batch, steps, warmup):
    """The reference's own CPU path is TensorFlow 0.12 (not installable: DESIGN.md); the timed stand-in is
    the oracle's restatement of the same graph and schedule in PyTorch-CPU fp32 on all host cores."""
    from oracle.models import DCGAN as OracleDCGAN
    torch.set_num_threads(os.cpu_count() or 1)
    ora = OracleDCGAN(batch_size=batch, seed=7)
    img = torch.tensor(np.random.RandomState(102).uniform(-1, 1, (batch, 64, 64, 3)).astype(np.float32))
    times = []
    for i in range(warmup + steps):
        z = torch.tensor(np.random.RandomState(1000 + i).uniform(-1, 1, (batch, 100)).astype(np.float32))
        t0 = time.perf_counter()
        ora.train_step(img, z)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return batch / sec, sec, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = min(args.steps, 10), min(args.warmup, 2)
    fps, sec, cores = cpu_reference_step_rate(args.batch, steps, max(1, warmup))
    sample = f"{steps} full steps (1 D + 2 G updates) of batch {args.batch} after {max(1, warmup)} warm-up, torch-CPU fp32 oneDNN"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": max(1, warmup),
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch), "global_batch": args.batch, "parallelism": "cpu",
                   "note": "reference TensorFlow-0.12 cannot run here; CPU restatement of the same graph (oracle/) on host cores"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def time_kernel(fn, flush, reps=5, launches=10):
    """Average device time (ms) of ONE launch of the kernel behind fn(), CUDA events on the launching stream.
    The host-side planning of a conv call (tensor-map encoding, ctypes) costs more than the kernel itself at these
    sizes, so each fn() enqueues `launches` back-to-back launches (gg_debug_set_repeat: plan once, launch n times);
    L2 is flushed before every timed batch (operands of one layer, 4-34 MB, then stay L2-resident across the batch,
    as they are inside the real step where the producer kernel has just written them)."""
    from gifgan import _cabi
    fn()
    torch.cuda.synchronize()
    _cabi.lib().gg_debug_set_repeat(launches)
    tot = 0.0
    try:
        for _ in range(reps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            e.synchronize()
            tot += s.elapsed_time(e)
    finally:
        _cabi.lib().gg_debug_set_repeat(1)
    return tot / reps / launches


def layer_rooflines(model, batch, precision, flush, reps=5, launches=10):
    """Time each conv layer's three kernels in isolation at the step's shapes; returns rows sorted by their
    share of the step (time x launches per step)."""
    import ctypes
    from gifgan import ops
    from gifgan._cabi import lib, check, ptr, stream, dt
    L = lib()
    dtp = ops.act_dtype()
    rows = []
    for name, hw, C, K in dcgan_layers():
        wvar = model.store.vars[name + "/w"]
        is_d = name.startswith("d_")
        for B_, uses in ((2 * batch, {"down": 1, "up": 1, "wgrad": 1}), (batch, {"down": 2, "up": 2, "wgrad": 0})) if is_d else \
                ((batch, {"up": 3, "down": 2, "wgrad": 2}),):
            ldt = torch.float32 if C == 3 else dtp
            large = torch.randn(B_, hw, hw, C, device="cuda").to(ldt)
            small = torch.randn(B_, hw // 2, hw // 2, K, device="cuda").to(dtp)
            g = ops._Geom(B_, (1, hw, hw), C, (1, hw // 2, hw // 2), K, (1, 5, 5), (1, 2, 2), (0, 1, 1))
            fl = conv_flops_per_image(hw, C, K) * B_
            fns = {
                "down": lambda: ops._run_down(g, large, wvar, None, small.dtype, None, 0.0, 4, out=small),
                "up": lambda: ops._run_up(g, small, wvar, None, large.dtype, None, 0.0, 4, out=large),
                "wgrad": lambda: ops._run_wgrad(g, large, small, wvar),
            }
            for kind, n_use in uses.items():
                if n_use == 0:
                    continue
                if is_d and kind == "up" and C == 3 and B_ == 2 * batch:
                    continue   # d_h0 dgrad is not needed in the D update
                ms = time_kernel(fns[kind], flush, reps, launches)
                tc = ops._tc_ok(C, K, large, small)
                path = "tcgen05" if tc else ("mma.sync" if (C == 3 and K % 64 == 0 and precision == "bf16") else "simt")
                rows.append(dict(kernel=f"{name}.{kind}[B={B_}]", ms=ms, flops=fl, tflops=fl / ms / 1e9, uses=n_use,
                                 path=path, share_ms=ms * n_use, order=len(rows)))
    rows.sort(key=lambda r: -r["share_ms"])
    return rows


def run_ours(args):
    from gifgan import ops
    from gifgan.dp import DataParallel
    from gifgan.model import DCGAN

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dp = DataParallel() if world > 1 else None
    rank = dp.rank if dp else 0
    local_rank = dp.local_rank if dp else 0
    torch.cuda.set_device(local_rank)
    peaks = load_peaks()

    ops.set_precision(args.precision, tensor_cores=None if args.tensor_cores < 0 else bool(args.tensor_cores))
    ops.reset_default_store(device="cuda", seed=7)
    B = args.batch
    model = DCGAN(None, batch_size=B, output_size=64, c_dim=3, dp=dp)
    if dp:
        dp.broadcast_parameters(model.store)

    n_batches = 4
    rs = np.random.RandomState(102 + rank)
    host_imgs = [torch.from_numpy(rs.uniform(-1, 1, (B, 64, 64, 3)).astype(np.float32)).pin_memory() for _ in range(n_batches)]
    host_z = [torch.from_numpy(np.random.RandomState(1000 + 17 * rank + i).uniform(-1, 1, (B, 100)).astype(np.float32)).pin_memory()
              for i in range(args.warmup + args.steps)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    # ---- warm-up (captures the CUDA graph on the first call) ---------------------------------
    for i in range(args.warmup):
        model.train_step(host_imgs[i % n_batches], host_z[i], use_graph=not args.eager)
    torch.cuda.synchronize()
    st, graph = model._static, model._graph

    # ---- value: batch resident in HBM, graph replay, per-step CUDA events, L2 flushed between steps
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if dp:
        dp.barrier()
    torch.cuda.synchronize()
    dev_ms = 0.0
    for i in range(args.steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        if args.eager:
            model._step_device(st["both"][:B], st["z"], None, False, st["loss_dev"])
        else:
            graph["graph"].replay()
            model.d_optim.t += 1; model.g_optim.t += 2
        e.record()
        e.synchronize()
        dev_ms += s.elapsed_time(e)
    torch.cuda.synchronize()
    if dp:
        dp.barrier()
    dev_ms = dp.max_over_ranks(dev_ms) if dp else dev_ms

    # ---- e2e: public API with host buffers (H2D of images + z, D2H of the losses) every step
    if dp:
        dp.barrier()
    torch.cuda.synchronize()
    e2e_ms = 0.0
    last = None
    for i in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        last = model.train_step(host_imgs[i % n_batches], host_z[args.warmup + i], use_graph=not args.eager)   # syncs on the loss read
        e2e_ms += (time.perf_counter() - t0) * 1e3
    if dp:
        dp.barrier()
    e2e_ms = dp.max_over_ranks(e2e_ms) if dp else e2e_ms
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        return
    frames = B * world
    value = frames * args.steps / (dev_ms / 1e3)
    e2e = frames * args.steps / (e2e_ms / 1e3)
    launches_per_step = (graph["launches"] if graph else 0)
    h2d = host_imgs[0].numel() * 4 + host_z[0].numel() * 4
    d2h = st["loss_host"].numel() * 4

    # ---- roofline of the dominant kernel (timed alone, burst peak) + whole-step tensor fraction
    rows = layer_rooflines(model, B, args.precision, flush)
    top = rows[0]
    step_fl = step_flops_per_image() * B
    ncu, ncu_file = ncu_by_layer()
    prof = ncu.get(top["kernel"], {})
    traffic = (prof["dram_read_bytes"] + prof["dram_write_bytes"]) if prof else None
    roofline = {"bound": "tensor", "kernel": top["kernel"], "path": top["path"], "achieved": top["tflops"], "peak": peaks["tf_burst"],
                "unit": "TFLOP/s", "frac": top["tflops"] / peaks["tf_burst"], "traffic": traffic, "peak_source": peaks["source"],
                "ncu": {"file": ncu_file, "tensor_pipe_active_pct": prof.get("tensor_pct"), "duration_us": prof.get("dur_us")} if prof else None,
                "flops_per_launch": top["flops"], "launch_ms": top["ms"],
                "step": {"gemm_tflop_per_step": step_fl / 1e12, "achieved_tflops": step_fl / (dev_ms / args.steps) / 1e9,
                         "frac_of_sustained": step_fl / (dev_ms / args.steps) / 1e9 / peaks["tf_sustained"]},
                "layers": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()} for r in rows[:12]]}

    # ---- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        fps, sec, cores = cpu_reference_step_rate(B, 3, 1)
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"3 full steps of batch {B} after 1 warm-up ({sec:.2f} s/step), oracle restatement (torch-CPU fp32), not TensorFlow"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": workload_name(B),
                   "global_batch": frames, "parallelism": f"dp{world}", "l2": "flushed (256 MB memset) before every timed step",
                   "bn": "per-replica statistics", "graph": not args.eager, "tensor_cores": bool(ops._USE_TC)},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches_per_step * args.steps * 2 + launches_per_step * args.warmup,
        "gpu_launches_per_step": launches_per_step,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "losses": last,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("GIFGAN_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--tensor-cores", type=int, default=-1, help="-1: library default, 0/1: force")
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU")
    ap.add_argument("--eager", action="store_true", help="no CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        run_ours(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

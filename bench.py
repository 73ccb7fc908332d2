#!/usr/bin/env python
"""bench.py -- GAN train frames/sec for gif-gan's conv-GAN training step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload dcgan|dcgan128|vid|mnist|recurrent]
                    [--precision bf16|fp32] [--batch B] [--clips C] [--no-extra] [--no-cpu-baseline]

A "step" is one iteration of the reference's train loop body: 1 discriminator update + 2 generator updates on one batch of
synthetic frames (models/recurrent_z/model.py:226-239, z_model_lib.py:217-239, rnn_test/recurrent_DCGAN.py:353-375).  The
three forward-only loss evaluations the reference loop also runs for its log line (model.py:241-243) are NOT part of the
timed step (optional), on either arm.

The headline line (default, every N) is BASELINE.json config 2: per-frame DCGAN 64x64x3, batch 64 per GPU (weak scaling).
`value` = frames/s with the batch resident in HBM (CUDA-graph replay, CUDA events, L2 flushed before every step, median of
--repeats repetitions of K steps, max over ranks); `e2e` = the same through the public API `DCGAN.train_step(host_images,
host_z)` with the host->device copies and the device->host loss read inside the timed region.  With N = 1 the other
configurations of BASELINE.json are measured in the same run and reported under `extra` (config 1 MNIST, config 3 video GAN
at 32 / 64 / 128 clips, the recurrent_image LSTM GAN, config 5's 128-px point); with N > 1, `extra` carries config 4: the
video GAN at a GLOBAL batch of 256 clips sharded over the ranks.  One JSON line on stdout (rank 0).
"""
import argparse
import hashlib
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
EVALS_NOTE = "1 D + 2 G updates per step; the reference loop's three forward-only loss evaluations (model.py:241-243) are excluded on both arms"


# ---------------------------------------------------------------------------------------------
# work accounting: exact valid-tap FLOPs
# ---------------------------------------------------------------------------------------------
def valid_taps(n, k=5, s=2):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    lo = total // 2
    return sum(1 for o in range(out) for t in range(k) if 0 <= s * o + t - lo < n)


def dcgan_layers(size=64, c=3, gf=64, df=64, z_dim=100):
    """(name, large_hw, C_large, K_small) for the eight 5x5 stride-2 layers."""
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


# exact GEMM FLOPs per step of the other configurations (per clip / per image)
VID_FLOPS_PER_CLIP = 44.19e9          # recurrent_z, 16 frames, image GAN frozen
REC_FLOPS_PER_CLIP = 105.1e9          # recurrent_image recurrent_DCGAN.py, 16 frames
MNIST_FLOPS_PER_IMAGE = 0.04e12 / 64  # config 1, ~0.04 TFLOP per step of 64


# the files that DEFINE each kernel family (kernel + every header it includes): a capture stays valid while they are unchanged
KERNEL_FAMILY_SOURCES = {"tcgen05": ("tc_tapgemm.cu", "tc_common.cuh", "common.cuh"), "mma.sync": ("conv_c3_mma.cu", "common.cuh")}


def kernel_source_sha(path=None, csrc=None):
    """Identity of the kernel sources: a committed ncu capture is only quoted when it was taken from THESE sources.
    path = "tcgen05" | "mma.sync": the files that define that kernel family; None: every .cu / .cuh of the library."""
    h = hashlib.sha256()
    d = csrc or os.path.join(ROOT, "gif-gan_b200", "csrc")
    files = KERNEL_FAMILY_SOURCES[path] if path else [f for f in sorted(os.listdir(d)) if f.endswith((".cu", ".cuh"))]
    for f in sorted(files):
        with open(os.path.join(d, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def ncu_by_layer(family="tcgen05"):
    """profiles/*_ncu_by_layer.json (tools/summarize_profiles.py): per-launch DRAM traffic and tensor-pipe activity of each layer
    kernel from a committed `ncu --set full` capture of tools/layer_kernels.py.  Only a capture whose recorded source identity
    of the kernel's family (`_kernel_source_sha_by_family`: the files that define the kernel, hashed from the commit the capture
    ran on) equals the current sources' is used -- otherwise `traffic` is null (never a stale number)."""
    import glob
    sha = kernel_source_sha(family)
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_by_layer.json")), reverse=True):
        with open(path) as f:
            d = json.load(f)
        if d.get("_kernel_source_sha_by_family", {}).get(family) == sha or d.get("_kernel_source_sha") == kernel_source_sha():
            return d, os.path.basename(path)
    return {}, None


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
# workloads: each returns an object with .frames (per rank per step), .flops (per rank per step), .replay(), .api_step(i), .launches
# ---------------------------------------------------------------------------------------------
class Workload:
    name = ""
    frames = 0
    flops = 0.0
    launches = 0
    h2d = 0
    d2h = 0

    def replay(self):          # one step, inputs resident, no host sync
        raise NotImplementedError

    def api_step(self, i):     # one step through the public API with host buffers; must sync on the result
        raise NotImplementedError


class DcganWorkload(Workload):
    """BASELINE config 2 (size 64) / config 5 point (size 128) / config 1 (mnist=True): gifgan.model.DCGAN."""

    def __init__(self, args, dp, rank, B, size=64, mnist=False, n_host=4, total_steps=32):
        from gifgan import ops
        from gifgan.model import DCGAN
        ops.reset_default_store(device="cuda", seed=7)
        self.B, self.size, self.mnist = B, size, mnist
        if mnist:
            self.model = DCGAN(None, batch_size=B, output_size=28, c_dim=1, y_dim=10, dataset_name="mnist", dp=dp)
            c, s = 1, 28
            self.name = f"tutorials/mnist conditional DCGAN 28x28x1 (BASELINE config 1), batch {B}/GPU, {EVALS_NOTE}"
            self.flops = MNIST_FLOPS_PER_IMAGE * B
        else:
            self.model = DCGAN(None, batch_size=B, output_size=size, c_dim=3, dp=dp)
            c, s = 3, size
            cfg = "BASELINE config 2" if size == 64 else "BASELINE config 5 point"
            self.name = f"DCGAN {size}x{size}x3 per-frame GAN ({cfg}: models/recurrent_z model.py), batch {B}/GPU, {EVALS_NOTE}"
            self.flops = step_flops_per_image(size) * B
        if dp:
            dp.broadcast_parameters(self.model.store)
        self.frames = B
        rs = np.random.RandomState(102 + rank)
        lo = 0.0 if mnist else -1.0
        self.host_imgs = [torch.from_numpy(rs.uniform(lo, 1, (B, s, s, c)).astype(np.float32)).pin_memory() for _ in range(n_host)]
        self.host_z = [torch.from_numpy(np.random.RandomState(1000 + 17 * rank + i).uniform(-1, 1, (B, 100)).astype(np.float32)).pin_memory()
                       for i in range(total_steps)]
        self.host_y = torch.from_numpy(np.eye(10, dtype=np.float32)[np.arange(B) % 10]).pin_memory() if mnist else None
        self.eager = args.eager
        self.h2d = self.host_imgs[0].numel() * 4 + self.host_z[0].numel() * 4 + (self.host_y.numel() * 4 if mnist else 0)
        self.i = 0

    def warm(self, n):
        for i in range(n):
            self.api_step(i)
        torch.cuda.synchronize()
        self.st, self.graph = self.model._static, self.model._graph
        self.launches = self.graph["launches"] if self.graph else 0
        self.d2h = self.st["loss_host"].numel() * 4

    def replay(self):
        m = self.model
        if self.eager:
            m._step_device(self.st["both"][:self.B], self.st["z"], self.st.get("y"), False, self.st["loss_dev"])
        else:
            self.graph["graph"].replay()
            m.d_optim.t += 1; m.g_optim.t += 2

    def api_step(self, i):
        return self.model.train_step(self.host_imgs[i % len(self.host_imgs)], self.host_z[i % len(self.host_z)], self.host_y,
                                     use_graph=not self.eager)


class VidWorkload(Workload):
    """BASELINE configs 3 / 4: gifgan.z_model_lib.VID_DCGAN, `clips` clips x 16 frames per rank, image GAN frozen."""

    def __init__(self, args, dp, rank, clips, T=16, what="BASELINE config 3"):
        from gifgan import ops
        from gifgan.z_model_lib import VID_DCGAN
        ops.reset_default_store(device="cuda", seed=7)
        with ops.variable_scope("video_gan"):
            self.model = VID_DCGAN(None, clips, 120, 100, T, 64, 64, 3, dp=dp)
        if dp:
            dp.broadcast_parameters(self.model.store)
        # stand-in for the loaded image-GAN checkpoint: seeded weights, non-trivial EMAs so that inference batch norm is exercised
        rs = np.random.RandomState(9)
        for k, v in self.model.store.vars.items():
            if k.endswith("moving_mean"):
                v.data.copy_(torch.tensor(rs.normal(0, 0.1, v.shape).astype(np.float32)))
            elif k.endswith("moving_variance"):
                v.data.copy_(torch.tensor(rs.uniform(0.5, 1.5, v.shape).astype(np.float32)))
        self.clips, self.T = clips, T
        self.frames = clips * T
        self.flops = VID_FLOPS_PER_CLIP * clips
        self.name = f"recurrent_z VID_DCGAN latent-MLP video GAN ({what}), {clips} clips x {T} frames 64x64x3 per GPU, image GAN frozen, {EVALS_NOTE}"
        self.img = torch.from_numpy(np.random.RandomState(103 + rank).uniform(-1, 1, (clips * T, 64, 64, 3)).astype(np.float32)).pin_memory()
        self.z = torch.from_numpy(np.random.RandomState(1000 + rank).uniform(-1, 1, (clips, 120)).astype(np.float32)).pin_memory()
        self.h2d = self.img.numel() * 4 + self.z.numel() * 4
        self.d2h = 7 * 4

    def warm(self, n):
        for _ in range(n):
            self.model.train_step(self.img, self.z)
        torch.cuda.synchronize()
        self.graph = self.model._graphs[(1, 2)]
        self.launches = self.graph["launches"]

    def replay(self):
        self.graph["graph"].replay()
        self.model.d_optim.t += 1; self.model.g_optim.t += 2

    def api_step(self, i):
        return self.model.train_step(self.img, self.z)


class RecurrentWorkload(Workload):
    """models/recurrent_image/rnn_test/recurrent_DCGAN.py: conv encoder -> LSTM -> deconv decoder, frame + clip discriminator."""

    def __init__(self, args, dp, rank, clips=40, T=16):
        from gifgan import ops
        from gifgan.recurrent_dcgan import RecurrentDCGAN
        ops.reset_default_store(device="cuda", seed=7)
        self.model = RecurrentDCGAN(batch_size=clips, video_length=T)
        self.frames = clips * T
        self.flops = REC_FLOPS_PER_CLIP * clips
        self.name = f"recurrent_image recurrent_DCGAN.py LSTM GAN, {clips} clips x {T} frames 64x64x3 per GPU (the script's batch_size 40), d_optim + 2 x g_optim per step, eager launches"
        self.inp = torch.from_numpy(np.random.RandomState(104 + rank).randint(0, 256, (clips, T + 1, 64, 64, 3)).astype(np.uint8)).pin_memory()
        self.dev_inp = None
        self.h2d = self.inp.numel()
        self.d2h = 8

    def warm(self, n):
        from gifgan import ops
        n0 = ops.cabi.launch_count()
        for _ in range(n):
            self.model.train_step(self.inp)
        torch.cuda.synchronize()
        self.launches = (ops.cabi.launch_count() - n0) // max(n, 1)
        self.dev_inp = self.inp.cuda()

    def replay(self):
        m = self.model
        m.update(self.dev_inp, "d"); m.update(self.dev_inp, "g"); m.update(self.dev_inp, "g")

    def api_step(self, i):
        return self.model.train_step(self.inp)


# ---------------------------------------------------------------------------------------------
def flush_l2(flush):
    """Evict the 126 MB L2: write a 256 MB buffer, then read it back.  The write alone would leave the cache full of DIRTY
    lines, and the measured kernel would then pay for their write-back on top of its own traffic (a 25 MB streaming kernel
    read 2x slower that way); after the read pass the cache holds clean lines of the flush buffer only."""
    flush.zero_()
    return flush.view(torch.int64).sum()


def timed(wl, steps, repeats, flush, dp, api=False):
    """`repeats` repetitions of `steps` steps.  Device arm: per-step CUDA events on the launching stream, L2 flushed (256 MB
    written and read back, outside the events) before every step.  API arm: wall clock around the public call (which synchronises on its
    device->host read), flush + synchronize before each call.  Returns the per-repetition totals in ms (max over ranks)."""
    totals = []
    for r in range(repeats):
        if dp:
            dp.barrier()
        torch.cuda.synchronize()
        ms = 0.0
        for i in range(steps):
            flush_l2(flush)
            if api:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                wl.api_step(r * steps + i)
                ms += (time.perf_counter() - t0) * 1e3
            else:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                wl.replay()
                e.record()
                e.synchronize()
                ms += s.elapsed_time(e)
        torch.cuda.synchronize()
        if dp:
            dp.barrier()
        totals.append(dp.max_over_ranks(ms) if dp else ms)
    return totals


def measure(wl, args, flush, dp, world, peaks, e2e=True):
    wl.warm(args.warmup)
    dev = timed(wl, args.steps, args.repeats, flush, dp)
    med = statistics.median(dev)
    out = {"workload": wl.name, "ms_per_step": med / args.steps, "value": wl.frames * world * args.steps / (med / 1e3), "unit": UNIT,
           "repeats_ms_per_step": [round(t / args.steps, 4) for t in dev], "gpu_launches_per_step": wl.launches,
           "gemm_tflop_per_step": wl.flops * world / 1e12,
           "achieved_tflops_per_gpu": wl.flops / (med / args.steps) / 1e9,
           "frac_of_sustained_bf16_peak": wl.flops / (med / args.steps) / 1e9 / peaks["tf_sustained"]}
    if e2e:
        api = timed(wl, args.steps, max(1, min(args.repeats, 3)), flush, dp, api=True)
        am = statistics.median(api)
        out["e2e"] = {"value": wl.frames * world * args.steps / (am / 1e3), "unit": UNIT, "h2d_bytes_per_step": wl.h2d, "d2h_bytes_per_step": wl.d2h,
                      "ms_per_step": am / args.steps}
    return out


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


def workload_name(batch):
    return f"DCGAN 64x64x3 per-frame GAN (BASELINE config 2: models/recurrent_z model.py), batch {batch}/GPU, {EVALS_NOTE}"


def run_reference(args):
    """The reference arm: the CPU restatement of the reference graph (oracle/, kind "port" -- TensorFlow 0.12 cannot be
    installed here) on the box's host cores, EXACTLY --steps timed steps after --warmup warm-ups of the same workload
    (one full 1 D + 2 G step of batch 64 is ~0.35 s on 16 cores, so the bounded sample is the whole step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    fps, sec, cores = cpu_reference_step_rate(args.batch, steps, warmup)
    sample = f"{steps} full steps (1 D + 2 G updates) of batch {args.batch} after {warmup} warm-up, torch-CPU fp32 oneDNN, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch), "global_batch": args.batch * max(1, args.gpus), "parallelism": f"dp{max(1, args.gpus)}",
                   "note": "reference TensorFlow-0.12 cannot run here; CPU restatement of the same graph (oracle/) on host cores, rank 0 only, one replica"},
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
            flush_l2(flush)
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
    from gifgan import ops
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
                try:        # the same launch COLD: one launch per L2 flush (operands from HBM), the pessimistic reading beside `ms`
                    ms_cold = time_kernel(fns[kind], flush, 3, 1)
                except Exception:       # noqa: BLE001
                    ms_cold = None
                tc = ops._tc_ok(C, K, large, small)
                path = "tcgen05" if tc else ("mma.sync" if (C == 3 and K % 64 == 0 and precision == "bf16") else "simt")
                # algorithmic bytes of the image-side (HBM-bound) layers: the fp32 image + the bf16 activation, each touched once
                nbytes = (B_ * hw * hw * C * 4 + B_ * (hw // 2) ** 2 * K * 2) if C == 3 else None
                rows.append(dict(kernel=f"{name}.{kind}[B={B_}]", ms=ms, flops=fl, tflops=fl / ms / 1e9, uses=n_use,
                                 path=path, share_ms=ms * n_use, order=len(rows), bytes=nbytes, ms_cold=ms_cold))
    rows.sort(key=lambda r: -r["share_ms"])
    return rows


def hbm_kernel_table(model, batch, flush, peaks, conv_rows):
    """Achieved HBM GB/s of the bandwidth-bound kernels of the step (north_star: norm / loss / optimizer kernels against the
    HBM roofline), each timed ALONE with CUDA events, L2 flushed (256 MB memset) before EVERY launch so that the operands come
    from HBM; algorithmic bytes = every operand touched once.  A single launch after a flush carries ~2 us of
    launch latency inside the events: the small kernels read low for that reason."""
    import ctypes
    from gifgan import ops
    from gifgan._cabi import lib, check, ptr, stream
    L = lib()
    rows = []

    def once(fn, reps=5):
        fn(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush_l2(flush)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); e.synchronize()
            tot += s.elapsed_time(e)
        return tot / reps

    def add(name, nbytes, ms, uses):
        gbs = nbytes / ms / 1e6
        rows.append(dict(kernel=name, bytes=int(nbytes), ms=round(ms, 5), gbs=round(gbs, 1), frac=round(gbs / peaks["hbm"], 4), launches_per_step=uses))

    # Adam over each optimiser group: p, g, m, v read + p, m, v written (28 B / param) + the bf16 shadow of p (2 B / param)
    for opt, uses, nm in ((model.d_optim, 1, "adam[D group]"), (model.g_optim, 2, "adam[G group]")):
        b, e = opt.range()
        ms = once(lambda: opt.apply())
        add(nm + " (adam_tick + adam_dev)", (e - b) * 30, ms, uses)
    # batch norm at the largest layer of the step: d_h1 on the 2B batch [2B,16,16,128]: fp32 pre-norm in, bf16 out
    B2, H, C = 2 * batch, 16, 128
    rows_n = B2 * H * H
    pre = torch.randn(rows_n, C, device="cuda")
    y = torch.empty(rows_n, C, dtype=torch.bfloat16, device="cuda")
    dy = torch.randn(rows_n, C, device="cuda").to(torch.bfloat16)
    dx = torch.empty_like(dy)
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    mm, mv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    sm, sr = torch.empty(2, C, device="cuda"), torch.empty(2, C, device="cuda")
    nb = L.gg_bn_workspace_bytes(C, 2)
    ws = torch.zeros(nb // 8, dtype=torch.float64, device="cuda")
    # statistics pass alone (what the GEMM epilogue replaces) + apply
    ms = once(lambda: check(L.gg_bn_fwd_train(ptr(pre), 0, ptr(y), 1, rows_n, C, 2, ptr(gamma), ptr(beta), ptr(mm), ptr(mv), ptr(sm), ptr(sr), 1e-5, 0.9,
                                              2, 0.2, ptr(ws), nb, stream()), "bn_fwd"))
    add("bn_fwd_train: colsum + bn_train_apply [256 x 16 x 16 x 128 -> lrelu]".replace("256", str(B2)), rows_n * C * (4 + 4 + 2), ms, 0)
    ws.zero_()
    check(L.gg_bn_fwd_train(ptr(pre), 0, ptr(y), 1, rows_n, C, 2, ptr(gamma), ptr(beta), ptr(mm), ptr(mv), ptr(sm), ptr(sr), 1e-5, 0.9, 2, 0.2, ptr(ws), nb, stream()))
    ms = once(lambda: check(L.gg_bn_fwd_train_stats(ptr(pre), 0, ptr(y), 1, rows_n, C, 2, ptr(gamma), ptr(beta), ptr(mm), ptr(mv), ptr(sm), ptr(sr), 1e-5, 0.9,
                                                    2, 0.2, ptr(ws), stream()), "bn_apply"))
    add(f"bn_train_apply [{B2} x 16 x 16 x 128 -> lrelu] (statistics from the GEMM epilogue)", rows_n * C * (4 + 2), ms, 21)
    ms = once(lambda: check(L.gg_bn_bwd(ptr(pre), 0, ptr(dy), 1, ptr(dx), 1, rows_n, C, 2, ptr(gamma), ptr(beta), ptr(sm), ptr(sr), None, None, 2, 0.2, 3,
                                        ptr(ws), nb, stream()), "bn_bwd_apply"))
    add(f"bn_bwd_apply [{B2} x 16 x 16 x 128] (reductions from the dgrad epilogue)", rows_n * C * (4 + 2 + 2), ms, 17)
    ms = once(lambda: check(L.gg_bn_bwd(ptr(pre), 0, ptr(dy), 1, ptr(dx), 1, rows_n, C, 2, ptr(gamma), ptr(beta), ptr(sm), ptr(sr), None, None, 2, 0.2, 1,
                                        ptr(ws), nb, stream()), "bn_bwd"))
    add(f"bn_bwd two-pass: memset + colsum + bn_bwd_apply [{B2} x 16 x 16 x 128] (not in the step any more: every reduction rides in a "
        "dgrad or loss-head launch)", rows_n * C * (2 * (4 + 2) + 2), ms, 0)
    # the generator's input projection (model.py:304: z[B,100] -> [B,8192], batch norm over 512 channels): streams its matrix
    zin, od, Cc = 100, 8192, 512
    zx = torch.randn(batch, zin, device="cuda")
    Wm, bm = torch.randn(zin, od, device="cuda") * 0.02, torch.zeros(od, device="cuda")
    yl = torch.empty(batch, od, device="cuda")
    st8 = torch.zeros(2 * Cc, dtype=torch.float64, device="cuda")
    ms = once(lambda: check(L.gg_linear_fwd_stats(ptr(zx), 0, ptr(Wm), ptr(bm), ptr(yl), batch, zin, od, Cc, 1, ptr(st8), stream()), "linear_fwd_stats"))
    add(f"g_h0_lin fwd + batch statistics [{batch} x 100 -> 8192] (thin_fwd)", zin * od * 4 + batch * od * 4, ms, 3)
    dyl = torch.randn(batch, od, device="cuda").to(torch.bfloat16)
    dWm = torch.zeros_like(Wm)
    ms = once(lambda: check(L.gg_linear_wgrad(ptr(zx), 0, ptr(dyl), 1, ptr(dWm), None, batch, zin, od, stream()), "linear_wgrad"))
    add(f"g_h0_lin wgrad [{batch} x 100 -> 8192] (thin_wgrad: dW += x^T dy)", batch * od * 2 + 2 * zin * od * 4, ms, 2)
    # image-side conv layers (3 channels <-> 64): HBM-bound by design (31-63 flop/B)
    for r in conv_rows:
        if r.get("bytes"):
            add(r["kernel"] + f" ({r['path']}, 10 back-to-back launches after one flush)", r["bytes"], r["ms"], r["uses"])
    # loss head of the D update (model.py:277 + 121-131): d_h3_lin over [2B, 8192] + both cross-entropy means in one launch; its
    # backward (Matrix / bias gradients, dh, d_bn3's backward reductions) in another.  Latency-bound: 2-8 MB per launch.
    R, F_, Cb = 2 * batch, 8192, 512
    h = torch.randn(R, F_, device="cuda").to(torch.bfloat16)
    w1, b1 = torch.randn(F_, device="cuda") * 0.02, torch.zeros(1, device="cuda")
    lg, parts, dl = torch.empty(R, device="cuda"), torch.empty(3, device="cuda"), torch.empty(R, device="cuda")
    tk = torch.zeros(4, dtype=torch.int32, device="cuda")
    I2, F2 = ctypes.c_int32 * 2, ctypes.c_float * 2
    ms = once(lambda: check(L.gg_loss_head_fwd(ptr(h), 1, ptr(w1), ptr(b1), R, F_, I2(0, batch), I2(batch, R), F2(1.0, 0.0), F2(1.0, 1.0), 2,
                                               ptr(lg), ptr(parts), ptr(dl), ptr(tk), stream()), "loss_head_fwd"))
    add(f"loss_head_fwd [{R} x {F_} -> logits, 2 cross-entropy means, dlogits] (latency-bound)", R * F_ * 2 + F_ * 4, ms, 3)
    pre = torch.randn(R, F_, device="cuda")
    dh, dW, db = torch.empty_like(h), torch.zeros(F_, device="cuda"), torch.zeros(1, device="cuda")
    smn, srs = torch.zeros(2, Cb, device="cuda"), torch.ones(2, Cb, device="cuda")
    gam, bet = torch.ones(Cb, device="cuda"), torch.zeros(Cb, device="cuda")
    sums = torch.zeros(2 * 2 * Cb, dtype=torch.float64, device="cuda")
    fz = ctypes.c_int32(0)
    ms = once(lambda: check(L.gg_loss_head_bwd(ptr(h), 1, ptr(dl), ptr(w1), R, F_, ptr(dW), ptr(db), ptr(dh), ptr(pre), ptr(smn), ptr(srs), ptr(gam),
                                               ptr(bet), 2, 0.2, 2, Cb, ptr(sums), ctypes.byref(fz), stream()), "loss_head_bwd"))
    add(f"loss_head_bwd [{R} x {F_}: dW, db, dh + d_bn3 backward reductions] (latency-bound)", R * F_ * (2 + 4 + 2) + 2 * F_ * 4, ms, 3)
    # decode tail of the input frames (z_model_lib.py:339-346: cv2.resize INTER_LINEAR + BGR->RGB + /127.5-1), config 3's batch of
    # 32 clips x 16 frames, uint8 128 x 128 -> fp32 64 x 64 (OpenCV's 2x shortcut: every source byte is read): byte work, HBM-bound
    nf, Hs, Sd = 32 * 16, 128, 64
    fr = torch.randint(0, 256, (nf, Hs, Hs, 3), dtype=torch.uint8, device="cuda")
    fo = torch.empty(nf, Sd, Sd, 3, device="cuda")
    ms = once(lambda: check(L.gg_frames_to_input(ptr(fr), nf, Hs, Hs, Hs * Hs * 3, Hs * 3, ptr(fo), Sd, Sd, 1, stream()), "frames_to_input"))
    add(f"frames_to_input [{nf} frames u8 {Hs}x{Hs}x3 -> fp32 {Sd}x{Sd}x3] (input side of the e2e path, not in the device-timed step)",
        nf * Hs * Hs * 3 + nf * Sd * Sd * 12, ms, 0)
    rows[-1]["cpu_reference"] = frames_cpu_reference(nf, Hs, Sd)      # the reference's own host code for the same frames, timed beside it
    return rows


def frames_cpu_reference(nf, Hs, Sd):
    """The reference's per-frame host code (z_model_lib.py:339-346: cv2.resize INTER_LINEAR, cv2.cvtColor BGR2RGB, utils.transform) on
    the same number of frames, one host thread as in the reference's loop.  None when OpenCV is missing; never raises."""
    try:
        import cv2
        frames = np.random.RandomState(0).randint(0, 256, (nf, Hs, Hs, 3)).astype(np.uint8)
        out = np.zeros((nf, Sd, Sd, 3))
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            for i in range(nf):
                im = cv2.cvtColor(cv2.resize(frames[i], (Sd, Sd), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2RGB)
                out[i] = np.array(im) / 127.5 - 1.
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return {"ms": round(best * 1e3, 3), "frames_per_s": round(nf / best, 1), "threads": 1,
                "what": "cv2.resize + cv2.cvtColor + x / 127.5 - 1 per frame on one host thread (the reference's loop), best of 3"}
    except Exception as e:      # noqa: BLE001 -- a missing / different OpenCV must not cost the bench line
        return {"unavailable": str(e)[:120]}


def run_ours(args):
    from gifgan import ops
    from gifgan.dp import DataParallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dp = DataParallel() if world > 1 else None
    rank = dp.rank if dp else 0
    local_rank = dp.local_rank if dp else 0
    torch.cuda.set_device(local_rank)
    peaks = load_peaks()
    ops.set_precision(args.precision, tensor_cores=None if args.tensor_cores < 0 else bool(args.tensor_cores))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    total = args.warmup + args.steps * (args.repeats + 3) + 4

    def make(kind, **kw):
        if kind == "dcgan":
            return DcganWorkload(args, dp, rank, kw.get("batch", args.batch), 64, total_steps=total)
        if kind == "dcgan128":
            return DcganWorkload(args, dp, rank, kw.get("batch", args.batch), 128, total_steps=total)
        if kind == "mnist":
            return DcganWorkload(args, None, rank, kw.get("batch", 64), 28, mnist=True, total_steps=total)
        if kind == "vid":
            return VidWorkload(args, dp, rank, kw.get("clips", args.clips), what=kw.get("what", "BASELINE config 3"))
        if kind == "recurrent":
            return RecurrentWorkload(args, None, rank, kw.get("clips", 40))
        raise ValueError(kind)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    wl = make(args.workload)
    main = measure(wl, args, flush, dp, world, peaks)
    clocks = sampler.stop() if sampler else None
    last = wl.api_step(0) if hasattr(wl, "api_step") else None

    # ---- roofline of the dominant conv kernel (timed alone, burst peak) + HBM-bound kernels: config 2 model only
    roofline = None
    if rank == 0 and args.workload == "dcgan" and not args.no_roofline:
        rows = layer_rooflines(wl.model, wl.B, args.precision, flush)
        tc_rows = [r for r in rows if r["path"] == "tcgen05"] or rows
        top = tc_rows[0]
        ncu, ncu_file = ncu_by_layer(top["path"] if top["path"] in KERNEL_FAMILY_SOURCES else "tcgen05")
        prof = ncu.get(top["kernel"], {})
        traffic = (prof["dram_read_bytes"] + prof["dram_write_bytes"]) if prof else None
        roofline = {"bound": "tensor", "kernel": top["kernel"], "path": top["path"], "achieved": top["tflops"], "peak": peaks["tf_burst"],
                    "unit": "TFLOP/s", "frac": top["tflops"] / peaks["tf_burst"], "traffic": traffic, "peak_source": peaks["source"],
                    "traffic_note": (f"ncu --set full capture {ncu_file} of these kernel sources" if prof else
                                     "null: no committed ncu capture matches the current sources of this kernel family (sha %s)" % kernel_source_sha(top["path"] if top["path"] in KERNEL_FAMILY_SOURCES else "tcgen05")),
                    "ncu": {"file": ncu_file, "tensor_pipe_active_pct": prof.get("tensor_pct"), "duration_us": prof.get("dur_us")} if prof else None,
                    "flops_per_launch": top["flops"], "launch_ms": top["ms"],
                    "launch_ms_cold": top.get("ms_cold"),
                    "frac_cold": (top["flops"] / top["ms_cold"] / 1e9 / peaks["tf_burst"]) if top.get("ms_cold") else None,
                    "timing_cold": "one launch per L2 flush (operands from HBM), 3 repetitions -- the pessimistic reading; `frac` is the in-step-like one",
                    "timing": "CUDA events, 10 back-to-back launches after one L2 flush (operands L2-warm for 9 of 10, as in the step where the producer kernel has just written them), 5 repetitions",
                    "step": {"gemm_tflop_per_step": main["gemm_tflop_per_step"], "achieved_tflops": main["achieved_tflops_per_gpu"],
                             "frac_of_sustained": main["frac_of_sustained_bf16_peak"]},
                    "layers": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k != "bytes"} for r in rows[:14]],
                    "hbm_kernels": hbm_kernel_table(wl.model, wl.B, flush, peaks, rows)}

    # ---- the other BASELINE configurations, same run
    extra = []
    if not args.no_extra and args.workload == "dcgan":
        del wl
        torch.cuda.empty_cache()
        plan = ([("vid", dict(clips=32)), ("vid", dict(clips=64, what="z_model.py CLI default batch")), ("vid", dict(clips=128, what="config 4 per-GPU shard at 2 GPUs")),
                 ("mnist", {}), ("recurrent", {}), ("dcgan128", {})] if world == 1 else
                [("vid", dict(clips=256 // world, what=f"BASELINE config 4: global 256 clips over {world} GPUs"))])
        for kind, kw in plan:
            try:
                w2 = make(kind, **kw)
                r = measure(w2, args if kind != "recurrent" else argparse.Namespace(**{**vars(args), "steps": min(args.steps, 5), "repeats": min(args.repeats, 3)}),
                            flush, dp if kind == "vid" else None, world if kind == "vid" else 1, peaks)
                r["kind"] = kind
                if kind == "vid":
                    r["clips_per_s"] = r["value"] / 16
                    r["global_clips"] = kw.get("clips", args.clips) * world
                extra.append(r)
                del w2
                torch.cuda.empty_cache()
            except Exception as ex:       # an extra must never take the headline line down with it
                extra.append({"kind": kind, "error": repr(ex)[:300]})

    if rank != 0:
        return
    # ---- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if not args.no_cpu_baseline and world == 1 and args.workload == "dcgan":
        fps, sec, cores = cpu_reference_step_rate(args.batch, 3, 1)
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"3 full steps of batch {args.batch} after 1 warm-up ({sec:.2f} s/step), oracle restatement (torch-CPU fp32), not TensorFlow"}

    launches_per_step = main["gpu_launches_per_step"]
    n_steps_run = args.warmup + args.steps * (args.repeats + max(1, min(args.repeats, 3))) + 1
    line = {
        "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": main["workload"], "global_batch": (args.batch * world if args.workload.startswith("dcgan") or args.workload == "mnist" else None), "global_frames": main["value"] * main["ms_per_step"] / 1e3,
                   "parallelism": f"dp{world}", "l2": "flushed before every timed step (256 MB written, then read back so that no dirty lines remain)",
                   "repeats": args.repeats, "statistic": "median over repeats of the K-step total, max over ranks per repeat",
                   "bn": "per-replica statistics", "graph": not args.eager, "tensor_cores": bool(ops._USE_TC)},
        "repeats_ms_per_step": main["repeats_ms_per_step"],
        "e2e": main.get("e2e"),
        "gpu_launches": launches_per_step * n_steps_run,
        "gpu_launches_per_step": launches_per_step,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "losses": last, "extra": extra,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--repeats", type=int, default=5, help="repetitions of the K-step timed region; the median is reported")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dcgan", choices=["dcgan", "dcgan128", "vid", "mnist", "recurrent"])
    ap.add_argument("--precision", default=os.environ.get("GIFGAN_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--tensor-cores", type=int, default=-1, help="-1: library default, 0/1: force")
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU (dcgan workloads)")
    ap.add_argument("--clips", type=int, default=32, help="clips per GPU (vid workload)")
    ap.add_argument("--eager", action="store_true", help="no CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configurations")
    ap.add_argument("--no-roofline", action="store_true", help="skip the per-kernel roofline / HBM tables")
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

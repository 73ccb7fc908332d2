#!/usr/bin/env python
"""Turns the raw ncu outputs a GPU visit leaves in gpurun_out/ (scratch) into the small summaries committed
under profiles/:
    python tools/summarize_profiles.py <tag>
  gpurun_out/launches_<tag>.csv          -> profiles/<tag>_launches.md      (one graph replay = one step, per kernel)
  gpurun_out/prof_layers_<tag>_raw.csv   -> profiles/<tag>_ncu_full.csv     (selected `ncu --set full` metrics per launch)
"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

KEEP = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
]


def us(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)


def launches(tag):
    path = os.path.join(G, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    lines = open(path).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = [r for r in csv.DictReader(lines[start:]) if r["Metric Name"] == "gpu__time_duration.sum"]
    flush = [i for i, r in enumerate(rows) if "FillFunctor<unsigned char" in r["Kernel Name"]]
    segs = [(flush[j], flush[j + 1] - flush[j] - 1) for j in range(len(flush) - 1)]
    segs = [s for s in segs if s[1] > 100]
    if not segs:
        return
    s0, n = segs[-2] if len(segs) > 1 else segs[-1]
    step = rows[s0 + 1:s0 + 1 + n]
    agg = collections.OrderedDict()
    tot = 0.0
    for r in step:
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        name = name if len(name) < 70 else name[:34] + ".." + name[-34:]
        k = (name, r["Grid Size"], r["Block Size"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us(r)
        tot += us(r)
    with open(os.path.join(P, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: launch list of ONE train step (one CUDA-graph replay between two L2 flushes)\n\n")
        f.write("Source: `ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 2 --warmup 3 "
                "--no-cpu-baseline` (tools/gpu_round.sh).\nPer-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n")
        f.write(f"launches in the step: {n}; sum of kernel durations: {tot:.1f} us\n\n")
        f.write("| kernel | grid | block | launches | total us | avg us | share |\n|---|---|---|---:|---:|---:|---:|\n")
        for (name, g, b), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {g} | {b} | {c} | {t:.1f} | {t / c:.2f} | {100 * t / tot:.1f}% |\n")
    print("wrote", f"{tag}_launches.md", n, "launches", round(tot, 1), "us")


def full(tag):
    path = os.path.join(G, f"prof_layers_{tag}_raw.csv")
    if not os.path.exists(path):
        return
    r = list(csv.reader(open(path)))
    hdr, units, rows = r[0], r[1], r[2:]
    idx = [(k, hdr.index(k)) for k in KEEP if k in hdr]
    with open(os.path.join(P, f"{tag}_ncu_full.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([k for k, _ in idx])
        w.writerow([units[i] for _, i in idx])
        for row in rows:
            w.writerow([row[i][:80] for _, i in idx])
    print("wrote", f"{tag}_ncu_full.csv", len(rows), "kernels")


def by_layer(tag):
    """Map the `ncu --set full` rows (two launches per layer kernel: warm-up + timed) back to bench.py's layer labels
    through the ORDER lines tools/layer_kernels.py prints."""
    raw, order = os.path.join(G, f"prof_layers_{tag}_raw.csv"), os.path.join(G, f"plain_layers_{tag}.log")
    if not (os.path.exists(raw) and os.path.exists(order)):
        return
    labels = [l.split()[1] for l in open(order) if l.startswith("ORDER ")]
    r = list(csv.reader(open(raw)))
    hdr, units, rows = r[0], r[1], r[2:]
    ix = {k: hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {}
    for i, lab in enumerate(labels):
        if 2 * i + 1 >= len(rows):
            break
        row = rows[2 * i + 1]                                     # the timed launch
        out[lab] = dict(kernel=row[ix["Kernel Name"]].split("(")[0], dur_us=float(row[ix["gpu__time_duration.sum"]]),
                        dram_read_bytes=float(row[ix["dram__bytes_read.sum"]]) * scale[units[ix["dram__bytes_read.sum"]]],
                        dram_write_bytes=float(row[ix["dram__bytes_write.sum"]]) * scale[units[ix["dram__bytes_write.sum"]]],
                        tensor_pct=float(row[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]]))
    import json
    sys.path.insert(0, ROOT)
    import bench
    # bench.py quotes `traffic` only from a capture of the CURRENT sources of the kernel's family.  Run this on the tree the capture
    # ran on (or pass GG_CAPTURE_CSRC=<checkout of that commit's gif-gan_b200/csrc>).
    csrc = os.environ.get("GG_CAPTURE_CSRC")
    out["_kernel_source_sha"] = bench.kernel_source_sha(None, csrc)
    out["_kernel_source_sha_by_family"] = {k: bench.kernel_source_sha(k, csrc) for k in bench.KERNEL_FAMILY_SOURCES}
    with open(os.path.join(P, f"{tag}_ncu_by_layer.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", f"{tag}_ncu_by_layer.json", len(out), "layers")


if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    tag = sys.argv[1]
    launches(tag)
    full(tag)
    by_layer(tag)

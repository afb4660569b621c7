"""Turn gpurun_out/ ncu exports into the committed summaries under profiles/ (run here, no GPU)."""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"


def launches(path, dst):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over\n"
                f"# `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-torch-baseline --no-parity-check` (config 2, "
                f"B=1024); {len(data)} launches = the whole process: set-up + training steps of the warm-up, timed, breakdown "
                f"and end-to-end legs (the end-to-end leg captures its two CUDA graphs, whose replays ncu lists kernel by kernel), "
                f"{tot:.1f} us total device time (cold-cache, serialised: compare SHARES)\n\n")
        f.write("| share | total us | launches | kernel |\n|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {100 * t / tot:5.1f}% | {t:9.1f} | {n} | `{k}` |\n")


def raw(rep, dst, title):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
    idx = [hdr.index(w) for w in want if w in hdr]
    ki = hdr.index("Kernel Name")
    import json
    traffic = {}
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    for r in data:
        def mb(i):
            v = float(r[i].replace(",", ""))
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]
        traffic[r[ki].split("(")[0].replace("void ", "").replace("cmmvae::", "")] = mb(ir) + mb(iw)
    json.dump(traffic, open(dst.replace(".md", "_dram_bytes.json"), "w"), indent=1)
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on` (per launch; never a bench value)\n\n")
        f.write("| kernel | " + " | ".join(f"{hdr[i]} [{units[i]}]" for i in idx) + " |\n")
        f.write("|---|" + "---|" * len(idx) + "\n")
        for r in data:
            f.write(f"| `{r[ki].split('(')[0]}` | " + " | ".join(r[i] for i in idx) + " |\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if os.path.exists(os.path.join(GP, "launches.csv")):
        launches(os.path.join(GP, "launches.csv"), os.path.join(OUT, f"{tag}_launch_list.md"))
    for name, title in (("topk.ncu-rep", "top kernels at the bench shape (B=1024, G=60530, H=1024, 5% nnz)"),
                        ("dec4096.ncu-rep", "fused decoder at B=4096 (BASELINE config 3 batch), G=60530, H=1024"),
                        ("dec8192.ncu-rep", "fused decoder at B=8192 (BASELINE config 4 batch), G=60530, H=1024")):
        p = os.path.join(GP, name)
        if os.path.exists(p):
            raw(p, os.path.join(OUT, f"{tag}_ncu_{name.split('.')[0]}.md"), title)
    print(os.listdir(OUT))

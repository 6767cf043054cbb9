"""Summarise ncu artefacts brought back in gpurun_out/ into small text files under profiles/ (run in the dev container).
  python tools/ncu_summary.py launches gpurun_out/launches_X.csv profiles/X_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof_X.ncu-rep profiles/X_full.txt"""
import collections
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        k = r[ik].split("(")[0][:70]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {src}: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# {'kernel':70s} launches   total_us  share\n")
        for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k:72s} {n:6d} {v / 1e3:10.1f} {v / tot * 100:6.1f}%\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# {src}: ncu --set full --clock-control none --import-source on\n")
        for r in rows[2:]:
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k} [{units[i]}] = {r[i][:110]}\n")
            f.write("\n")
    print(open(dst).read())


def traffic(src, dst):
    """Launch list with per-launch duration and DRAM bytes (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
    dram__bytes_write.sum): writes the text table `dst` and, next to it, <dst minus _launches.txt>_traffic.json (what
    bench.py reports as roofline.traffic)."""
    import json
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    iid, ik, im, iu, iv = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
    per = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
        except ValueError:
            continue
        d = per.setdefault(r[iid], {"k": r[ik].split("(")[0].replace("void ", "").replace("embclip::", "")[:64]})
        d[r[im]] = v
    tc = ("conv_gemm", "gemm2sm", "conv3x3_halo", "bneck_tail")
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in per.values():
        a = agg[d["k"]]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot_us = sum(a[1] for a in agg.values())
    tot_b = sum(a[2] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {src}: one encoder step (B=256, all three heads): ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none\n")
        f.write("# (cold-cache, serialised: compare SHARES; dram = read + write bytes summed over the kernel's launches)\n")
        f.write(f"# {'kernel':62s} launches  total_us  share   dram_MB\n")
        for k, (n, us, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k:64s} {n:6d} {us:9.1f} {us / tot_us * 100:6.1f}% {b / 1e6:9.1f}\n")
        f.write(f"# total {tot_us:.1f} us, {tot_b / 1e9:.2f} GB DRAM traffic per step\n# per launch, in order:\n")
        for i, d in enumerate(per.values()):
            f.write(f"{i:3d} {d['k']:64s} {d.get('gpu__time_duration.sum', 0):8.1f} us  rd {d.get('dram__bytes_read.sum', 0) / 1e6:8.1f} MB  wr {d.get('dram__bytes_write.sum', 0) / 1e6:8.1f} MB\n")
    sel = [d for d in per.values() if any(t in d["k"] for t in tc)]
    out = {"source": f"{dst} (ncu dram__bytes_read.sum + dram__bytes_write.sum, one B=256 step)",
           "tensor_core_conv_kernels": {"launches": len(sel),
                                        "dram_bytes_per_step": sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in sel),
                                        "time_us_cold": sum(d.get("gpu__time_duration.sum", 0) for d in sel)},
           "all_kernels": {"launches": len(per), "dram_bytes_per_step": tot_b}}
    json.dump(out, open(dst.replace("_launches.txt", "_traffic.json"), "w"), indent=0)
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])

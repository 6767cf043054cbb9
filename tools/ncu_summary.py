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


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])

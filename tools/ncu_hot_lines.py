"""Hottest source lines of an ncu report (needs -lineinfo + --import-source on): `ncu -i rep --page source --csv` reduced to the
N lines with the most warp-stall samples, with the dominant stall reasons.  Run where the .ncu-rep is (the GPU box).
  python tools/ncu_hot_lines.py prof.ncu-rep out.txt [N]"""
import csv
import io
import subprocess
import sys


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = ""
    for extra in (["--print-source", "cuda,sass"], ["--print-source", "sass"], []):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + extra, capture_output=True, text=True).stdout
        if raw.count("\n") > 20:
            break
    out = []
    # the CSV holds one table per kernel launch, separated by header rows
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = None
    tables = []
    for r in rows:
        if r and r[0] in ("#", "Address", "Line", "Source") or (r and "Source" in r and "# Samples" in " ".join(r)):
            hdr = r
            tables.append((hdr, []))
        elif hdr is not None and len(r) == len(hdr):
            tables[-1][1].append(r)
    for ti, (hdr, body) in enumerate(tables[:3]):
        def col(name):
            for i, h in enumerate(hdr):
                if h.strip() == name:
                    return i
            return None
        isrc = col("Source")
        isamp = col("# Samples") if col("# Samples") is not None else col("Warp Stall Sampling (All Samples)")
        if isrc is None or isamp is None:
            out.append(f"table {ti}: columns {hdr[:12]} ...")
            continue
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") or "Stall" in h and "Sampling" not in h]
        def num(x):
            try:
                return float(x.replace(",", ""))
            except ValueError:
                return 0.0
        tot = sum(num(r[isamp]) for r in body) or 1.0
        body.sort(key=lambda r: -num(r[isamp]))
        out.append(f"== launch {ti}: {int(tot)} stall samples")
        for r in body[:top]:
            reasons = sorted(((num(r[i]), h) for i, h in stall_cols if num(r[i]) > 0), reverse=True)[:3]
            out.append(f"{num(r[isamp]) / tot * 100:5.1f}%  {r[isrc].strip()[:150]}   [{', '.join(f'{h}:{int(v)}' for v, h in reasons)}]")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:top + 5]))


if __name__ == "__main__":
    main()

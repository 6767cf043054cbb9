"""Summarise `ncu --page source --csv` output (SASS view, kept as text under gpurun_out/) into profiles/: per kernel launch the
distribution of warp-stall reasons over all sampled instructions and the hottest SASS instructions.
  python tools/ncu_source_summary.py gpurun_out/prof_X_source.csv profiles/X_stalls.txt"""
import collections
import csv
import sys

csv.field_size_limit(10 ** 9)


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src, errors="replace")))
    out = [f"# {src}: ncu --set full --import-source on, --page source --csv (SASS view; first launches of the capture)"]
    kernel, hdr, body = None, None, []

    seen = set()

    def flush():
        if not hdr or not body:
            return
        key = (kernel, len(body), body[0][2] if body else "")
        if key in seen or len(body) < 400:            # repeated table of the same launch / truncated last table of the excerpt
            return
        seen.add(key)
        isamp = hdr.index("# Samples")
        isrc = hdr.index("Source")
        stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        num = lambda x: float(x.replace(",", "")) if x.replace(",", "").replace(".", "", 1).isdigit() else 0.0
        tot = sum(num(r[isamp]) for r in body) or 1.0
        agg = collections.Counter()
        for r in body:
            for i, h in stall:
                agg[h] += num(r[i])
        st = sum(agg.values()) or 1.0
        out.append(f"\n== {kernel[:150]}")
        out.append(f"   {int(tot)} samples over {len(body)} SASS instructions in this excerpt; stall reasons: " +
                   ", ".join(f"{h[6:]} {v / st * 100:.0f}%" for h, v in agg.most_common(6)))
        for r in sorted(body, key=lambda r: -num(r[isamp]))[:14]:
            top = sorted(((num(r[i]), h[6:]) for i, h in stall if num(r[i]) > 0), reverse=True)[:2]
            out.append(f"   {num(r[isamp]) / tot * 100:5.1f}%  {r[isrc].strip()[:90]:90s} [{', '.join(f'{h} {int(v)}' for v, h in top)}]")

    for r in rows:
        if r and r[0] == "Kernel Name":
            flush()
            kernel, hdr, body = r[1], None, []
        elif r and r[0] == "Address":
            hdr = r
        elif hdr is not None and len(r) == len(hdr):
            body.append(r)
    flush()
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:60]))


if __name__ == "__main__":
    main()

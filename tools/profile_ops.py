"""Per-op device timing of one encoder forward (CUDA events inside the library), with each op's algorithmic
FLOPs / bytes and the achieved TFLOP/s and GB/s.  Writes gpurun_out/ops_<tag>.json and prints a table."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embclip_b200.encoder import ClipRN50Encoder  # noqa: E402
from embclip_b200.synthetic import synthetic_rn50_state_dict  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    tag = sys.argv[2] if len(sys.argv) > 2 else "ops"
    enc = ClipRN50Encoder(synthetic_rn50_state_dict(), "cuda:0")
    frames = torch.randn(B, 224, 224, 3, device="cuda")
    if os.environ.get("PROFILE_U8"):
        frames = torch.randint(0, 256, (B, 224, 224, 3), device="cuda", dtype=torch.uint8)
    heads = ("trunk", "avgpool", "attnpool")
    for _ in range(3):
        enc(frames, heads)
    torch.cuda.synchronize()
    runs = [enc.profile(frames, heads) for _ in range(5)]
    names = [n for n, _ in runs[0]]
    ms = [sorted(r[i][1] for r in runs)[len(runs) // 2] for i in range(len(names))]
    acts = enc.activations(B)
    shapes = {k: tuple(v.shape) for k, v in acts.items()}
    pinfo = {n: s for n, _, s, _, _ in enc.param_infos}
    rows = []
    prev = (B, 224, 224, 3)
    for n, t in zip(names, ms):
        flop = byts = 0
        if n + ".w" in pinfo and n in shapes and n != "stem.conv1":
            cout, k = pinfo[n + ".w"]
            m = shapes[n][0] * shapes[n][1] * shapes[n][2]
            conv3 = ".conv2" in n or n in ("stem.conv2", "stem.conv3")
            src = {"stem.conv2": "stem.conv1", "stem.conv3": "stem.conv2"}.get(n, n.replace(".conv2", ".conv1"))
            m_in = shapes[src][0] * shapes[src][1] * shapes[src][2] if conv3 else m     # fused avgpool: 4x the output pixels
            flop = 2.0 * m_in * cout * k
            kin = k // 9 if conv3 else k
            byts = m_in * kin * 2 + m * cout * (4 if acts[n].dtype == torch.float32 else 2) + cout * k * 2
            if ".conv3" in n and k == cout // 4:
                byts += m * cout * 2     # residual read
        rows.append(dict(op=n, ms=t, shape=shapes.get(n), gflop=flop / 1e9, mb=byts / 1e6,
                         tflops=flop / (t * 1e-3) / 1e12 if t > 0 else 0, gbs=byts / (t * 1e-3) / 1e9 if t > 0 else 0))
    total = sum(ms)
    print(f"B={B} total {total:.3f} ms  ({B / total * 1e3:.0f} frames/s if back-to-back)")
    for r in rows:
        print(f"{r['op']:22s} {r['ms']:7.3f} ms  {str(r['shape']):24s} {r['gflop']:8.1f} GF {r['tflops']:7.1f} TF/s  {r['mb']:8.1f} MB {r['gbs']:7.0f} GB/s")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(batch=B, total_ms=total, ops=rows), open(os.path.join(ROOT, "gpurun_out", f"ops_{tag}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

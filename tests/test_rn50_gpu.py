"""GPU parity of the CLIP-RN50 encoder (through the C ABI) against the oracle.

Two bars (DESIGN.md "Numerics"):
  * vs oracle/fp16_path.py -- same rounding points as the kernels, so only fp32 accumulation order differs:
    rel-L2 per frame <= 2e-4, layer by layer.  This is the "do the kernels compute the design" check.
  * vs oracle/clip_model.py in fp32 -- the north-star bar: rel-L2 per frame <= 1e-3 on every head.
Golden fixtures (tests/golden/rn50_b2_*.pt, made by tests/golden/make_golden.py from the oracle) pin the
same check without needing the oracle at run time.
"""
import os

import pytest
import torch

from conftest import synthetic_frames

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel_l2(a, b):
    a, b = a.float().flatten(1), b.float().flatten(1)
    return ((a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-12)).max().item()


@pytest.fixture(scope="module")
def encoder(built_lib, rn50_visual):
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from embclip_b200.encoder import ClipRN50Encoder
    return ClipRN50Encoder(rn50_visual.state_dict(), "cuda:0")


@pytest.mark.parametrize("batch", [2, 5])
def test_rn50_per_op_vs_fp16_path(encoder, rn50_visual, batch):
    """Per-op isolation: every op's output is compared with the oracle evaluated on the inputs the kernels
    actually saw (feed=GPU activations).  Only fp32 summation order differs -> rel-L2 <= 1e-4 per op."""
    from oracle.fp16_path import rn50_fp16_path
    frames = synthetic_frames(batch, seed=batch)
    out = encoder(frames.cuda(), want=("trunk", "avgpool", "attnpool"))
    torch.cuda.synchronize()
    acts = {k: v.cpu() for k, v in encoder.activations(batch).items()}
    nchw = frames.permute(0, 3, 1, 2).contiguous()
    ref = rn50_fp16_path(rn50_visual, nchw, feed=acts)
    report = [(name, rel_l2(t, ref[name].reshape(t.shape))) for name, t in acts.items()]
    bad = [x for x in report if not x[1] <= 1e-4]
    assert not bad, f"ops off: {bad}; all: {report}"
    assert rel_l2(out["trunk"].cpu(), ref["trunk_nchw"]) <= 1e-6          # pure layout change of the fed tensor
    assert rel_l2(out["avgpool"].cpu(), ref["avgpool"]) <= 1e-5
    assert rel_l2(out["attnpool"].cpu(), ref["attnpool"]) <= 1e-4
    # chained end to end the two fp16 pipelines decorrelate (see oracle/fp16_path.py); bound it loosely
    chained = rn50_fp16_path(rn50_visual, nchw)
    assert rel_l2(out["trunk"].cpu(), chained["trunk_nchw"]) <= 1.5e-3


def test_rn50_vs_fp32_oracle(encoder, rn50_visual):
    frames = synthetic_frames(4, seed=7)
    with torch.no_grad():
        t = rn50_visual.trunk(frames.permute(0, 3, 1, 2).contiguous())
        ap = rn50_visual.attnpool(t)
    out = encoder(frames.cuda(), want=("trunk", "avgpool", "attnpool"))
    torch.cuda.synchronize()
    e_trunk, e_avg, e_ap = rel_l2(out["trunk"].cpu(), t), rel_l2(out["avgpool"].cpu(), t.mean((2, 3))), rel_l2(out["attnpool"].cpu(), ap)
    print(f"rel-L2 vs fp32 oracle: trunk {e_trunk:.3e} avgpool {e_avg:.3e} attnpool {e_ap:.3e}")
    assert e_trunk <= 1e-3, e_trunk
    assert e_avg <= 1e-3, e_avg
    assert e_ap <= 1e-3, e_ap


def test_rn50_golden(encoder):
    path = os.path.join(GOLDEN, "rn50_b2_seed0.pt")
    g = torch.load(path)
    frames = synthetic_frames(2, seed=0)
    assert torch.equal(frames[:, ::37, ::41].contiguous(), g["frames_probe"]), "synthetic frame generator drifted"
    out = encoder(frames.cuda(), want=("trunk", "avgpool", "attnpool"))
    torch.cuda.synchronize()
    assert rel_l2(out["avgpool"].cpu(), g["avgpool"]) <= 1e-3
    assert rel_l2(out["attnpool"].cpu(), g["attnpool"]) <= 1e-3
    assert rel_l2(out["trunk"].cpu()[:, ::16], g["trunk_every16"]) <= 1e-3


def test_rn50_head_selection_and_determinism(encoder):
    frames = synthetic_frames(3, seed=11).cuda()
    a = encoder(frames, want=("trunk",))["trunk"].clone()
    b = encoder(frames, want=("trunk", "attnpool"))
    torch.cuda.synchronize()
    assert torch.equal(a, b["trunk"])                                   # bit-identical run to run
    c = encoder(frames[:1], want=("trunk",))["trunk"]
    torch.cuda.synchronize()
    assert torch.equal(c[0], a[0])                                      # batch-invariant (no cross-frame mixing)
    assert set(encoder(frames, want=("avgpool",)).keys()) == {"avgpool"}
    empty = encoder(frames[:0], want=("trunk",))["trunk"]
    assert empty.shape == (0, 2048, 7, 7)


def test_rn50_rejects_bad_input(encoder):
    with pytest.raises(ValueError):
        encoder(torch.zeros(2, 3, 224, 224, device="cuda"))
    with pytest.raises(ValueError):
        encoder(torch.zeros(2, 224, 224, 3, device="cuda", dtype=torch.float16))
    with pytest.raises(ValueError):
        encoder(torch.zeros(2, 224, 224, 3))


def test_rn50_large_batch_properties(encoder):
    """BASELINE size (B=256): size-independent properties -- every frame's result equals the same frame run
    in a small batch (frames are independent), and duplicated frames give identical rows."""
    B = 256
    small = synthetic_frames(8, seed=5).cuda()
    frames = small.repeat(B // 8, 1, 1, 1)
    big = encoder(frames, want=("trunk", "attnpool"))
    ref = encoder(small, want=("trunk", "attnpool"))
    torch.cuda.synchronize()
    for k in ("trunk", "attnpool"):
        assert torch.equal(big[k][:8], ref[k]), k
        assert torch.equal(big[k].view(B // 8, 8, -1), ref[k].view(1, 8, -1).expand(B // 8, -1, -1)), k


def test_rn50_uint8_frames(encoder, rn50_visual):
    """Raw uint8 frames: the stem kernel applies (v / 255 - mean) / std itself.  Same
    result as handing over host-normalised fp32 frames, up to the rounding of the normalisation (one FMA on the device,
    div-sub-div on the host) -- and within the north-star bar of the fp32 oracle."""
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (3, 224, 224, 3), generator=g, dtype=torch.uint8)
    mean, std = torch.tensor(encoder.CLIP_RGB_MEANS), torch.tensor(encoder.CLIP_RGB_STDS)
    f32 = (u8.float() / 255.0 - mean) / std
    a = encoder(u8.cuda(), want=("trunk", "attnpool"))
    a = {k: v.clone() for k, v in a.items()}
    b = encoder(f32.cuda(), want=("trunk", "attnpool"))
    torch.cuda.synchronize()
    # stem.conv1 differs by 9e-6 (0.06 % of its fp16 roundings flip); the flips cascade through 50 fp16 layers exactly as
    # they do between any two fp16 pipelines with different summation order (oracle/fp16_path.py): bound it the same way
    for k in ("trunk", "attnpool"):
        assert rel_l2(a[k].cpu(), b[k].cpu()) <= 1.5e-3, k
    acts_u8 = encoder.activations(3)          # (second forward overwrote the workspace: re-run the raw path for its stem)
    encoder(u8.cuda(), want=("trunk",))
    torch.cuda.synchronize()
    stem_u8 = encoder.activations(3)["stem.conv1"].clone()
    encoder(f32.cuda(), want=("trunk",))
    torch.cuda.synchronize()
    assert rel_l2(stem_u8.cpu(), encoder.activations(3)["stem.conv1"].cpu()) <= 5e-5
    with torch.no_grad():
        t = rn50_visual.trunk(f32.permute(0, 3, 1, 2).contiguous())
    assert rel_l2(a["trunk"].cpu(), t) <= 1e-3
    with pytest.raises(ValueError):
        encoder(torch.zeros(2, 224, 224, 3, device="cuda", dtype=torch.int32))


def test_profile_lists_every_launch_for_both_frame_dtypes(encoder):
    """embclip_rn50_profile / _profile_u8: one (name, ms) per launch of the forward, same op list for fp32 and raw uint8 frames
    (the uint8 entry exists because the fp32 one would read a uint8 buffer as floats), and the profiled pass leaves the same
    outputs as a plain forward."""
    g = torch.Generator().manual_seed(11)
    u8 = torch.randint(0, 256, (2, 224, 224, 3), generator=g, dtype=torch.uint8).cuda()
    mean, std = torch.tensor(encoder.CLIP_RGB_MEANS), torch.tensor(encoder.CLIP_RGB_STDS)
    f32 = ((u8.cpu().float() / 255.0 - mean) / std).cuda()
    heads = ("trunk", "avgpool", "attnpool")
    ops_f = encoder.profile(f32, heads)
    ops_u = encoder.profile(u8, heads)
    torch.cuda.synchronize()
    assert [n for n, _ in ops_f] == [n for n, _ in ops_u]
    assert len(ops_f) == encoder.launches_per_forward(heads)
    assert all(0.0 < ms < 50.0 for _, ms in ops_f + ops_u)
    names = [n for n, _ in ops_f]
    assert names[0] == "stem.conv1" and "layer4.2.conv3" in names and "attnpool.c" in names
    assert "layer2.0.xpool" not in names          # written by layer1.2's fused tail (pooled-output variant), not a launch
    ref = {k: v.clone() for k, v in encoder(u8, want=heads).items()}
    encoder.profile(u8, heads)
    out = encoder(u8, want=heads)
    torch.cuda.synchronize()
    for k in heads:
        assert torch.equal(out[k], ref[k]), k


def _heavy_tailed_oracle(seed: int, calibrated: bool):
    """Weight sets with heavy-tailed per-channel BatchNorm statistics (VERDICT r1 "numerics margin").

    calibrated=False: every BN channel's scale gamma / sqrt(var) is multiplied by a log-normal factor (sigma 0.7: x0.25 .. x4
    at two sigma; 2 % of the channels x4 more, 2 % x0.05), split at random between gamma and the running variance, with the
    layer's overall gain kept -- per-channel scales span two orders of magnitude, the network's dynamics stay those of the
    default set.
    calibrated=True: gains log-normal (sigma 0.6, same outliers), shifts N(0, 0.3), and the running mean / variance SET TO the
    statistics of the activations that reach each layer (one train-mode pass with momentum 1), as a trained checkpoint's are.
    Centring on the true mean removes the large common-mode part of every post-ReLU sum, so each layer amplifies the relative
    perturbation it receives (measured x1.1 .. x1.2 per op): ANY one-rounding-per-op fp16 pipeline -- including the reference's
    own in-tree CUDA call, which runs CLIP converted to fp16 (thor_image_features.py:57) -- lands at 5e-2 .. 7e-2 on such a
    set; the test pins the kernels to that ideal fp16 path (oracle/fp16_path.py), not to 1e-3."""
    from oracle.clip_model import build_rn50, freeze_model, init_synthetic_rn50_visual
    torch.manual_seed(seed)
    m = init_synthetic_rn50_visual(build_rn50().visual, seed=1000 + seed)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, mod in m.named_modules():
            if not isinstance(mod, torch.nn.BatchNorm2d):
                continue
            c = mod.weight.numel()
            u = torch.rand(c, generator=g)
            if calibrated:
                gain = torch.exp(0.6 * torch.randn(c, generator=g))
                gain = torch.where(u < 0.02, gain * 5.0, gain)
                gain = torch.where(u > 0.98, gain * 0.02, gain)
                last = name.endswith("bn3") and name.startswith("layer") or "downsample" in name
                mod.weight.copy_(gain * (0.5 if last else 1.0))
                mod.bias.copy_(0.3 * torch.randn(c, generator=g))
                mod.momentum = 1.0
            else:
                sc = torch.exp(0.7 * torch.randn(c, generator=g))
                sc = torch.where(u < 0.02, sc * 4.0, sc)
                sc = torch.where(u > 0.98, sc * 0.05, sc)
                sc = sc / sc.pow(2).mean().sqrt()
                a = torch.rand(c, generator=g)
                mod.weight.mul_(sc.pow(a))
                mod.running_var.div_(sc.pow(2 * (1 - a)))
        if calibrated:
            m.train()
            m.trunk(synthetic_frames(4, seed=50 + seed).permute(0, 3, 1, 2).contiguous())     # running stats := batch stats
    return freeze_model(m)


def test_rn50_numerics_margin_sweep(built_lib):
    """Numerics margin over weight seeds and heavy-tailed BN statistics (run with -s for the table).  Each row: the GPU
    encoder vs the fp32 oracle, next to the IDEAL one-rounding-per-op fp16 path evaluated on the CPU (oracle/fp16_path.py:
    same rounding points, torch fp32 sums) vs the same oracle -- the floor of any fp16-operand kernel on that weight set.
      * 5 uncalibrated heavy-tailed sets: inside the 1e-3 bar whenever the ideal path is; never more than 10 % above the ideal path;
      * 2 calibrated (trained-like) sets: the ideal fp16 path itself is at 5e-2 .. 7e-2 (see _heavy_tailed_oracle); the kernels
        must sit on it (<= 1.15 x), which is the strongest statement available without the real checkpoint."""
    from embclip_b200.encoder import ClipRN50Encoder
    from oracle.fp16_path import rn50_fp16_path
    rows = []
    for calibrated, seed in [(False, s_) for s_ in range(5)] + [(True, 0), (True, 1)]:
        oracle = _heavy_tailed_oracle(seed, calibrated)
        frames = synthetic_frames(1, seed=90 + seed)
        nchw = frames.permute(0, 3, 1, 2).contiguous()
        with torch.no_grad():
            t = oracle.trunk(nchw)
            ap = oracle.attnpool(t)
        ideal = rn50_fp16_path(oracle, nchw)
        i_t, i_a = rel_l2(ideal["trunk_nchw"], t), rel_l2(ideal["attnpool"], ap)
        enc = ClipRN50Encoder(oracle.state_dict(), "cuda:0")
        out = enc(frames.cuda(), want=("trunk", "avgpool", "attnpool"))
        torch.cuda.synchronize()
        assert torch.isfinite(out["trunk"]).all()
        rows.append((("calibrated" if calibrated else "heavy-tailed"), seed, rel_l2(out["trunk"].cpu(), t), i_t,
                     rel_l2(out["avgpool"].cpu(), t.mean((2, 3))), rel_l2(out["attnpool"].cpu(), ap), i_a))
        del enc
    for r in rows:
        print("%-12s seed %d: trunk %.3e (ideal fp16 path %.3e)  avgpool %.3e  attnpool %.3e (ideal %.3e)" % r)
    ht = [r for r in rows if r[0] == "heavy-tailed"]
    inside = sum(1 for r in ht if max(r[2], r[4], r[5]) <= 1e-3)
    print(f"numerics margin: {inside}/{len(ht)} heavy-tailed weight sets inside the 1e-3 bar on every head; worst trunk "
          f"{max(r[2] for r in ht):.3e} (ideal fp16 path {max(r[3] for r in ht):.3e})")
    for r in rows:
        assert r[2] <= 1.15 * r[3] + 2e-5, r                       # the kernels ARE the ideal fp16 path
        assert r[5] <= 1.15 * r[6] + 5e-5, r
        if r[0] == "heavy-tailed":
            assert r[2] <= max(1e-3, 1.1 * r[3]) and r[4] <= 1e-3, r

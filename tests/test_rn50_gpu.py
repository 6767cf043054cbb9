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


def _heavy_tailed_oracle(seed: int):
    """A trained-like weight set with heavy-tailed per-channel statistics (VERDICT r1 "numerics margin"): BatchNorm gains
    log-normal (sigma 0.6; 2 % of the channels x5, 2 % nearly dead), shifts N(0, 0.3), and the running mean / variance
    CALIBRATED to the statistics of the activations that reach each layer (one train-mode pass with momentum 1, as a
    trained checkpoint's are) -- so per-channel scales span two orders of magnitude while the network stays in its
    operating range."""
    from oracle.clip_model import build_rn50, freeze_model, init_synthetic_rn50_visual
    torch.manual_seed(seed)
    m = init_synthetic_rn50_visual(build_rn50().visual, seed=1000 + seed)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, mod in m.named_modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                c = mod.weight.numel()
                gain = torch.exp(0.6 * torch.randn(c, generator=g))
                u = torch.rand(c, generator=g)
                gain = torch.where(u < 0.02, gain * 5.0, gain)
                gain = torch.where(u > 0.98, gain * 0.02, gain)
                last = name.endswith("bn3") and name.startswith("layer") or "downsample" in name
                mod.weight.copy_(gain * (0.5 if last else 1.0))
                mod.bias.copy_(0.3 * torch.randn(c, generator=g))
                mod.momentum = 1.0
        m.train()
        m.trunk(synthetic_frames(4, seed=50 + seed).permute(0, 3, 1, 2).contiguous())     # running stats := batch stats
    return freeze_model(m)


def test_rn50_numerics_margin_sweep(built_lib):
    """>= 5 weight seeds x heavy-tailed BN statistics: the worst trunk / avg-pool / attention-pool rel-L2 against the fp32
    oracle, reported (run with -s) and held to the 1e-3 bar."""
    from embclip_b200.encoder import ClipRN50Encoder
    rows = []
    for seed in range(6):
        oracle = _heavy_tailed_oracle(seed)
        frames = synthetic_frames(2, seed=90 + seed)
        with torch.no_grad():
            t = oracle.trunk(frames.permute(0, 3, 1, 2).contiguous())
            ap = oracle.attnpool(t)
        enc = ClipRN50Encoder(oracle.state_dict(), "cuda:0")
        out = enc(frames.cuda(), want=("trunk", "avgpool", "attnpool"))
        torch.cuda.synchronize()
        assert torch.isfinite(out["trunk"]).all()
        rows.append((seed, rel_l2(out["trunk"].cpu(), t), rel_l2(out["avgpool"].cpu(), t.mean((2, 3))), rel_l2(out["attnpool"].cpu(), ap),
                     t.abs().max().item()))
        del enc
    for r in rows:
        print("heavy-tailed seed %d: trunk %.3e avgpool %.3e attnpool %.3e (max |trunk| %.1f)" % r)
    worst = max(max(r[1:4]) for r in rows)
    print(f"numerics margin: worst rel-L2 over {len(rows)} heavy-tailed weight sets = {worst:.3e} (bar 1e-3)")
    assert worst <= 1e-3, rows

"""Host-side weight preparation (embclip_b200/packing.py) against the oracle, on CPU: BN folding, tap-major
3x3 layout, conv3+downsample K-concatenation, attention-pool repacking, and the blob layout contract."""
import ctypes as C

import torch
import torch.nn.functional as F

from embclip_b200 import packing


def test_infer_cfg(rn50_visual):
    sd = {"visual." + k: v for k, v in rn50_visual.state_dict().items()}
    assert packing.infer_rn_cfg(sd) == dict(layers=(3, 4, 6, 3), width=64, heads=32, output_dim=1024, input_resolution=224)
    assert packing.infer_rn_cfg(rn50_visual.state_dict())["layers"] == (3, 4, 6, 3)


def test_folded_layers_match_oracle(rn50_visual):
    m = rn50_visual
    pk = packing.packed_tensors(m.state_dict())
    torch.manual_seed(0)
    with torch.no_grad():
        # stem conv1: [27, 32] fp32, (kh, kw, ci) major
        x = torch.randn(1, 3, 32, 32)
        ref = m.bn1(m.conv1(x))
        w = pk["stem.conv1.w"].view(3, 3, 3, 32).permute(3, 2, 0, 1)
        got = F.conv2d(x, w, pk["stem.conv1.b"], stride=2, padding=1)
        assert torch.allclose(got, ref, atol=1e-5, rtol=1e-5)
        # a 3x3: [Cout, (kh, kw, ci)] fp16 -- compare against the fp32-folded weights rounded the same way
        blk = m.layer2[1]
        x = torch.randn(1, 128, 8, 8)
        ref = blk.bn2(blk.conv2(x))
        w = pk["layer2.1.conv2.w"].float().view(128, 3, 3, 128).permute(0, 3, 1, 2)
        got = F.conv2d(x, w, pk["layer2.1.conv2.b"], padding=1)
        assert ((got - ref).norm() / ref.norm()).item() < 5e-4          # only the fp16 weight rounding
        # fused conv3 + downsample of a stride-2 block: [W3 | Wd] over [pool(t) ; pool(x)]
        blk = m.layer3[0]
        t, xin = torch.randn(1, 256, 8, 8), torch.randn(1, 512, 8, 8)
        ref = blk.bn3(blk.conv3(blk.avgpool(t))) + blk.downsample(xin)
        a = torch.cat([F.avg_pool2d(t, 2), F.avg_pool2d(xin, 2)], 1).permute(0, 2, 3, 1).reshape(-1, 768)
        wcat = pk["layer3.0.conv3.w"].float()
        assert wcat.shape == (1024, 768)
        got = (a @ wcat.t() + pk["layer3.0.conv3.b"]).view(1, 4, 4, 1024).permute(0, 3, 1, 2)
        assert ((got - ref).norm() / ref.norm()).item() < 5e-4
        # identity blocks keep a plain conv3
        assert pk["layer3.1.conv3.w"].shape == (1024, 256)
        # attention pool repacking
        ap = m.attnpool
        assert torch.equal(pk["attnpool.kT.w"], ap.k_proj.weight.t().contiguous().half())
        assert torch.equal(pk["attnpool.q.w"], (ap.q_proj.weight * 0.125).half())
        assert torch.equal(pk["attnpool.q.b"], ap.q_proj.bias * 0.125)


def test_blob_layout_matches_library(built_lib, rn50_visual):
    from embclip_b200 import _lib
    lib = _lib.load()
    cfg = _lib.RN50Cfg()
    cfg.layers[:] = (3, 4, 6, 3)
    cfg.width, cfg.heads, cfg.output_dim, cfg.input_resolution = 64, 32, 1024, 224
    h = C.c_void_p()
    assert lib.embclip_rn50_create(C.byref(cfg), C.byref(h)) == 0
    infos = []
    for i in range(lib.embclip_rn50_num_params(h)):
        pi = _lib.ParamInfo()
        lib.embclip_rn50_param_info(h, i, C.byref(pi))
        infos.append((pi.name.decode(), "f16" if pi.dtype == 0 else "f32", tuple(pi.shape[:pi.ndim]), int(pi.offset), int(pi.nbytes)))
    blob = packing.pack_blob(rn50_visual.state_dict(), infos)
    assert blob.numel() == lib.embclip_rn50_blob_bytes(h)
    pk = packing.packed_tensors(rn50_visual.state_dict())
    assert set(pk) == {n for n, *_ in infos}                      # nothing asked for is missing, nothing extra
    name, dt, shape, off, nb = next(x for x in infos if x[0] == "layer1.0.conv3.w")
    assert shape == (256, 128)                                    # 64 (conv3) + 64 (downsample) along K
    assert torch.equal(blob[off:off + nb].view(torch.float16).view(shape), pk[name])
    lib.embclip_rn50_destroy(h)


def test_transformer_tower_packing_matches_library_plan(built_lib):
    """ViT-B/32 and text tower: every tensor the library's plan asks for is produced by the packer with the right dtype /
    shape; shape inference mirrors clip/model.py build_model; the q third of in_proj carries the 1/8 attention scale."""
    import ctypes as C
    from embclip_b200 import _lib
    from embclip_b200.vit import infer_text_cfg, infer_vit_cfg, packed_tower_tensors
    from oracle.clip_model import build_vit_b32, init_synthetic_transformer
    torch.manual_seed(0)
    sd = init_synthetic_transformer(build_vit_b32(), seed=1234).state_dict()
    vcfg, tcfg = infer_vit_cfg(sd), infer_text_cfg(sd)
    assert (vcfg["width"], vcfg["layers"], vcfg["heads"], vcfg["output_dim"], vcfg["patch_size"], vcfg["input_resolution"]) == (768, 12, 12, 512, 32, 224)
    assert (tcfg["width"], tcfg["layers"], tcfg["heads"], tcfg["output_dim"], tcfg["context_length"], tcfg["vocab_size"]) == (512, 12, 8, 512, 77, 49408)
    lib = _lib.load()
    for cfg in (vcfg, tcfg):
        h = C.c_void_p()
        assert lib.embclip_tf_create(C.byref(_lib.TFCfg(**cfg)), C.byref(h)) == 0
        tensors = packed_tower_tensors(sd, cfg)
        names = set()
        for i in range(lib.embclip_tf_num_params(h)):
            pi = _lib.ParamInfo()
            assert lib.embclip_tf_param_info(h, i, C.byref(pi)) == 0
            name = pi.name.decode()
            names.add(name)
            t = tensors[name]
            assert tuple(t.shape) == tuple(pi.shape[:pi.ndim]), name
            assert t.dtype == (torch.float16 if pi.dtype == _lib.DTYPE_F16 else torch.float32), name
            assert pi.offset % 256 == 0
        assert names == set(tensors)
        assert lib.embclip_tf_workspace_bytes(h, 64) > 0
        assert lib.embclip_tf_destroy(h) == 0
    pk = packed_tower_tensors(sd, vcfg)
    w = sd["visual.transformer.resblocks.3.attn.in_proj_weight"]
    assert torch.equal(pk["blk3.qkv.w"][:768], (w[:768] * 0.125).half()) and torch.equal(pk["blk3.qkv.w"][768:], w[768:].half())
    conv = sd["visual.conv1.weight"]
    assert torch.equal(pk["patch.w"][5, (7 * 32 + 9) * 3 + 2], conv[5, 2, 7, 9].half())
    bad = dict(vcfg, width=640, heads=10)
    h = C.c_void_p()
    assert lib.embclip_tf_create(C.byref(_lib.TFCfg(**bad)), C.byref(h)) < 0

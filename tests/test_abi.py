"""The C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol include/embclip_b200.h
declares (no compute calls here).  Host-side plan queries that need no device are exercised too."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "embclip_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(embclip_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_python_binds(built_lib):
    from embclip_b200 import _lib
    assert set(header_symbols()) == set(_lib.SIGNATURES), "include/embclip_b200.h and _lib.SIGNATURES disagree"


def test_library_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(built_lib)
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    from embclip_b200 import _lib
    assert _lib.load().embclip_abi_version() == 6


def test_sass_is_blackwell_native(built_lib):
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):      # tcgen05.mma, TMA load/store, tcgen05.ld
        assert mnemonic in sass, f"{mnemonic} missing from SASS"
    assert "HMMA." not in sass.replace("UTCHMMA", ""), "legacy mma.sync path present"
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout


def test_plan_queries_without_gpu(built_lib):
    from embclip_b200 import _lib
    lib = _lib.load()
    cfg = _lib.RN50Cfg()
    cfg.layers[:] = (3, 4, 6, 3)
    cfg.width, cfg.heads, cfg.output_dim, cfg.input_resolution = 64, 32, 1024, 224
    h = C.c_void_p()
    assert lib.embclip_rn50_create(C.byref(cfg), C.byref(h)) == 0
    n = lib.embclip_rn50_num_params(h)
    names, end = [], 0
    for i in range(n):
        pi = _lib.ParamInfo()
        assert lib.embclip_rn50_param_info(h, i, C.byref(pi)) == 0
        assert pi.offset % 256 == 0 and pi.offset >= end
        end = pi.offset + pi.nbytes
        names.append(pi.name.decode())
    assert lib.embclip_rn50_blob_bytes(h) >= end
    # 1 stem conv1 + 2 stem convs + 16 blocks x 3 convs, each (w, b); attnpool: pos, q(w,b), kT, v(w,b), c(w,b)
    assert n == 2 * (3 + 48) + 8 + 1          # + stem.conv1.wtc (tensor-core stem weights)
    assert "layer4.2.conv3.w" in names and "attnpool.kT.w" in names
    # workspace scales linearly with batch (up to 1 KiB alignment per tensor)
    w1, w8 = lib.embclip_rn50_workspace_bytes(h, 1), lib.embclip_rn50_workspace_bytes(h, 8)
    assert 0 < w1 and 7.9 * w1 < w8 <= 8 * w1
    # trunk launches: 3 stem convs (pool fused into conv3), 16 x 3 convs minus the 3 conv1s that layer 1's bneck_tail launches
    # compute (layer1.1, layer1.2, layer2.0) and the 2 that layer 2's bneck_tail_stream launches compute (layer2.2, layer2.3),
    # 3 identity-path pools minus the one layer1.2's bneck_tail writes itself (pooled-output variant); + heads
    assert lib.embclip_rn50_launches_per_forward(h, 0, 0, 0) == 3 + 48 - 3 - 2 + 3 - 1
    assert lib.embclip_rn50_launches_per_forward(h, 1, 1, 1) == 3 + 48 - 3 - 2 + 3 - 1 + 2 + 6
    # error paths: state and argument checks happen before any CUDA call
    assert lib.embclip_rn50_forward(h, None, 1, None, None, None, None, 0, None) == -1
    bad = _lib.RN50Cfg()
    bad.layers[:] = (3, 4, 6, 3)
    bad.width, bad.heads, bad.output_dim, bad.input_resolution = 80, 40, 640, 224
    h2 = C.c_void_p()
    assert lib.embclip_rn50_create(C.byref(bad), C.byref(h2)) == -1
    assert b"width" in lib.embclip_last_error()
    # RN50x16 (width 96) at its native 384 x 384: 145 tokens -> planned without the attention-pool head; at AllenAct's 224 x 224
    # with output_dim 0 (checkpoint positional embedding does not fit) likewise; stem channels are carried as 64
    x16 = _lib.RN50Cfg()
    x16.layers[:] = (6, 8, 18, 8)
    x16.width, x16.heads, x16.output_dim, x16.input_resolution = 96, 48, 768, 384
    h3 = C.c_void_p()
    assert lib.embclip_rn50_create(C.byref(x16), C.byref(h3)) == 0
    names = {}
    for i in range(lib.embclip_rn50_num_params(h3)):
        pi = _lib.ParamInfo()
        assert lib.embclip_rn50_param_info(h3, i, C.byref(pi)) == 0
        names[pi.name.decode()] = tuple(pi.shape[:pi.ndim])
    assert names["stem.conv1.w"] == (27, 64) and names["stem.conv2.w"] == (64, 9 * 64) and names["stem.conv3.w"] == (96, 9 * 64)
    assert names["layer4.7.conv3.w"] == (3072, 768) and not any(n.startswith("attnpool") for n in names)
    assert lib.embclip_rn50_launches_per_forward(h3, 1, 1, 0) == 3 + 3 * 40 + 3 + 2
    assert lib.embclip_rn50_destroy(h3) == 0
    x16.input_resolution = 224
    assert lib.embclip_rn50_create(C.byref(x16), C.byref(h3)) == 0
    assert any(lib.embclip_rn50_param_info(h3, i, C.byref(pi)) == 0 and pi.name.startswith(b"attnpool") for i in range(lib.embclip_rn50_num_params(h3)))
    assert lib.embclip_rn50_destroy(h3) == 0
    assert lib.embclip_rn50_destroy(h) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from embclip_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU"):
        _lib.load()


def test_actor_critic_plan_matches_upstream_state_dict(built_lib):
    """The flat parameter layout the library reports carries the upstream (AllenAct) state_dict names and shapes,
    in 256-B aligned, non-overlapping slots; 3,480,775 parameters."""
    from embclip_b200 import _lib
    from oracle.allenact_models import ResnetTensorNavActorCritic
    lib = _lib.load()
    cfg = _lib.ACCfg(feat_channels=2048, feat_pixels=49, compress_hidden=128, compress_out=32, goal_dims=32,
                     combine_hidden=128, combine_out=32, hidden=512, num_actions=6, num_goals=12)
    h = C.c_void_p()
    assert lib.embclip_ac_create(C.byref(cfg), C.byref(h)) == 0
    ref = {k: tuple(v.shape) for k, v in ResnetTensorNavActorCritic().state_dict().items()}
    got, end, total = {}, 0, 0
    for i in range(lib.embclip_ac_num_params(h)):
        pi = _lib.ParamInfo()
        assert lib.embclip_ac_param_info(h, i, C.byref(pi)) == 0
        assert pi.offset % 256 == 0 and pi.offset >= end and pi.dtype == _lib.DTYPE_F32
        end = pi.offset + pi.nbytes
        got[pi.name.decode()] = tuple(pi.shape[:pi.ndim])
        total += pi.nbytes // 4
    assert got == ref
    assert total == 3_480_775
    assert lib.embclip_ac_param_floats(h) * 4 >= end
    assert lib.embclip_ac_workspace_bytes(h, 128, 60) > 0
    bad = _lib.ACCfg(feat_channels=2048, feat_pixels=49, compress_hidden=128, compress_out=64, goal_dims=32,
                     combine_hidden=128, combine_out=32, hidden=512, num_actions=6, num_goals=12)
    h2 = C.c_void_p()
    assert lib.embclip_ac_create(C.byref(bad), C.byref(h2)) < 0 and b"compress_out" in lib.embclip_last_error()
    assert lib.embclip_ac_destroy(h) == 0

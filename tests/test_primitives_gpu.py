"""GPU parity of the primitive ops of libembclip_b200.so (called through the C ABI) against plain fp32
PyTorch on the same fp16-rounded operands.  Tolerance: the kernels accumulate in fp32 and round the result
to fp16 once, so |err| <= 2^-11 * |ref| + accumulation-order noise; we allow rtol 2e-3 / atol 2e-3*scale."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib(built_lib):
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from embclip_b200 import _lib
    return _lib.load()


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(lib, rc):
    assert rc == 0, lib.embclip_last_error().decode()


def _close(out, ref, what):
    scale = ref.abs().max().item() + 1e-6
    err = (out.float() - ref).abs()
    tol = 2e-3 * ref.abs() + 2e-3 * scale
    bad = err > tol
    if bad.any():
        idx = bad.nonzero()
        rows = idx[:, 0].unique()[:16].tolist()
        cols = idx[:, -1].unique()[:16].tolist()
        pytest.fail(f"{what}: {int(bad.sum())}/{bad.numel()} elements off, max err {err.max().item():.4g} "
                    f"(scale {scale:.3g}); first bad rows {rows} cols {cols}")


def _gemm_case(lib, M, N, K0, K1=0, bias=True, res=False, relu=False, out_f32=False, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    a0 = rn(M, K0).half()
    a1 = rn(M, K1).half() if K1 else None
    w = (rn(N, K0 + K1) * (K0 + K1) ** -0.5).half()
    b = rn(N) if bias else None
    r = rn(M, N).half() if res else None
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if out_f32 else torch.float16)
    _check(lib, lib.embclip_gemm_f16(_ptr(a0), _ptr(a1), _ptr(w), _ptr(b), _ptr(r), _ptr(out), M, N, K0, K1,
                                     int(relu), int(out_f32), _stream()))
    torch.cuda.synchronize()
    a = a0.float() if a1 is None else torch.cat([a0.float(), a1.float()], 1)
    ref = a @ w.float().t()
    if bias:
        ref = ref + b
    if res:
        ref = ref + r.float()
    if relu:
        ref = ref.relu()
    _close(out, ref, f"gemm M{M} N{N} K{K0}+{K1} bias{bias} res{res} relu{relu} f32{out_f32}")


@pytest.mark.parametrize("M,N,K", [
    (128, 128, 64),        # one tile, one k-block
    (128, 128, 256),       # pipeline wrap (4 k-blocks)
    (256, 128, 1024),      # 16 k-blocks > stages, 2 tiles
    (1000, 256, 512),      # partial last M tile, 2 N tiles
    (128 * 150, 64, 64),   # more tiles than SMs: persistent loop + accumulator double-buffering
    (3136, 64, 256),       # layer1 conv1 shape (one frame)
    (392, 512, 128),       # layer2 conv3 shape, BN 128
    (98, 2048, 512),       # layer4 conv3 shape: M < 128
    (5, 1024, 2048),       # attnpool c_proj at tiny batch
])
def test_gemm_plain(lib, M, N, K):
    _gemm_case(lib, M, N, K)


@pytest.mark.parametrize("N,K", [(32, 32), (64, 32), (32, 64), (64, 64), (128, 64), (128, 96), (64, 288)])
def test_gemm_tile_variants(lib, N, K):
    """every (BN, BK) instantiation incl. the 64-B-swizzle BK=32 path"""
    _gemm_case(lib, 777, N, K, relu=True)


def test_gemm_epilogues(lib):
    _gemm_case(lib, 500, 256, 128, res=True, relu=True)
    _gemm_case(lib, 500, 256, 128, bias=False)
    _gemm_case(lib, 500, 256, 128, out_f32=True, relu=True)
    _gemm_case(lib, 300, 128, 64, out_f32=True, res=True)


@pytest.mark.parametrize("M,N,K", [(128 * 37 + 5, 1024, 256), (128 * 19, 2048, 256), (128 * 150, 256, 256)])
def test_gemm_residual_wide_tile(lib, M, N, K):
    """conv3 + identity residual of layers 3 / 4 at batch size: enough 128 x 256 tiles for every SM -> conv_gemm_kernel<256, 64, true>
    (two staging chunks per epilogue warp, ring of four = one tile; ragged last M tile; persistent loop over several tiles per CTA)."""
    _gemm_case(lib, M, N, K, res=True, relu=True, seed=M % 89)


def test_gemm_k_concat(lib):
    """[W3 | Wd] . [t ; x]: the fused conv3 + downsample of a bottleneck's first block"""
    _gemm_case(lib, 784, 512, 128, K1=256, relu=True)
    _gemm_case(lib, 3136, 256, 64, K1=64, relu=True)
    _gemm_case(lib, 98, 2048, 512, K1=1024, relu=True, out_f32=True)


@pytest.mark.parametrize("M,N,K", [
    (2048, 256, 256),          # smallest shape routed to the CTA-pair kernel: 8 pair tiles, 4 k-blocks
    (4096 + 37, 512, 1024),    # partial last pair tile, 2 N tiles, pipeline wraps
    (256 * 80, 256, 512),      # more pair tiles than CTA pairs: persistent loop + accumulator double-buffering
    (12544, 2048, 512),        # layer4 conv3 shape
    (25600, 768, 3072),        # ViT-B/32 mlp.c_proj at B = 512
])
def test_gemm_cta_pair(lib, M, N, K):
    """shapes launch_gemm routes to gemm2sm_kernel (cta_group::2): N % 256 == 0, K >= 256, M >= 2048"""
    _gemm_case(lib, M, N, K, relu=True, seed=M % 97)


def test_gemm_cta_pair_epilogues(lib):
    _gemm_case(lib, 3000, 256, 256, res=True, relu=True)                 # fp16 residual read in the epilogue
    _gemm_case(lib, 3000, 512, 256, bias=False)
    _gemm_case(lib, 3000, 256, 512, out_f32=True, relu=True)             # fp32 rows stored directly
    _gemm_case(lib, 2500, 1024, 256, K1=512, relu=True)                  # K-concat second source (conv3 + downsample)
    _gemm_case(lib, 2500, 256, 320)                                      # K not a multiple of 64 -> stays on the 1-CTA kernel


def test_gemm_grouped(lib):
    """per-head contractions of AttentionPool2d"""
    torch.manual_seed(1)
    B, heads, E, hd = 6, 32, 2048, 64
    q = torch.randn(B, E, device="cuda").half()
    wkT = (torch.randn(E, E, device="cuda") * E ** -0.5).half()          # [c, (h,d)]
    out = torch.full((B, heads * E), float("nan"), device="cuda", dtype=torch.float16)
    _check(lib, lib.embclip_gemm_grouped_f16(_ptr(q), E, _ptr(wkT), E, E, None, _ptr(out), B, heads * E, hd,
                                             E, hd, hd, E, 0, 0, _stream()))
    torch.cuda.synchronize()
    ref = torch.einsum("bhd,chd->bhc", q.float().view(B, heads, hd), wkT.float().view(E, heads, hd)).reshape(B, heads * E)
    _close(out, ref, "grouped qk")

    xbar = torch.randn(B, heads * E, device="cuda").half()
    wv = (torch.randn(E, E, device="cuda") * E ** -0.5).half()
    bv = torch.randn(E, device="cuda")
    out2 = torch.full((B, E), float("nan"), device="cuda", dtype=torch.float16)
    _check(lib, lib.embclip_gemm_grouped_f16(_ptr(xbar), heads * E, _ptr(wv), E, E, _ptr(bv), _ptr(out2), B, E, E,
                                             hd, E, 0, 0, 0, 0, _stream()))
    torch.cuda.synchronize()
    ref2 = torch.einsum("bhc,hdc->bhd", xbar.float().view(B, heads, E), wv.float().view(heads, hd, E)).reshape(B, E) + bv
    _close(out2, ref2, "grouped v")


@pytest.mark.parametrize("B,H,W,Cin,Cout", [
    (2, 16, 16, 64, 64),     # box 16x8
    (3, 56, 56, 64, 64),     # layer1 conv2: box 56x2 / partial coverage
    (2, 28, 28, 128, 128),   # layer2
    (3, 14, 14, 256, 256),   # layer3: box 14x9, second tile half out of bounds
    (5, 7, 7, 512, 512),     # layer4: two images per tile, odd batch
    (2, 112, 112, 32, 32),   # stem conv2: BK 32 / BN 32
    (1, 112, 112, 32, 64),   # stem conv3
    (1, 10, 12, 32, 96),     # odd spatial size, N tile 32 x 3
])
def test_conv3x3(lib, B, H, W, Cin, Cout):
    _conv3x3_case(lib, B, H, W, Cin, Cout, pool=False)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [
    (2, 112, 112, 32, 64),   # stem conv3 + pool
    (3, 56, 56, 128, 128),   # layer2.0 conv2 + pool
    (3, 28, 28, 256, 256),   # layer3.0
    (5, 14, 14, 512, 512),   # layer4.0: whole images per tile
    (2, 12, 20, 64, 64),     # odd strip split
])
def test_conv3x3_fused_pool(lib, B, H, W, Cin, Cout):
    _conv3x3_case(lib, B, H, W, Cin, Cout, pool=True)


def _conv3x3_case(lib, B, H, W, Cin, Cout, pool):
    g = torch.Generator(device="cuda").manual_seed(H * 1000 + Cin)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).half()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * (9 * Cin) ** -0.5).half()
    b = torch.randn(Cout, device="cuda", generator=g)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    oshape = (B, H // 2, W // 2, Cout) if pool else (B, H, W, Cout)
    out = torch.full(oshape, float("nan"), device="cuda", dtype=torch.float16)
    _check(lib, lib.embclip_conv3x3_f16(_ptr(x), _ptr(wk), _ptr(b), _ptr(out), B, H, W, Cin, Cout, 1, int(pool), _stream()))
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).relu()
    if pool:
        ref = F.avg_pool2d(ref.half().float(), 2)     # the kernel rounds the full-resolution map to fp16 before averaging
    ref = ref.permute(0, 2, 3, 1)
    _close(out.reshape(-1, Cout), ref.reshape(-1, Cout), f"conv3x3 B{B} {H}x{W} {Cin}->{Cout} pool{pool}")


def test_avgpool2(lib):
    x = torch.randn(3, 28, 28, 128, device="cuda").half()
    out = torch.empty(3, 14, 14, 128, device="cuda", dtype=torch.float16)
    _check(lib, lib.embclip_avgpool2_f16(_ptr(x), _ptr(out), 3, 28, 28, 128, _stream()))
    torch.cuda.synchronize()
    ref = F.avg_pool2d(x.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert torch.equal(out, ref.half())     # fp32 sum of 4, one rounding: bit-exact


def test_stem_conv1(lib):
    torch.manual_seed(3)
    x = torch.randn(2, 224, 224, 3, device="cuda")
    w = torch.randn(32, 3, 3, 3, device="cuda") * 27 ** -0.5
    b = torch.randn(32, device="cuda")
    wk = w.permute(2, 3, 1, 0).reshape(27, 32).contiguous()
    out = torch.empty(2, 112, 112, 32, device="cuda", dtype=torch.float16)
    _check(lib, lib.embclip_stem_conv1(_ptr(x), _ptr(wk), _ptr(b), _ptr(out), 2, 224, 32, _stream()))
    torch.cuda.synchronize()
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, b, stride=2, padding=1).relu().permute(0, 2, 3, 1)
    _close(out.reshape(-1, 32), ref.reshape(-1, 32), "stem conv1")


def test_error_paths(lib):
    a = torch.zeros(4, 48, device="cuda", dtype=torch.float16)
    rc = lib.embclip_gemm_f16(_ptr(a), None, _ptr(a), None, None, _ptr(a), 4, 48, 48, 0, 0, 0, _stream())
    assert rc == -1 and b"multiples of 32" in lib.embclip_last_error()
    assert lib.embclip_gemm_f16(None, None, None, None, None, None, 4, 64, 64, 0, 0, 0, _stream()) == -1


def _tail_case(lib, M, n1, down, seed=0):
    """bneck_tail: out = relu([y2 | x0] W3^T + b3 (+ res)) rounded to fp16; y1 = relu(out W1^T + b1) from that fp16 tile."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    y2 = rn(M, 64).relu().half()
    x0 = rn(M, 64).relu().half() if down else None
    k3 = 128 if down else 64
    w3 = (rn(256, k3) * k3 ** -0.5).half()
    b3 = rn(256)
    res = None if down else rn(M, 256).relu().half()
    w1 = (rn(n1, 256) * 256 ** -0.5).half()
    b1 = rn(n1)
    out = torch.full((M, 256), float("nan"), device="cuda", dtype=torch.float16)
    y1 = torch.full((M, n1), float("nan"), device="cuda", dtype=torch.float16)
    _check(lib, lib.embclip_bneck_tail_f16(_ptr(y2), _ptr(x0), _ptr(w3), _ptr(b3), _ptr(res), _ptr(out), _ptr(w1), _ptr(b1),
                                           _ptr(y1), M, n1, _stream()))
    torch.cuda.synchronize()
    a = torch.cat([y2.float(), x0.float()], 1) if down else y2.float()
    ref_out = a @ w3.float().t() + b3
    if res is not None:
        ref_out = ref_out + res.float()
    ref_out = ref_out.relu()
    _close(out, ref_out, f"bneck_tail out M{M} n1 {n1} down{down}")
    ref_y1 = (out.float() @ w1.float().t() + b1).relu()          # fed with the kernel's own fp16 x' (its rounding point)
    _close(y1, ref_y1, f"bneck_tail y1 M{M} n1 {n1} down{down}")


@pytest.mark.parametrize("M,n1,down", [
    (128, 64, False),            # one tile
    (128, 64, True),             # K-concatenated downsample conv (layer1.0)
    (128, 128, False),           # layer1.2 -> layer2.0.conv1
    (3136, 64, False),           # one frame: partial last tile (3136 = 24.5 x 128)
    (3136 * 5, 64, True),
    (128 * 148 * 3 + 77, 128, False),   # three tiles per CTA + ragged tail: quarter-buffer recycling, accumulator parity
    (128 * 148 * 4, 64, False),
    (50, 64, True),              # fewer rows than one tile
])
def test_bneck_tail(lib, M, n1, down):
    _tail_case(lib, M, n1, down, seed=M % 97)


def _tail_pool_case(lib, B, H, W, mode, seed=0):
    """bneck_tail, pooled-output variant: the launch writes pool(x') (2x2 average, or the window's top-left pixel) and the next
    conv1 of x', but not x' itself.  The same launch without the pooling gives x' bit for bit (same tiles' arithmetic), so the
    pooled tensor must EQUAL the pool of that x', and y1 must equal the plain launch's y1."""
    import torch.nn.functional as F
    g = torch.Generator(device="cuda").manual_seed(seed)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    M, n1 = B * H * W, 128
    y2 = rn(M, 64).relu().half()
    w3 = (rn(256, 64) / 8).half()
    b3 = rn(256)
    res = rn(M, 256).relu().half()
    w1 = (rn(n1, 256) / 16).half()
    b1 = rn(n1)
    out = torch.full((M, 256), float("nan"), device="cuda", dtype=torch.float16)
    y1_plain = torch.full((M, n1), float("nan"), device="cuda", dtype=torch.float16)
    _check(lib, lib.embclip_bneck_tail_f16(_ptr(y2), None, _ptr(w3), _ptr(b3), _ptr(res), _ptr(out), _ptr(w1), _ptr(b1),
                                           _ptr(y1_plain), M, n1, _stream()))
    pooled = torch.full((M // 4, 256), float("nan"), device="cuda", dtype=torch.float16)
    y1 = torch.full((M, n1), float("nan"), device="cuda", dtype=torch.float16)
    _check(lib, lib.embclip_bneck_tail_pool_f16(_ptr(y2), _ptr(w3), _ptr(b3), _ptr(res), _ptr(pooled), mode, W, _ptr(w1), _ptr(b1),
                                                _ptr(y1), M, n1, _stream()))
    torch.cuda.synchronize()
    x = out.view(B, H, W, 256)
    if mode == 1:
        x = x.float()
        ref = ((((x[:, 0::2, 0::2] + x[:, 0::2, 1::2]) + x[:, 1::2, 0::2]) + x[:, 1::2, 1::2]) * 0.25).half()
    else:
        ref = x[:, 0::2, 0::2]
    assert torch.equal(pooled.view(B, H // 2, W // 2, 256), ref), f"pooled output differs (B{B} {H}x{W} mode {mode})"
    assert torch.equal(y1, y1_plain), f"conv1' output differs from the plain launch (B{B} {H}x{W} mode {mode})"
    ref_out = (y2.float() @ w3.float().t() + b3 + res.float()).relu()
    _close(out, ref_out, f"bneck_tail (plain, for the pooled case) out B{B} {H}x{W}")


@pytest.mark.parametrize("B,H,W,mode", [
    (1, 2, 56, 1),               # one tile of two image rows
    (2, 56, 56, 1),              # layer 1 of CLIP RN50: AvgPool2d(2) in front of layer2.0's downsample conv
    (2, 56, 56, 2),              # torchvision: stride-2 1x1 downsample conv reads x'[:, ::2, ::2]
    (160, 4, 56, 1),             # several tiles per CTA: accumulator / staging parities
    (3, 8, 32, 1),               # narrower rows: 64-row tiles
    (1, 6, 64, 2),               # the widest supported row: 128-row tiles
])
def test_bneck_tail_pool(lib, B, H, W, mode):
    _tail_pool_case(lib, B, H, W, mode, seed=B * 7 + W)


def test_bneck_tail_pool_rejects_bad_arguments(lib):
    z = torch.zeros(2 * 2 * 56, 256, device="cuda", dtype=torch.float16)
    b = torch.zeros(256, device="cuda")
    call = lambda mode, W, M, n1: lib.embclip_bneck_tail_pool_f16(_ptr(z), _ptr(z), _ptr(b), _ptr(z), _ptr(z), mode, W, _ptr(z), _ptr(b),
                                                                  _ptr(z), M, n1, _stream())
    assert call(1, 56, 224, 128) == 0
    assert call(3, 56, 224, 128) != 0        # max pooling is not a downsample-branch mode
    assert call(1, 57, 228, 128) != 0        # odd width
    assert call(1, 72, 144, 128) != 0        # two rows do not fit one 128-row tile
    assert call(1, 56, 168, 128) != 0        # M is not whole row pairs
    assert call(1, 56, 224, 64) != 0         # only the 128-wide next conv1 is built
    assert b"bneck_tail" in lib.embclip_last_error()
    torch.cuda.synchronize()


def _tail_stream_case(lib, M, n3, seed=0):
    """bneck_tail_stream: out = relu(y2 W3^T + b3 + res) rounded to fp16; y1 = relu(out W1^T + b1) from that fp16 tile."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    k3, n1 = 128, 128
    y2 = rn(M, k3).relu().half()
    w3 = (rn(n3, k3) * k3 ** -0.5).half()
    b3 = rn(n3)
    res = rn(M, n3).relu().half()
    w1 = (rn(n1, n3) * n3 ** -0.5).half()
    b1 = rn(n1)
    out = torch.full((M, n3), float("nan"), device="cuda", dtype=torch.float16)
    y1 = torch.full((M, n1), float("nan"), device="cuda", dtype=torch.float16)
    _check(lib, lib.embclip_bneck_tail_stream_f16(_ptr(y2), _ptr(w3), _ptr(b3), _ptr(res), _ptr(out), _ptr(w1), _ptr(b1), _ptr(y1),
                                                  M, k3, n3, n1, _stream()))
    torch.cuda.synchronize()
    ref_out = (y2.float() @ w3.float().t() + b3 + res.float()).relu()
    _close(out, ref_out, f"bneck_tail_stream out M{M} n3 {n3}")
    ref_y1 = (out.float() @ w1.float().t() + b1).relu()          # fed with the kernel's own fp16 x' (its rounding point)
    _close(y1, ref_y1, f"bneck_tail_stream y1 M{M} n3 {n3}")


@pytest.mark.parametrize("M,n3", [
    (128, 512),                  # one tile, 8 quarters: both rings wrap, lagged conv1' flush
    (128, 256),                  # 4 quarters (= ring depth)
    (784, 512),                  # one layer-2 frame: partial last tile
    (128 * 148 * 3 + 77, 512),   # three tiles per CTA + ragged tail: accumulator parity, A ring of two tiles
    (128 * 148 * 2, 1024),       # 16 quarters per tile
    (50, 512),                   # fewer rows than one tile
])
def test_bneck_tail_stream(lib, M, n3):
    _tail_stream_case(lib, M, n3, seed=M % 97)


def test_bneck_tail_rejects_bad_arguments(lib):
    z = torch.zeros(128, 256, device="cuda", dtype=torch.float16)
    b = torch.zeros(256, device="cuda")
    args = lambda x0, res, n1: (_ptr(z), x0, _ptr(z), _ptr(b), res, _ptr(z), _ptr(z), _ptr(b), _ptr(z), 128, n1, _stream())
    assert lib.embclip_bneck_tail_f16(*args(None, None, 64)) != 0            # neither downsample source nor residual
    assert lib.embclip_bneck_tail_f16(*args(_ptr(z), _ptr(z), 64)) != 0      # both
    assert lib.embclip_bneck_tail_f16(*args(None, _ptr(z), 96)) != 0         # unsupported conv1 width
    assert b"bneck_tail" in lib.embclip_last_error()

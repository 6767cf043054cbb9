"""How long is one link of a dependent chain of small launches?  (rollout-sized batches: the forward is 53 dependent kernels)"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embclip_b200 import _lib  # noqa: E402

lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
dev = "cuda"


def chain(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def chain_graph(fn, n=200):
    """the same chain replayed from a CUDA graph: no host cost per launch -> the device-side cost of one link"""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(5):
            fn_s = fn
            fn_s()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * n) * 1e3


def gemm(M, N, K, res=False):
    a = torch.randn(M, K, device=dev).half()
    w = torch.randn(N, K, device=dev).half()
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev).half() if res else None
    # ping-pong so every launch depends on the previous one's output when shapes allow (N == K)
    o = torch.empty(M, N, device=dev, dtype=torch.float16)
    return lambda: lib.embclip_gemm_f16(a.data_ptr(), None, w.data_ptr(), b.data_ptr(), r.data_ptr() if res else None, o.data_ptr(), M, N, K, 0, 1, 0,
                                        torch.cuda.current_stream().cuda_stream)


def conv3(B, H, Cc):
    x = torch.randn(B, H, H, Cc, device=dev).half()
    w = torch.randn(Cc, 9 * Cc, device=dev).half()
    b = torch.randn(Cc, device=dev)
    o = torch.empty(B, H, H, Cc, device=dev, dtype=torch.float16)
    return lambda: lib.embclip_conv3x3_f16(x.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), B, H, H, Cc, Cc, 1, 0,
                                           torch.cuda.current_stream().cuda_stream)


for name, fn in [("gemm 128x128x64 (1 tile, 1 k-block)", gemm(128, 128, 64)),
                 ("gemm 128x128x1024 (1 tile, 16 k-blocks)", gemm(128, 128, 1024)),
                 ("gemm 1568x256x1024 (l3 conv1 @B=8)", gemm(1568, 256, 1024)),
                 ("gemm 1568x1024x256+res (l3 conv3 @B=8)", gemm(1568, 1024, 256, res=True)),
                 ("gemm 392x512x2048 (l4 conv1 @B=8)", gemm(392, 512, 2048)),
                 ("gemm 392x2048x512+res (l4 conv3 @B=8)", gemm(392, 2048, 512, res=True)),
                 ("conv3x3 8x14x14x256 (l3 conv2 @B=8)", conv3(8, 14, 256)),
                 ("conv3x3 8x7x7x512 (l4 conv2 @B=8)", conv3(8, 7, 512)),
                 ("conv3x3 8x28x28x128 (l2 conv2 @B=8)", conv3(8, 28, 128))]:
    print(f"{name:45s} {chain(fn):7.2f} us per launch from the host (PDL), {chain_graph(fn):7.2f} us replayed from a CUDA graph")
z = torch.zeros(1024, device=dev)
print(f"{'torch z.add_(1) (1 block)':45s} {chain(lambda: z.add_(1)):7.2f} us per launch from the host, {chain_graph(lambda: z.add_(1)):7.2f} us from a graph")

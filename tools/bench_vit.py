#!/usr/bin/env python
"""CUDA-event timing of th_vit_attention against torch's eager attention (the reference's formula) and torch SDPA
at the three token counts (B = 3 views, 3 heads of 64).  One JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def timed(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    import __graft_entry__ as entry
    entry.build()
    from transhuman_b200 import ops
    dev, B, H, D = "cuda:0", 3, 3, 64
    scale = D ** -0.5
    res = {}
    for N in (300, 1500, 6000):
        qkv = torch.randn((B, N, 3 * H * D), device=dev)

        def eager():
            q, k, v = qkv.reshape(B, N, 3, H, D).permute(2, 0, 3, 1, 4)
            a = ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
            return (a @ v).transpose(1, 2).reshape(B, N, H * D)

        def sdpa():
            q, k, v = qkv.reshape(B, N, 3, H, D).permute(2, 0, 3, 1, 4)
            return torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, H * D)

        ours = timed(lambda: ops.vit_attention(qkv, H, scale))
        flops = 4.0 * B * H * N * N * D
        res[str(N)] = {"ours_ms": ours, "ours_algorithmic_tflops": flops / ours / 1e9,
                       "torch_eager_fp32_ms": timed(eager), "torch_sdpa_fp32_ms": timed(sdpa),
                       "max_abs_vs_eager": float((ops.vit_attention(qkv, H, scale) - eager()).abs().max())}
    print(json.dumps(res))


if __name__ == "__main__":
    main()

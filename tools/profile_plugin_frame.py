#!/usr/bin/env python
"""One render_fast(batch) frame through the plugin with the genuine Network (300 tokens by default), after warm-up,
bracketed by cudaProfilerStart/Stop -- for `ncu --profile-from-start off` launch lists of a whole frame."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def main():
    import __graft_entry__ as entry
    entry.build()
    from tools import bench_plugin
    n_tok = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    renderer, batch = bench_plugin.make(n_tok, "cuda:0")
    with torch.no_grad():
        for _ in range(3):
            renderer.render_fast(dict(batch))
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        renderer.render_fast(dict(batch))
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()

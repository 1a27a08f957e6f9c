#!/usr/bin/env python
"""A/B of the token-grid K-NN (TH_TOKEN_GRID threshold) on the configs with many tokens: run as
`TH_TOKEN_GRID=1024 python tools/ab_token_grid.py` vs `TH_TOKEN_GRID=1000000 ...`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench


class Args:
    pass


def main():
    import __graft_entry__ as entry
    entry.build()
    out = bench.other_configs(Args(), torch.device("cuda:0"))
    c3, c5 = out["c3_512x512x128_1500tok"], out["c5_grid256_6000tok"]
    print(json.dumps({"TH_TOKEN_GRID": os.environ.get("TH_TOKEN_GRID"), "c3_culled_ms": c3["culled"]["ms_per_frame"],
                      "c3_dense_ms": c3["dense"]["ms_per_frame"], "c5_ms": c5["ms"]}))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-step device time of 40 consecutive dense C2 frames from a cold start, next to the SM clock / power that
nvidia-smi reports -- shows how long the power-cap transient lasts after load begins (bench.py's warm-up policy)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench


class Args:
    size, samples, tokens, views, simt, plain_maps = 512, 64, 300, 3, False, False


def main():
    import __graft_entry__ as entry
    entry.build()
    from transhuman_b200 import ops
    dev = torch.device("cuda:0")
    fr, frame, host_rays = bench.build_workload(Args, 0, dev)
    rays = tuple(r.to(dev) for r in host_rays)
    smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader,nounits",
                            "-lms", "200"], stdout=subprocess.PIPE, text=True)
    ops.render_rays(frame, *rays, 64)          # allocation / first-touch
    torch.cuda.synchronize()
    time.sleep(3.0)                            # idle: let the GPU cool / clock down like a fresh bench start
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
    ev[0].record()
    for i in range(40):
        ops.render_rays(frame, *rays, 64)
        ev[i + 1].record()
    torch.cuda.synchronize()
    smi.terminate()
    rows = [l.strip() for l in smi.stdout.read().splitlines() if l.strip()]
    print(json.dumps({"ms_per_step": [round(ev[i].elapsed_time(ev[i + 1]), 1) for i in range(40)],
                      "smi_clock_power_temp_every_200ms": rows[-60:]}))


if __name__ == "__main__":
    main()

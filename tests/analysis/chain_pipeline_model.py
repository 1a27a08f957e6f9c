#!/usr/bin/env python
"""A small event model of the chain kernel's job pipeline (csrc/mlp_chain.cu), for PLANNING only.

It walks the real job program (``th_debug_chain_program``) through four in-order roles -- loader (3-stage ring,
operand bandwidth, one fence per dependency), MMA issuer (tensor floor per k-block, TMEM reuse waits), epilogue
(measured cost per epilogue kind) and the attention mix -- and reports the steady-state cycles per 256-point
unit.  The constants below are the ones measured on the B200 with TH_CHAIN_STATS / ncu (profiles/README.md);
with them the model reproduces the measured unit (313 kcycles) and the shape of the per-job waits, so it can rank
ideas before GPU time is spent on them.  Numbers printed here are model output, not measurements."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.test_cabi_host import _pack  # noqa: E402
from tests.test_chain_program import EPI_ALPHA, EPI_IMG, EPI_KEEP, EPI_RGB, EPI_SCORES, _program  # noqa: E402
from transhuman_b200 import _lib, synth  # noqa: E402

K = dict(
    mma_kb_256=1536.0,       # 12 MMAs x 128 cycles (M = 256, N = 256, K = 16 on a CTA pair)
    mma_kb_128=768.0,
    epi_img_256=6200.0,      # image epilogue of an N = 256 job (TH_CHAIN_STATS: 1353-1450 kcycles / 15 units / 15 jobs)
    epi_scores=4400.0,
    epi_head=6500.0,
    epi_keep=200.0,
    mix_total=46000.0,       # in-place attention mix per unit (4 k-blocks)
    fence=1500.0,            # gpu-scope + proxy fence on the loader thread
    lat_l2=1300.0,           # bulk copy issue -> bytes landed, operand in L2
    lat_hbm=2300.0,          # ... chunk input streamed from HBM
    ingest=45.0,             # bytes per cycle an SM takes in for operands (SM <-> L2 port under load)
    issue=280.0,             # loader instructions per k-block
    stages=3,
    # calibration (not measured one by one): hand-over latencies that the wait-time profile implies
    hop_epi=1300.0,          # per job on the epilogue warps: named barrier, bias staging, accumulator-full wake-up
    hop_dep=900.0,           # counter release -> the loader's poll (nanosleep) sees it
    hop_full=350.0,          # bytes landed -> the MMA thread's mbarrier wait returns
    hop_commit=600.0,        # last MMA issued -> commit arrives (stage free / accumulator full)
)


def simulate(jobs, has_mix, units=12, k=K, drop_mix_wait=False):
    n = len(jobs)
    last_scores = max((j for j, jb in enumerate(jobs) if jb["epi"] == EPI_SCORES), default=-1)
    epi_done, mma_done = {}, {}
    stage_free = [0.0] * k["stages"]
    loader_t = mma_t = epi_t = 0.0
    ring = 0
    port_free = 0.0
    unit_end = []
    waits = np.zeros((n, 2))
    for u in range(units):
        mix_done = [None] * 4
        fenced = set()
        for j, jb in enumerate(jobs):
            G = u * n + j
            # ---- loader: every k-block of the job, in segment order
            arrive = []
            for sg in jb["segs"]:
                for kb in range(sg["kbs"]):
                    t = loader_t
                    if not sg["chunk"] and sg["dep"] >= 0 and ("j", sg["dep"]) not in fenced:
                        t = max(t, epi_done[u * n + sg["dep"]] + k["hop_dep"]) + k["fence"]
                        fenced.add(("j", sg["dep"]))
                    if sg["dep_mix"] and ("m", kb) not in fenced:
                        t = max(t, mix_done[kb] + k["hop_dep"] if not drop_mix_wait else t) + k["fence"]
                        fenced.add(("m", kb))
                    t = max(t, stage_free[ring % k["stages"]]) + k["issue"]
                    nbytes = 32768 + jb["N"] * 128
                    port_free = max(port_free, t) + nbytes / k["ingest"]
                    arrive.append(max(t + (k["lat_hbm"] if sg["chunk"] else k["lat_l2"]), port_free))
                    loader_t = t
                    ring += 1
            # ---- MMA issuer
            t0 = mma_t
            if G - jb["wait_back"] >= 0:
                t0 = max(t0, epi_done.get(G - jb["wait_back"], 0.0))
            waits[j, 0] += t0 - mma_t
            t = t0
            per_kb = k["mma_kb_256"] if jb["N"] == 256 else k["mma_kb_128"]
            base_ring = ring - len(arrive)
            for i, a in enumerate(arrive):
                a += k["hop_full"]
                waits[j, 1] += max(0.0, a - t)
                t = max(t, a) + per_kb
                stage_free[(base_ring + i) % k["stages"]] = t + k["hop_commit"]
            mma_t = mma_done[G] = t
            # ---- epilogue
            e0 = max(epi_t, t + k["hop_commit"]) + k["hop_epi"]
            dur = {EPI_IMG: k["epi_img_256"] * jb["N"] / 256.0, EPI_KEEP: k["epi_keep"], EPI_SCORES: k["epi_scores"],
                   EPI_ALPHA: k["epi_head"], EPI_RGB: k["epi_head"]}[jb["epi"]]
            epi_t = epi_done[G] = e0 + dur
            if j == last_scores and has_mix:
                for kb in range(4):
                    mix_done[kb] = epi_t + k["mix_total"] * (kb + 1) / 4.0
        unit_end.append(mma_t)
    per_unit = (unit_end[-1] - unit_end[3]) / (len(unit_end) - 4)
    return per_unit, waits / units


def main():
    lib = _lib.load()
    blob = _pack(lib, synth.make_weights(seed=9), 3)
    head, jobs = _program(lib, blob, 3, 256, 0, 0)
    base, waits = simulate(jobs, head["has_mix"])
    floor = sum(jb["nkb"] * (K["mma_kb_256"] if jb["N"] == 256 else K["mma_kb_128"]) for jb in jobs)
    print(f"default program: {len(jobs)} jobs, tensor floor {floor / 1e3:.0f} kcycles per unit; model {base / 1e3:.0f} "
          f"(measured 313); MMA waits per unit: TMEM {waits[:, 0].sum() / 1e3:.0f}, operands {waits[:, 1].sum() / 1e3:.0f} "
          f"(measured 43 / 89)")
    print("  per job (TMEM / operand wait, kcycles): " +
          " ".join(f"j{j}={waits[j, 0] / 1e3:.1f}/{waits[j, 1] / 1e3:.1f}" for j in range(len(jobs))))

    def report(label, per_unit):
        print(f"  {label:68s} {per_unit / 1e3:6.0f} kcycles per unit  ({100 * (per_unit / base - 1):+5.1f} %)")

    print("what-if (model):")
    for name, kk in (("image epilogue 20 % faster (bias through the MMA, fused ReLU)", dict(K, epi_img_256=K["epi_img_256"] * 0.8)),
                     ("image epilogue 35 % faster", dict(K, epi_img_256=K["epi_img_256"] * 0.65)),
                     ("attention mix 2x faster", dict(K, mix_total=K["mix_total"] / 2)),
                     ("no fences at all", dict(K, fence=0.0)),
                     ("4-stage ring", dict(K, stages=4)),
                     ("operand ingest 64 B/clk instead of 45", dict(K, ingest=64.0))):
        report(name, simulate(jobs, head["has_mix"], k=kk)[0])
    report("mix entirely off the critical path", simulate(jobs, head["has_mix"], drop_mix_wait=True)[0])
    h2, j2 = _program(lib, blob, 3, 256, 0, 1)
    report("pre-mapped feature maps, X copied through identity jobs (in tree, flag)", simulate(j2, h2["has_mix"])[0])
    # pre-mapped without the copies: the X jobs disappear, their readers take the chunk image
    j3 = []
    x_jobs = [j for j, jb in enumerate(j2) if jb["epi"] == EPI_IMG and jb["segs"][0]["chunk"] and jb["nkb"] == 4
              and jb["out_off"] >= 3 * h2["scr_act"]][:3]
    remap = {}
    for j, jb in enumerate(j2):
        if j in x_jobs:
            continue
        remap[j] = len(j3)
        j3.append(jb)
    import copy
    j3 = copy.deepcopy(j3)
    for jb in j3:
        for sg in jb["segs"]:
            if not sg["chunk"] and sg["dep"] in x_jobs:
                sg["chunk"], sg["dep"] = 1, -1
            elif not sg["chunk"] and sg["dep"] >= 0:
                sg["dep"] = remap[sg["dep"]]
    report("pre-mapped, no copies (mix and key embeds read X from the chunk image)", simulate(j3, h2["has_mix"])[0])
    kk = dict(K, epi_img_256=K["epi_img_256"] * 0.8)
    report("  ... and the 20 % faster image epilogue", simulate(j3, h2["has_mix"], k=kk)[0])
    report("  ... and the mix off the critical path", simulate(j3, h2["has_mix"], k=kk, drop_mix_wait=True)[0])


if __name__ == "__main__":
    main()

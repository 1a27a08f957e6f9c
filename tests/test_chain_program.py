"""CPU check of the job program the layer-chained kernel runs (csrc/mlp_chain.cu).

``th_debug_chain_program`` returns the program ``mlp_forward_chain`` builds -- jobs, their operand segments
(scratch slots / chunk images), weight k-blocks, epilogue kinds, TMEM columns -- without touching a device.
This test (1) checks its static invariants: segment and weight k-block counts add up, every scratch operand
was stored by the job it names, TMEM columns are never overwritten while an epilogue may still read them;
(2) INTERPRETS the program in float64 with the matrices of the packed blob and compares the result with the
oracle's MLP.  It covers the default program, the density-only program and the experimental pre-mapped one,
so a wiring mistake in a new program variant shows up here before it costs GPU time."""
import ctypes as C
import struct

import numpy as np
import pytest
import torch

from tests.test_cabi_host import _matrices, _pack, lib  # noqa: F401  (fixture)
from transhuman_b200 import synth

EPI_IMG, EPI_KEEP, EPI_SCORES, EPI_ALPHA, EPI_RGB, EPI_MIX = 0, 2, 3, 4, 5, 6
MAX_SEG = 5
IMG_NAMES = ["fc0", "ar0", "k0", "k1", "v", "fc1", "fc2", "fc3m", "f", "view", "t", "fc1f", "gvf"]


def _program(lib, blob, V, P, alpha_only, premapped):
    cap = 8 + 32 * (16 + 6 * MAX_SEG)
    table = np.zeros(cap, dtype=np.int64)
    n = lib.th_debug_chain_program(blob.ctypes.data, V, P, alpha_only, premapped, table.ctypes.data, cap)
    assert n > 0, n
    head = table[:8]
    jobs = []
    for j in range(int(head[0])):
        t = table[8 + j * (16 + 6 * MAX_SEG):8 + (j + 1) * (16 + 6 * MAX_SEG)]
        keys = ["N", "relu", "epi", "out_off", "tmem_col", "wait_back", "view", "nseg", "nkb", "wimg", "bias", "bias2",
                "reader", "shift", "out_aset"]
        job = {k: int(t[i]) for i, k in enumerate(keys)}
        job["segs"] = [dict(zip(["chunk", "off", "tile_off", "kbs", "dep", "dep_mix"], map(int, t[16 + 6 * s:22 + 6 * s])))
                       for s in range(job["nseg"])]
        for sg in job["segs"]:
            sg["aset"], sg["dep_mix"] = sg["dep_mix"] >> 1, sg["dep_mix"] & 1
        jobs.append(job)
    return {"njobs": int(head[0]), "V": int(head[1]), "has_mix": int(head[2]), "alpha_only": int(head[3]),
            "Pp": int(head[4]), "scr_act": int(head[5]), "tile_img": int(head[6]), "deferred": int(head[7])}, jobs


def _images(blob, V):
    offs = struct.unpack_from("<54Q", blob, 16)
    shapes = {"fc0": (256, 256), "ar0": (256, 384), "k0": (128, 256), "k1": (128, 256), "v": (256, 512),
              "fc1": (256, 256), "fc2": (256, 256), "fc3m": (256, 256 * V), "f": (256, 640), "view": (128, 320),
              "t": (128, 128 * V + 384), "fc1f": (256, 512), "gvf": (128, 704), "gvfp": (128, 448),
              "tp": (128, 128 * V + 128), "xid": (256, 256)}
    img = {n: offs[30 + i] for i, n in enumerate(IMG_NAMES)}
    img.update({"gvfp": offs[51], "tp": offs[52], "xid": offs[53]})
    # fc_1' cut into S / X parts and halves of 128 rows (after img_off[24] / img_inv_scale[24] / n_img in the header)
    tail = struct.unpack_from("<12Q", blob, 16 + 57 * 8 + 24 * 8 + 24 * 4 + 8)
    for i, n in enumerate(["fc1s0", "fc1s1", "fc1x0", "fc1x1"]):
        img[n] = tail[8 + i]
        shapes[n] = (128, 256)
    return {n: (o, shapes[n]) for n, o in img.items()}


def _executed(head, jobs, n_units=4):
    """The job sequence one cluster executes for n_units units: iteration `it` runs the shift-0 jobs of unit `it` and
    the shift -1 (deferred tail) jobs of unit `it - 1`; a deferred tail needs one draining iteration."""
    n_iter = n_units + (1 if head["deferred"] else 0)
    return [(j, it, it + jb["shift"]) for it in range(n_iter) for j, jb in enumerate(jobs)
            if 0 <= it + jb["shift"] < n_units]


def _check_tmem(head, jobs):
    """No job's MMAs may overwrite TMEM columns that an earlier epilogue can still read: executed job number G starts
    after the epilogue of executed job number G - wait_back; a kept accumulator (key embeds, Y_i of the TMEM-side
    mix) is read until the last score / mix epilogue of its unit."""
    n = head["njobs"]
    flip_on = 256 if n & 1 else 0
    seq = _executed(head, jobs)
    where = {(j, ui): g for g, (j, it, ui) in enumerate(seq)}
    last_reader = {}
    for j, jb in enumerate(jobs):
        last = j
        if jb["epi"] == EPI_KEEP:   # independent of the builder's bookkeeping: who reads a kept accumulator?
            k = j + 1
            while jobs[k]["epi"] == EPI_KEEP:
                k += 1
            kind = jobs[k]["epi"]                             # key embeds -> scores, Y_i -> the mix epilogues
            assert kind in (EPI_SCORES, EPI_MIX)
            while k + 1 < n and jobs[k + 1]["epi"] == kind:
                k += 1
            last = k                                          # the last job of that contiguous run
            assert jb["reader"] == last and jobs[last]["shift"] == jb["shift"], (j, jb["reader"], last)
        last_reader[j] = last
    for g, (j, it, ui) in enumerate(seq):
        c0 = (jobs[j]["tmem_col"] + (flip_on if it & 1 else 0)) & 511
        c1 = c0 + jobs[j]["N"]
        assert c1 <= 512
        for a in range(g):
            ja, ita, uia = seq[a]
            a0 = (jobs[ja]["tmem_col"] + (flip_on if ita & 1 else 0)) & 511
            if a0 < c1 and c0 < a0 + jobs[ja]["N"]:
                reader = where[(last_reader[ja], uia)]
                assert g - jobs[j]["wait_back"] >= reader, \
                    f"executed job {g} (job {j}, wait_back {jobs[j]['wait_back']}) overwrites columns job {ja} is read from until {reader}"


def _check_scratch(head, jobs):
    """Scratch tiles across units (deferred tail): a tile may only be overwritten by an epilogue when every job that
    reads the old contents has been ISSUED earlier (MMAs complete in order, and an epilogue stores after its own job's
    MMAs).  Slot set A alternates with the unit's parity, slot B does not."""
    V, scr = head["V"], head["scr_act"]
    seq = _executed(head, jobs)

    def tiles(off, aset, ui, nbytes):
        base = off + (V * scr if (aset and ui & 1) else 0)
        return set(range(base // head["tile_img"], (base + nbytes + head["tile_img"] - 1) // head["tile_img"]))

    writes, reads = [], []          # (exec index, tiles)
    for g, (j, it, ui) in enumerate(seq):
        jb = jobs[j]
        for sg in jb["segs"]:
            if not sg["chunk"]:
                reads.append((g, tiles(sg["off"], sg["aset"], ui, sg["kbs"] * head["tile_img"]), j))
        if jb["epi"] in (EPI_IMG, EPI_MIX):
            writes.append((g, tiles(jb["out_off"], jb["out_aset"], ui, jb["N"] // 64 * head["tile_img"]), j, ui))
    for gw, tw, jw, uiw in writes:
        # readers of the PREVIOUS contents: reads executed after the previous write of the tile and before this one must
        # all precede gw in issue order (trivially true) -- the hazard is a reader issued AFTER gw that wants the OLD data
        for gr, tr, jr in reads:
            if gr > gw and tr & tw:
                # it must be a legitimate consumer of THIS write: its segment names jw as producer, or re-reads in place
                deps = {sg["dep"] for sg in jobs[jr]["segs"] if not sg["chunk"]}
                later_writer = any(g2 > gw and g2 < gr and (t2 & tr & tw) for g2, t2, _, _ in writes)
                assert jw in deps or later_writer or -1 in deps, (jw, jr, sorted(tr & tw)[:3])


def _slot(head, off):
    """(slot key, first k-block) of a scratch offset.  Slot A(v) at v * scr; slot B(v) behind it -- at full tile size,
    or, in the deferred-tail layout, at half size behind BOTH parity copies of slot set A."""
    V, scr = head["V"], head["scr_act"]
    if off < V * scr:
        key, inside = ("A", off // scr), off % scr
    elif head["deferred"]:
        assert off >= 2 * V * scr
        key, inside = ("B", (off - 2 * V * scr) // (scr // 2)), (off - 2 * V * scr) % (scr // 2)
    else:
        key, inside = ("B", (off - V * scr) // scr), (off - V * scr) % scr
    assert inside % head["tile_img"] == 0
    return key, inside // head["tile_img"]


def _interpret(blob, head, jobs, data):
    """data: chunk buffers by name, each (views or 1, P, C) float64.  Returns raw (P,4) or alpha (P)."""
    V, Pp, scr = head["V"], head["Pp"], head["scr_act"]
    mats = _matrices(blob, V)
    images = _images(blob, V)
    f32 = lambda off, n: torch.from_numpy(blob[off:off + 4 * n].view(np.float32).astype(np.float64))
    # chunk block (mlp_carve): byte offsets from the first buffer
    o_pix = V * Pp * 256 * 4
    o_pm = o_pix + V * Pp * 384 * 4 + 4 * V * Pp * 256 * 4 + 2 * V * Pp * 128 * 4
    chunk = {0: "rep", o_pix: "pix", o_pix + V * Pp * 256 * 4: "p2", o_pm: "pix_mean", o_pm + Pp * 384 * 4: "vd"}
    slots, written_by = {}, {}
    kept, scores, alpha, out_final, Aw = {}, {}, None, None, None
    data = dict(data)
    mixed_in_chunk = False
    # one unit: its shift-0 jobs in program order, then its deferred (shift -1) jobs, as the kernel runs them
    order = [j for j, jb in enumerate(jobs) if jb["shift"] == 0] + [j for j, jb in enumerate(jobs) if jb["shift"] == -1]
    for j in order:
        jb = jobs[j]
        parts = []
        for sg in jb["segs"]:
            C_ = sg["kbs"] * 64
            if sg["chunk"]:
                name = chunk[sg["off"]]
                view = sg["tile_off"] // (Pp // 128)
                assert sg["tile_off"] % (Pp // 128) == 0 and sg["dep"] == -1
                assert not sg["dep_mix"] or (name == "pix" and mixed_in_chunk), (j, name)
                buf = data[name]
                assert buf.shape[-1] == C_, (j, name, buf.shape, C_)
                parts.append(buf[view if buf.shape[0] > 1 else 0])
            else:
                slot, kb0 = _slot(head, sg["off"])
                assert slot in slots, f"job {j} reads scratch slot {slot} before anything stored it"
                assert bool(sg["aset"]) == bool(head["deferred"] and slot[0] == "A")
                if sg["dep"] >= 0:
                    for kb in range(kb0, kb0 + sg["kbs"]):
                        assert written_by[(slot, kb)] == sg["dep"], (j, slot, kb, written_by[(slot, kb)], sg["dep"])
                assert bool(sg["dep_mix"]) == (head["has_mix"] == 1 and slot[0] == "B" and
                                               jobs[written_by[(slot, kb0)]]["epi"] == EPI_IMG
                                               and jb["epi"] == EPI_IMG and jb["N"] == 256 and len(jb["segs"]) == 2)
                parts.append(slots[slot][:, 64 * kb0:64 * kb0 + C_])
        A = torch.cat(parts, -1)
        assert A.shape[1] == jb["nkb"] * 64
        name = next(n for n, (o, (N, K)) in images.items() if o <= jb["wimg"] < o + N * K * 4)
        o, (N, K) = images[name]
        assert N == jb["N"] and (jb["wimg"] - o) % (N * 256) == 0
        kb0 = (jb["wimg"] - o) // (N * 256)
        W = mats[name][0][:, kb0 * 64:kb0 * 64 + A.shape[1]]
        assert W.shape[1] == A.shape[1], (j, name, kb0)
        out = A @ W.T
        bias = f32(jb["bias"], N) if jb["bias"] >= 0 else None
        if jb["epi"] in (EPI_IMG, EPI_MIX):
            out = out + bias if bias is not None else out
            if jb["epi"] == EPI_MIX:   # + sum_i A[i][j] Y_i, Y_i = the kept accumulators of the X part
                assert Aw is not None and len(kept) == V
                out = out + sum(Aw[:, i, jb["view"], None] * kept[i] for i in range(V))
            out = torch.relu(out) if jb["relu"] else out
            slot, kb0 = _slot(head, jb["out_off"])
            assert 64 * kb0 + N <= 256 and bool(jb["out_aset"]) == bool(head["deferred"] and slot[0] == "A")
            if slot not in slots or jb["epi"] == EPI_IMG:
                slots[slot] = torch.zeros((out.shape[0], 256), dtype=torch.float64)   # a whole new tile
            slots[slot][:, 64 * kb0:64 * kb0 + N] = out
            for kb in range(kb0, kb0 + N // 64):
                written_by[(slot, kb)] = j
        elif jb["epi"] == EPI_KEEP:
            assert bias is None
            kept[jb["view"]] = out
        elif jb["epi"] == EPI_SCORES:
            kp = out + bias
            b2 = f32(jb["bias2"], 128)
            for jv in range(V):
                scores[(jb["view"], jv)] = (kp * (kept[jv] + b2)).sum(-1) / np.sqrt(128.0)
            if jb["view"] == V - 1:
                S_ = torch.stack([torch.stack([scores[(i, jv)] for jv in range(V)], -1) for i in range(V)], 1)  # (P,i,j)
                Aw = torch.softmax(S_, dim=1)
                kept = {}                                        # the key embeds are dead once the scores exist
                if head["has_mix"] == 1 and (("B", 0) in slots):
                    X = torch.stack([slots[("B", i)] for i in range(V)], 0)                   # (i,P,256)
                    XT = torch.einsum("pij,ipc->jpc", Aw, X)
                    for jv in range(V):
                        slots[("B", jv)] = XT[jv]
                elif head["has_mix"] == 1:  # no X jobs: the mix works on the X tiles of the chunk image
                    data["pix"] = torch.einsum("pij,ipc->jpc", Aw, data["pix"])
                    mixed_in_chunk = True
        elif jb["epi"] == EPI_ALPHA:
            O = torch.relu(out + bias)
            alpha = O @ mats["afc"][0][0] + mats["afc"][1][0]
        elif jb["epi"] == EPI_RGB:
            T = torch.relu(out + bias)
            out_final = T @ mats["rgb"][0].T + mats["rgb"][1]
        else:
            raise AssertionError(f"job {j}: unknown epilogue {jb['epi']}")
    return alpha if head["alpha_only"] else torch.cat([out_final, alpha[:, None]], -1)


@pytest.mark.parametrize("V", [1, 2, 3])
@pytest.mark.parametrize("variant", ["default", "alpha_only", "premapped", "premapped_alpha_only", "premapped_tmem",
                                     "premapped_tmem_alpha_only", "premapped_defer", "premapped_defer_alpha_only",
                                     "premapped_deep", "premapped_deep_alpha_only"])
def test_chain_program(lib, V, variant, monkeypatch):
    monkeypatch.setenv("TH_CHAIN_DEFER", "1" if "defer" in variant else "2" if "deep" in variant else "0")
    from oracle import transhuman_oracle as orc
    w = synth.make_weights(seed=9)
    blob = _pack(lib, w, V)
    P = 200
    premapped, alpha_only = variant.startswith("premapped"), variant.endswith("alpha_only")
    # th_debug_chain_program's `premapped`: 2 = TMEM-side mix, 3 = in-place mix by the mix warps
    head, jobs = _program(lib, blob, V, P, int(alpha_only), (2 if "tmem" in variant else 3) if premapped else 0)
    assert head["V"] == V and head["Pp"] == 256 and head["njobs"] == len(jobs) <= 32
    if "tmem" in variant:    # mix on the accumulator side: no mix warps, X never rewritten
        assert head["has_mix"] == 0 and sum(jb["epi"] == EPI_MIX for jb in jobs) == 2 * V
    elif premapped:
        assert head["has_mix"] == 1 and not any(jb["epi"] == EPI_MIX for jb in jobs)
    _check_tmem(head, jobs)
    _check_scratch(head, jobs)
    if "deep" in variant:                                     # everything behind the scores runs one unit late
        assert head["deferred"] == 1
        assert [jb["shift"] for jb in jobs] == [0] * (3 * V) + [-1] * (2 * V + 1 if alpha_only else 3 * V + 2)
    elif "defer" not in variant:
        assert head["deferred"] == 0 and all(jb["shift"] == 0 for jb in jobs)
    else:                                                     # opt-in: the tail of the network deferred by one unit
        assert head["deferred"] == 1 and sum(jb["shift"] == -1 for jb in jobs) == (1 if alpha_only else V + 2)

    g = torch.Generator().manual_seed(4)
    rep = torch.randn((V, 255, P), generator=g, dtype=torch.float64)
    pix = torch.randn((V, 384, P), generator=g, dtype=torch.float64)
    vd = torch.randn((1, P, 27), generator=g, dtype=torch.float64) * (0.0 if alpha_only else 1.0)
    w64 = {k: torch.from_numpy(np.asarray(v)).double() for k, v in w.items()}
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        want = orc.mlp_forward(w64, rep, pix, vd, progressive=False)[0]
    finally:
        torch.set_default_dtype(old)
    mats = _matrices(blob, V)
    pix_r = pix.permute(0, 2, 1)
    data = {"rep": torch.cat([rep.permute(0, 2, 1), torch.zeros((V, P, 1), dtype=torch.float64)], -1),
            "vd": torch.cat([vd[0], torch.zeros((P, 37), dtype=torch.float64)], -1)[None]}
    if premapped:  # what k_features PRE blends out of the pre-mapped maps
        pre = pix_r @ mats["pre"][0].T + mats["pre"][1]
        data["pix"] = torch.relu(pre[..., :256])
        data["p2"] = pre[..., 256:384]
        data["pix_mean"] = pre[..., 384:].sum(0)[None]
    else:
        data["pix"] = pix_r
        data["pix_mean"] = pix_r.mean(0)[None]
    got = _interpret(blob, head, jobs, data)
    ref = want[:, 3] if alpha_only else want
    assert (got - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())

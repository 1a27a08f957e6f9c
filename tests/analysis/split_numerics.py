#!/usr/bin/env python
"""CPU model of the tensor-core path's arithmetic ("Precision of the tensor path"):
every GEMM operand is split x = hi + lo with hi = fp16(x), lo = fp16(x - hi) and the products
A_hi B_hi + A_lo B_hi + A_hi B_lo are accumulated in fp32.  This script evaluates the chain kernel's layer
program from the packed blob (the same program tests/test_cabi_host.py checks in float64) with that operand
rounding -- accumulation in float64, so it isolates the OPERAND precision -- on the samples of a small
synthetic frame, composites the rays and prints the deviation from the oracle for 3, 2 and 1 products and
for single-pass TF32 / BF16 operands.  It answers "are three products needed for the 1e-4 RGB bar?"."""
import os
import struct
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402

from oracle import transhuman_oracle as orc  # noqa: E402  (checker only)
from tests.test_cabi_host import _matrices, _pack  # noqa: E402
from transhuman_b200 import _lib, synth  # noqa: E402


def f16(x):
    return x.to(torch.float16).to(torch.float64)


def f16_rz(x):
    """fp16 rounding toward zero (cvt.rz.f16.f32)"""
    h = x.to(torch.float16)
    over = h.to(torch.float64).abs() > x.abs()
    toward0 = torch.nextafter(h, torch.zeros_like(h))
    return torch.where(over, toward0, h).to(torch.float64)


def trunc_bits(x, bits):
    """round-to-nearest to `bits` explicit mantissa bits (TF32: 10, BF16: 7), float32 exponent range"""
    xi = x.to(torch.float32).contiguous().view(torch.int32).to(torch.int64)
    drop = 23 - bits
    xi = (xi + (1 << (drop - 1))) >> drop << drop
    return xi.to(torch.int32).view(torch.float32).to(torch.float64)


def make_lin(m, scheme):
    def lin(name, x):
        W, b = m[name]
        if scheme == "exact" or name in ("afc", "rgb"):      # the heads are fp32 dot products in the epilogue
            return x @ W.T + b
        if scheme in ("tf32", "bf16"):
            bits = 10 if scheme == "tf32" else 7
            return trunc_bits(x, bits) @ trunc_bits(W, bits).T + b
        # "3rz": the activation's high half rounded toward zero (what cvt.rz.relu.f16x2.f32 would give the
        # epilogue for free together with the ReLU); weights as shipped
        xh, wh = (f16_rz(x) if scheme == "3rz" else f16(x)), f16(W)
        xl, wl = f16(x - xh), f16(W - wh)
        out = xh @ wh.T
        if scheme in ("3", "3rz", "2a"):
            out = out + xl @ wh.T                            # activation low part
        if scheme in ("3", "3rz", "2w"):
            out = out + xh @ wl.T                            # weight low part
        return out + b
    return lin


def program(m, lin, rep_r, pix_r, vd_r, V):
    relu = torch.relu
    P = rep_r.shape[1]
    S = relu(lin("fc0", rep_r))
    X = relu(lin("ar0", pix_r))
    KP, KS = lin("k0", X), lin("k1", S)
    A = torch.softmax(torch.einsum("ipc,jpc->pij", KP, KS) / np.sqrt(128.0), dim=1)
    XT = torch.einsum("pij,ipc->jpc", A, X)
    inter = relu(lin("fc2", relu(lin("fc1f", torch.cat([S, XT], -1)))))
    alpha = lin("afc", relu(lin("fc3m", torch.cat(list(inter), -1))))
    G = relu(lin("gvf", torch.cat([inter, pix_r, vd_r.expand(V, P, 64)], -1)))
    T = relu(lin("t", torch.cat(list(G) + [pix_r.mean(0)], -1)))
    return torch.cat([lin("rgb", T), alpha], -1)


def main(H=20, W=20, S=16, seed=21, shift=-12.0):
    lib = _lib.load()
    fr = synth.make_frame(H=H, W=W, n_class=300, V=3, feat_hw=28, seed=seed, alpha_bias_shift=shift)
    tf = orc.to_torch_frame(fr)
    tokens = orc.build_tokens(tf)
    V = 3
    m = _matrices(_pack(lib, fr["weights"], V), V)
    ray_o, ray_d = tf["ray_o"][None], tf["ray_d"][None]
    pts, z_vals = orc.get_sampling_points(ray_o, ray_d, tf["near"][None], tf["far"][None], S)
    xyz = pts.clone().flatten(1, 2)
    pts_s = orc.world2smpl(pts, tf["Rh"][None], tf["Th"][None]).flatten(1, 2)
    vd = orc.view_embed(ray_d)[:, :, None].repeat(1, 1, S, 1).contiguous().view(1, -1, 27)
    pf = orc.get_pixel_aligned_feature(xyz, tf["input_R"], tf["input_T"], tf["input_K"], tf["pixel_feat_map"],
                                       tf["pixel_feat_map"].shape[-2:])
    rep = orc.human_representation(pts_s[0], tokens[0], tokens[1], tf["holder"], K=7)
    P = rep.shape[-1]
    rep_r = torch.cat([rep.double().permute(0, 2, 1), torch.zeros((V, P, 1), dtype=torch.float64)], -1)
    pix_r = pf.double().permute(0, 2, 1)
    vd_r = torch.cat([vd[0].double(), torch.zeros((P, 37), dtype=torch.float64)], -1)

    def composite(raw):
        rgb, acc, _, depth = orc.raw2outputs(raw.float().reshape(-1, S, 4), z_vals.view(-1, S), ray_d.view(-1, 3), False)
        return rgb, acc

    ref_raw = program(m, make_lin(m, "exact"), rep_r, pix_r, vd_r, V)
    ref_rgb, ref_acc = composite(ref_raw)
    # rays whose last sample sits on the sign step of alpha (excluded by the parity tests as well)
    edge = (ref_raw.reshape(-1, S, 4)[:, -1, 3].abs() < 1e-3)
    print(f"frame {H}x{W}x{S}, {P} samples, |raw| <= {ref_raw.abs().max():.1f}, {int(edge.sum())} knife-edge rays excluded")
    names = {"3": "fp16 hi/lo, 3 products (shipped)", "3rz": "3 products, activation hi rounded toward zero", "2a": "2 products: A_hi B_hi + A_lo B_hi (weights fp16)",
             "2w": "2 products: A_hi B_hi + A_hi B_lo (activations fp16)", "1": "1 product: fp16 x fp16",
             "tf32": "single pass, TF32 operands", "bf16": "single pass, BF16 operands"}
    for scheme, label in names.items():
        raw = program(m, make_lin(m, scheme), rep_r, pix_r, vd_r, V)
        rgb, acc = composite(raw)
        d_raw = ((raw - ref_raw).abs() / ref_raw.abs().clamp_min(1.0)).max().item()
        d_rgb = (rgb - ref_rgb)[~edge].abs().max().item()
        d_acc = (acc - ref_acc)[~edge].abs().max().item()
        print(f"  {label:52s} raw {d_raw:.2e} (rel. max(1,|raw|))   rgb_map {d_rgb:.2e}   acc_map {d_acc:.2e}")


if __name__ == "__main__":
    main()
    main(H=24, W=24, S=16, seed=11, shift=-15.0)

"""Algorithmic byte model of the rasterizer hot path, per view.

`roofline.achieved` in bench.py is these bytes divided by measured time.  The model is the
*reference algorithm's* unavoidable traffic -- in particular the sort is charged
`passes = ceil((32 + getHigherMsb(tiles)) / 8)` passes over 24 bytes/instance regardless of how the
sort is actually implemented here -- evaluated with the run's measured P, V, L, N, tiles.
"""
from __future__ import annotations

import math


def higher_msb(n):
    msb, step = 16, 16
    while step > 1:
        step //= 2
        msb = msb + step if (n >> msb) else msb - step
    if n >> msb:
        msb += 1
    return msb


def stage_bytes(P, V, L, W, H, *, sh_degree=3, sh_coeffs=16, use_sh=True, extended=True, training=True):
    """Returns {stage: bytes} using the term table of SURVEY.md Appendix E (F1..F9, B1..B4)."""
    N = W * H
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    passes = math.ceil((32 + higher_msb(tiles)) / 8)
    c_feat = 7 if extended else 3
    c_out = 8 if extended else 3
    sh_in = 12 * (sh_degree + 1) ** 2 * V if use_sh else 12 * V
    f = {}
    f["preprocess"] = (44 * P + (0) + sh_in + (12 * V if extended else 0)                      # F1 + F2
                       + 8 * P + 28 * V + (12 * V if use_sh else 0)                            # F3
                       + ((24 * V + (3 * V if use_sh else 0)) if training else 0))
    f["scan"] = 8 * P                                                                          # F4
    f["duplicate"] = 8 * P + 12 * V + 12 * L                                                   # F5
    f["sort"] = passes * 24 * L + 8 * L                                                        # F6
    f["ranges"] = 8 * L + 16 * tiles                                                           # F7
    f["blend_fwd"] = (28 + 4 * c_feat) * L + (4 * c_out + 8) * N                               # F8 + F9
    b = {}
    b["blend_bwd"] = (28 + 4 * c_feat) * L + (4 * c_out + 8) * N + (4 * c_feat + 24) * V       # B1 + B2
    b["preprocess_bwd"] = (4 * P + (64 + 4 * c_feat + 24) * V                                  # B3
                           + ((12 * (sh_degree + 1) ** 2 + 3) * V if use_sh else 0)
                           + 56 * P + (12 * sh_coeffs * P if use_sh else 12 * P)               # B4
                           + (12 * P if extended else 0))
    return dict(forward=f, backward=b, passes=passes, tiles=tiles,
                bytes_fwd=sum(f.values()), bytes_bwd=sum(b.values()))

"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT.

ctypes front-end of oracle/gs_oracle.c (the plain-C CPU restatement of the reference
rasterizer).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module; the product package never does.

`forward()` runs the reference's stage order (DGR/cuda_rasterizer/rasterizer_impl.cu:198-336):
preprocess -> inclusive scan -> duplicateWithKeys -> stable radix sort on
[0, 32+getHigherMsb(tiles)) -> identifyTileRanges -> render, and returns every
intermediate so tests can compare stage by stage.  `backward()` mirrors
rasterizer_impl.cu:340-434.  `extended=True` gives the SDP-GS outputs (depth, alpha,
feature; SURVEY.md Appendix D -- parity-unpinned by reference code) by blending the 8
channels [r,g,b,z,1,f0,f1,f2] with background [bg,0,0,0,0,0].
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "cpu"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle_cpu.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "gs_oracle.c")):
            build()
        _LIB = C.CDLL(path)
        _LIB.gso_scan.restype = C.c_int64
        _LIB.gso_higher_msb.restype = C.c_uint32
        _LIB.gso_num_threads.restype = C.c_int
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def num_threads():
    return int(lib().gso_num_threads())


def set_num_threads(n):
    lib().gso_set_num_threads(C.c_int(int(n)))


def higher_msb(n):
    return int(lib().gso_higher_msb(C.c_uint32(n)))


def mark_visible(means3D, viewmatrix):
    means3D, viewmatrix = _f32(means3D), _f32(viewmatrix)
    out = np.zeros(means3D.shape[0], dtype=np.uint8)
    lib().gso_mark_visible(C.c_int(means3D.shape[0]), _p(means3D), _p(viewmatrix), _p(out))
    return out.astype(bool)


def forward(means3D, opacities, cam, bg, *, shs=None, colors_precomp=None, scales=None, rotations=None,
            cov3D_precomp=None, sh_degree=3, scale_modifier=1.0, extended=False, features=None,
            confidence=None, stop_after=None):
    """Returns a dict with every stage's output.  `cam` has width,height,tanfovx,tanfovy,viewmatrix,projmatrix,campos."""
    L_ = lib()
    means3D = _f32(means3D)
    P = means3D.shape[0]
    W, H = int(cam.width), int(cam.height)
    opac = _f32(opacities).reshape(-1)
    if confidence is not None:
        opac = (opac * _f32(confidence).reshape(-1)).astype(np.float32)
    shs, colors_precomp, scales, rotations, cov3D_precomp = map(_f32, (shs, colors_precomp, scales, rotations, cov3D_precomp))
    M = 0 if shs is None else shs.shape[1]
    view, proj, campos = _f32(cam.viewmatrix), _f32(cam.projmatrix), _f32(cam.campos)
    bg = _f32(bg)

    o = dict(P=P, W=W, H=H, M=M, D=sh_degree)
    o["radii"] = np.zeros(P, np.int32)
    o["means2D"] = np.zeros((P, 2), np.float32)
    o["depths"] = np.zeros(P, np.float32)
    o["cov3D"] = np.zeros((P, 6), np.float32)
    o["rgb"] = np.zeros((P, 3), np.float32)
    o["conic_opacity"] = np.zeros((P, 4), np.float32)
    o["tiles_touched"] = np.zeros(P, np.uint32)
    o["clamped"] = np.zeros((P, 3), np.uint8)
    o["rect"] = np.zeros((P, 4), np.uint32)
    L_.gso_preprocess(C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D), _p(scales), C.c_float(scale_modifier),
                      _p(rotations), _p(opac), _p(shs), _p(cov3D_precomp), _p(colors_precomp), _p(view), _p(proj),
                      _p(campos), C.c_int(W), C.c_int(H), C.c_float(cam.tanfovx), C.c_float(cam.tanfovy),
                      _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]), _p(o["rgb"]),
                      _p(o["conic_opacity"]), _p(o["tiles_touched"]), _p(o["clamped"]), _p(o["rect"]))
    if stop_after == "preprocess":
        return o
    o["point_offsets"] = np.zeros(P, np.uint32)
    Lr = int(L_.gso_scan(C.c_int(P), _p(o["tiles_touched"]), _p(o["point_offsets"])))
    o["num_rendered"] = Lr
    keys_u = np.zeros(max(Lr, 1), np.uint64)
    vals_u = np.zeros(max(Lr, 1), np.uint32)
    L_.gso_duplicate_with_keys(C.c_int(P), _p(o["means2D"]), _p(o["depths"]), _p(o["point_offsets"]), _p(o["radii"]),
                               C.c_int(W), C.c_int(H), _p(keys_u), _p(vals_u))
    gx, gy = (W + 15) // 16, (H + 15) // 16
    bit = higher_msb(gx * gy)
    keys = np.zeros_like(keys_u)
    vals = np.zeros_like(vals_u)
    L_.gso_sort_pairs(C.c_int64(Lr), _p(keys_u), _p(vals_u), _p(keys), _p(vals), C.c_int(32 + bit))
    o["keys_unsorted"], o["values_unsorted"] = keys_u[:Lr], vals_u[:Lr]
    o["point_list_keys"], o["point_list"] = keys[:Lr], vals[:Lr]
    o["sort_bits"] = 32 + bit
    o["ranges"] = np.zeros((gx * gy, 2), np.uint32)
    L_.gso_tile_ranges(C.c_int64(Lr), _p(keys), C.c_int(gx * gy), _p(o["ranges"]))
    if stop_after == "binning":
        return o

    rgb = colors_precomp if colors_precomp is not None else o["rgb"]
    if extended:
        feat = _f32(features) if features is not None else rgb
        chans = np.concatenate([rgb, o["depths"][:, None], np.ones((P, 1), np.float32), feat], axis=1)
        bgc = np.concatenate([bg, np.zeros(5, np.float32)])
    else:
        chans, bgc = rgb, bg
    chans = np.ascontiguousarray(chans, np.float32)
    bgc = np.ascontiguousarray(bgc, np.float32)
    Cn = chans.shape[1]
    o["channels"], o["bg_channels"] = chans, bgc
    o["final_T"] = np.zeros((H, W), np.float32)
    o["n_contrib"] = np.zeros((H, W), np.uint32)
    out = np.zeros((Cn, H, W), np.float32)
    pl = np.ascontiguousarray(vals)
    L_.gso_render_forward(C.c_int(W), C.c_int(H), C.c_int(Cn), _p(o["ranges"]), _p(pl), _p(o["means2D"]), _p(chans),
                          _p(o["conic_opacity"]), _p(bgc), _p(o["final_T"]), _p(o["n_contrib"]), _p(out))
    o["out_all"] = out
    o["color"] = out[:3]
    if extended:
        o["depth"], o["alpha"], o["feature"] = out[3:4], out[4:5], out[5:8]
    o["_inputs"] = dict(means3D=means3D, shs=shs, colors_precomp=colors_precomp, scales=scales, rotations=rotations,
                        cov3D_precomp=cov3D_precomp, view=view, proj=proj, campos=campos, bg=bg, cam=cam,
                        scale_modifier=scale_modifier, extended=extended, confidence=confidence, features=features)
    return o


def backward(o, dL_dcolor, dL_ddepth=None, dL_dalpha=None, dL_dfeature=None):
    """Gradients for the forward result `o` (dict from forward()).  Mirrors rasterizer_impl.cu:340-434."""
    L_ = lib()
    i = o["_inputs"]
    P, W, H, M, D = o["P"], o["W"], o["H"], o["M"], o["D"]
    cam = i["cam"]
    chans, bgc = o["channels"], o["bg_channels"]
    Cn = chans.shape[1]
    if i["extended"]:
        z = lambda a, c: np.zeros((c, H, W), np.float32) if a is None else _f32(a).reshape(c, H, W)
        dpix = np.concatenate([_f32(dL_dcolor), z(dL_ddepth, 1), z(dL_dalpha, 1), z(dL_dfeature, 3)], axis=0)
    else:
        dpix = _f32(dL_dcolor)
    dpix = np.ascontiguousarray(dpix, np.float32)
    g = {}
    g["means2D"] = np.zeros((P, 3), np.float32)
    g["conic"] = np.zeros((P, 4), np.float32)
    dop = np.zeros(P, np.float32)
    dch = np.zeros((P, Cn), np.float32)
    pl = np.ascontiguousarray(o["point_list"])
    L_.gso_render_backward(C.c_int(P), C.c_int(W), C.c_int(H), C.c_int(Cn), _p(o["ranges"]), _p(pl), _p(o["means2D"]),
                           _p(chans), _p(o["conic_opacity"]), _p(bgc), _p(o["final_T"]), _p(o["n_contrib"]), _p(dpix),
                           _p(g["means2D"]), _p(g["conic"]), _p(dop), _p(dch))
    if i["confidence"] is not None:
        dop = dop * _f32(i["confidence"]).reshape(-1)
    g["opacities"] = dop.reshape(P, 1)
    dcolor = np.ascontiguousarray(dch[:, :3])
    dz = np.ascontiguousarray(dch[:, 3]) if i["extended"] else None
    if i["extended"]:
        if i["features"] is not None:
            g["features"] = np.ascontiguousarray(dch[:, 5:8])
        else:  # feature channels alias the colours (gaussian_renderer/__init__.py:298)
            dcolor = np.ascontiguousarray(dcolor + dch[:, 5:8])
    g["means3D"] = np.zeros((P, 3), np.float32)
    g["cov3D"] = np.zeros((P, 6), np.float32)
    g["shs"] = np.zeros((P, M, 3), np.float32)
    g["scales"] = np.zeros((P, 3), np.float32)
    g["rotations"] = np.zeros((P, 4), np.float32)
    focal_y = H / (2.0 * cam.tanfovy)
    focal_x = W / (2.0 * cam.tanfovx)
    cov3D = i["cov3D_precomp"] if i["cov3D_precomp"] is not None else o["cov3D"]
    L_.gso_preprocess_backward(
        C.c_int(P), C.c_int(D), C.c_int(M), _p(i["means3D"]), _p(o["radii"]), _p(i["shs"]), _p(o["clamped"]),
        _p(i["scales"]), _p(i["rotations"]), C.c_float(i["scale_modifier"]), _p(cov3D), _p(i["view"]), _p(i["proj"]),
        C.c_float(focal_x), C.c_float(focal_y), C.c_float(cam.tanfovx), C.c_float(cam.tanfovy), _p(i["campos"]),
        _p(g["means2D"]), _p(g["conic"]), _p(dcolor), _p(dz), _p(g["means3D"]), _p(g["cov3D"]), _p(g["shs"]),
        _p(g["scales"]), _p(g["rotations"]))
    g["colors_precomp"] = dcolor
    g["dz"] = dz
    return g

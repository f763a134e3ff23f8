"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT.

ctypes driver for oracle/_ref/libref_rasterizer.so: the UNMODIFIED reference CUDA rasterizer
(submodules/diff-gaussian-rasterization) compiled for sm_100a by oracle/Makefile behind the torch-free
shim oracle/ref_build/ref_capi.cu.  Needs a GPU.  Used by tests/ (parity), tests/golden/make_golden.py
(fixtures) and bench.py's `--impl reference` arm.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_rasterizer.so")
_LIB = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(LIB_PATH)
        for f in ("ref_required_geom", "ref_required_image", "ref_required_binning"):
            getattr(_LIB, f).restype = C.c_size_t
            getattr(_LIB, f).argtypes = [C.c_int]
        _LIB.ref_last_error.restype = C.c_char_p
    return _LIB


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _t(a, dev="cuda"):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)


class RefState:
    pass


def forward(means3D, opacities, cam, bg, *, shs=None, colors_precomp=None, scales=None, rotations=None,
            cov3D_precomp=None, sh_degree=3, scale_modifier=1.0, binning_cap_bytes=None, decode=True):
    """One call of CudaRasterizer::Rasterizer::forward.  Returns a RefState with color, radii, the three
    byte buffers and (decode=True) every decodable intermediate as numpy arrays."""
    L = lib()
    s = RefState()
    s.means3D, s.opac = _t(means3D), _t(opacities).reshape(-1)
    s.shs, s.colors, s.scales, s.rots, s.cov = _t(shs), _t(colors_precomp), _t(scales), _t(rotations), _t(cov3D_precomp)
    s.view, s.proj, s.campos, s.bg = _t(cam.viewmatrix), _t(cam.projmatrix), _t(cam.campos), _t(bg)
    s.cam, s.D, s.mod = cam, int(sh_degree), float(scale_modifier)
    P = s.P = s.means3D.shape[0]
    W, H = int(cam.width), int(cam.height)
    s.M = 0 if s.shs is None else int(s.shs.shape[1])
    s.color = torch.zeros((3, H, W), dtype=torch.float32, device="cuda")
    s.radii = torch.zeros((P,), dtype=torch.int32, device="cuda")
    gb, ib = L.ref_required_geom(P), L.ref_required_image(W * H)
    s.geom = torch.zeros((gb,), dtype=torch.uint8, device="cuda")
    s.img = torch.zeros((ib,), dtype=torch.uint8, device="cuda")
    needed = C.c_size_t(0)

    def call(binning, cap):
        torch.cuda.synchronize()
        return L.ref_forward(
            C.c_int(P), C.c_int(s.D), C.c_int(s.M), _p(s.bg), C.c_int(W), C.c_int(H), _p(s.means3D), _p(s.shs), _p(s.colors),
            _p(s.opac), _p(s.scales), C.c_float(s.mod), _p(s.rots), _p(s.cov), _p(s.view), _p(s.proj), _p(s.campos),
            C.c_float(cam.tanfovx), C.c_float(cam.tanfovy), C.c_int(0), _p(s.color), _p(s.radii),
            _p(s.geom), C.c_size_t(gb), _p(binning), C.c_size_t(cap), C.byref(needed), _p(s.img), C.c_size_t(ib), C.c_int(0))

    if binning_cap_bytes is None:
        n = call(None, 0)  # sizing call: fails with -2 after the reference's own D2H of num_rendered
        if n == -2:
            cap = int(needed.value)
            s.binning = torch.zeros((cap,), dtype=torch.uint8, device="cuda")
            n = call(s.binning, cap)
        else:  # num_rendered == 0 can succeed with an empty buffer
            s.binning = torch.zeros((max(int(needed.value), 1),), dtype=torch.uint8, device="cuda")
    else:
        s.binning = torch.zeros((binning_cap_bytes,), dtype=torch.uint8, device="cuda")
        n = call(s.binning, binning_cap_bytes)
    if n < 0:
        raise RuntimeError("reference forward failed: " + L.ref_last_error().decode())
    torch.cuda.synchronize()
    s.num_rendered = int(n)
    if decode:
        _decode(s)
    return s


def _decode(s):
    L = lib()
    P, R = s.P, s.num_rendered
    W, H = int(s.cam.width), int(s.cam.height)
    off = (C.c_int64 * 9)()
    L.ref_geom_layout(C.c_void_p(0), C.c_int(P), off)
    g = s.geom.cpu().numpy()
    view = lambda buf, o, dt, n: np.frombuffer(buf, dtype=dt, count=n, offset=int(o)).copy()
    s.depths = view(g, off[0], np.float32, P)
    s.clamped = view(g, off[1], np.uint8, 3 * P).reshape(P, 3)
    s.means2D = view(g, off[3], np.float32, 2 * P).reshape(P, 2)
    s.cov3D = view(g, off[4], np.float32, 6 * P).reshape(P, 6)
    s.conic_opacity = view(g, off[5], np.float32, 4 * P).reshape(P, 4)
    s.rgb = view(g, off[6], np.float32, 3 * P).reshape(P, 3)
    s.tiles_touched = view(g, off[7], np.uint32, P)
    s.point_offsets = view(g, off[8], np.uint32, P)
    off3 = (C.c_int64 * 3)()
    L.ref_image_layout(C.c_void_p(0), C.c_int(W * H), off3)
    im = s.img.cpu().numpy()
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    s.final_T = view(im, off3[0], np.float32, W * H).reshape(H, W)
    s.n_contrib = view(im, off3[1], np.uint32, W * H).reshape(H, W)
    s.ranges = view(im, off3[2], np.uint32, 2 * tiles).reshape(tiles, 2)
    if R > 0:
        off4 = (C.c_int64 * 4)()
        L.ref_binning_layout(C.c_void_p(0), C.c_int(R), off4)
        b = s.binning.cpu().numpy()
        s.point_list = view(b, off4[0], np.uint32, R)
        s.point_list_keys = view(b, off4[2], np.uint64, R)
    else:
        s.point_list = np.zeros(0, np.uint32)
        s.point_list_keys = np.zeros(0, np.uint64)


def backward(s, dL_dcolor):
    """One call of CudaRasterizer::Rasterizer::backward on the state of forward().  Returns numpy grads."""
    L = lib()
    P, M = s.P, s.M
    W, H = int(s.cam.width), int(s.cam.height)
    dpix = _t(dL_dcolor)
    z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device="cuda")
    g = dict(means2D=z(P, 3), conic=z(P, 4), opacities=z(P, 1), colors=z(P, 3), means3D=z(P, 3), cov3D=z(P, 6),
             shs=z(P, max(M, 1), 3), scales=z(P, 3), rotations=z(P, 4))
    rc = L.ref_backward(
        C.c_int(P), C.c_int(s.D), C.c_int(M), C.c_int(s.num_rendered), _p(s.bg), C.c_int(W), C.c_int(H), _p(s.means3D),
        _p(s.shs), _p(s.colors), _p(s.scales), C.c_float(s.mod), _p(s.rots), _p(s.cov), _p(s.view), _p(s.proj), _p(s.campos),
        C.c_float(s.cam.tanfovx), C.c_float(s.cam.tanfovy), _p(s.radii), _p(s.geom), _p(s.binning), _p(s.img), _p(dpix),
        _p(g["means2D"]), _p(g["conic"]), _p(g["opacities"]), _p(g["colors"]), _p(g["means3D"]), _p(g["cov3D"]),
        _p(g["shs"]), _p(g["scales"]), _p(g["rotations"]), C.c_int(0))
    if rc != 0:
        raise RuntimeError("reference backward failed: " + L.ref_last_error().decode())
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in g.items()}
    if M == 0:
        out["shs"] = np.zeros((P, 0, 3), np.float32)
    return out


def mark_visible(means3D, cam):
    L = lib()
    m, v, p = _t(means3D), _t(cam.viewmatrix), _t(cam.projmatrix)
    present = torch.zeros((m.shape[0],), dtype=torch.bool, device="cuda")
    L.ref_mark_visible(C.c_int(m.shape[0]), _p(m), _p(v), _p(p), _p(present))
    torch.cuda.synchronize()
    return present.cpu().numpy()


class RefRasterize(torch.autograd.Function):
    """The reference rasterizer as an autograd op, mirroring `_RasterizeGaussians`
    (DGR/diff_gaussian_rasterization/__init__.py:44-155) with the allocation pattern of its binding
    (DGR/rasterize_points.cu:60-77 outputs + byte buffers, :151-159 nine zero-filled gradient tensors).
    Used by bench.py's reference arm to time a training iteration of the reference's stock path.
    `cfg` = dict(cam, view, proj, campos, bg, D, binning_bytes)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cfg):
        L = lib()
        cam = cfg["cam"]
        P, W, H = means3D.shape[0], int(cam.width), int(cam.height)
        M = 0 if sh is None else int(sh.shape[1])
        color = torch.full((3, H, W), 0.0, dtype=torch.float32, device=means3D.device)
        radii = torch.full((P,), 0, dtype=torch.int32, device=means3D.device)
        gb, ib = L.ref_required_geom(P), L.ref_required_image(W * H)
        geom = torch.empty((gb,), dtype=torch.uint8, device=means3D.device)
        img = torch.empty((ib,), dtype=torch.uint8, device=means3D.device)
        binning = torch.empty((cfg["binning_bytes"],), dtype=torch.uint8, device=means3D.device)
        needed = C.c_size_t(0)
        means3D, opacities = means3D.contiguous(), opacities.contiguous()
        sh = None if sh is None else sh.contiguous()
        colors_precomp = None if colors_precomp is None else colors_precomp.contiguous()
        scales, rotations = scales.contiguous(), rotations.contiguous()
        n = L.ref_forward(C.c_int(P), C.c_int(cfg["D"]), C.c_int(M), _p(cfg["bg"]), C.c_int(W), C.c_int(H), _p(means3D), _p(sh),
                          _p(colors_precomp), _p(opacities), _p(scales), C.c_float(1.0), _p(rotations), None, _p(cfg["view"]),
                          _p(cfg["proj"]), _p(cfg["campos"]), C.c_float(cam.tanfovx), C.c_float(cam.tanfovy), C.c_int(0), _p(color),
                          _p(radii), _p(geom), C.c_size_t(gb), _p(binning), C.c_size_t(cfg["binning_bytes"]), C.byref(needed), _p(img),
                          C.c_size_t(ib), C.c_int(0))
        if n < 0:
            raise RuntimeError("reference forward failed: " + L.ref_last_error().decode())
        ctx.cfg, ctx.n, ctx.M = cfg, int(n), M
        ctx.has_sh = sh is not None
        ctx.save_for_backward(means3D, sh if sh is not None else torch.empty(0), colors_precomp if colors_precomp is not None else torch.empty(0),
                              scales, rotations, radii, geom, binning, img)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, g_color, _g_radii):
        L = lib()
        means3D, sh, colors, scales, rotations, radii, geom, binning, img = ctx.saved_tensors
        cfg, cam = ctx.cfg, ctx.cfg["cam"]
        P, W, H, M = means3D.shape[0], int(cam.width), int(cam.height), ctx.M
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=means3D.device)
        g = dict(means2D=z(P, 3), conic=z(P, 2, 2), opacities=z(P, 1), colors=z(P, 3), means3D=z(P, 3), cov3D=z(P, 6),
                 shs=z(P, max(M, 1), 3), scales=z(P, 3), rotations=z(P, 4))
        shp = sh if ctx.has_sh else None
        colp = colors if not ctx.has_sh else None
        rc = L.ref_backward(C.c_int(P), C.c_int(cfg["D"]), C.c_int(M), C.c_int(ctx.n), _p(cfg["bg"]), C.c_int(W), C.c_int(H), _p(means3D),
                            _p(shp), _p(colp), _p(scales), C.c_float(1.0), _p(rotations), None, _p(cfg["view"]), _p(cfg["proj"]),
                            _p(cfg["campos"]), C.c_float(cam.tanfovx), C.c_float(cam.tanfovy), _p(radii), _p(geom), _p(binning), _p(img),
                            _p(g_color.contiguous()), _p(g["means2D"]), _p(g["conic"]), _p(g["opacities"]), _p(g["colors"]),
                            _p(g["means3D"]), _p(g["cov3D"]), _p(g["shs"]), _p(g["scales"]), _p(g["rotations"]), C.c_int(0))
        if rc != 0:
            raise RuntimeError("reference backward failed")
        return (g["means3D"], g["means2D"], g["shs"] if ctx.has_sh else None, g["colors"] if not ctx.has_sh else None,
                g["opacities"], g["scales"], g["rotations"], None)

"""Shared test plumbing: named parity cases and uniform runners for the three implementations.

  run_oracle    -- oracle/gs_oracle.c through oracle/cpu_oracle.py            (CPU, the checker)
  run_product   -- sdp-gs_b200 through the C-ABI / drop-in Python surface      (GPU, the thing tested)
  run_reference -- oracle/_ref/libref_rasterizer.so, the unmodified reference (GPU, pins the oracle)

Every runner returns a dict with the same keys so comparisons read the same everywhere.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sdp-gs_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from b200gs import synthetic as syn  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> (config, view index, options)
CASES = {
    "tiny_sh3_ext": dict(config="tiny", view=0, extended=True, sh_degree=3, bg=(0.0, 0.0, 0.0)),
    "small_sh3": dict(config="small", view=1, extended=False, sh_degree=3, bg=(0.0, 0.0, 0.0)),
    "inside_sh2_white_ext": dict(config="inside", view=0, extended=True, sh_degree=2, bg=(1.0, 1.0, 1.0)),
    "small_precomp": dict(config="small", view=0, extended=False, sh_degree=0, bg=(0.2, 0.4, 0.6), colors_precomp=True,
                          cov3D_precomp=True),
    "tiny_sh0_mod": dict(config="tiny", view=1, extended=False, sh_degree=0, bg=(0.0, 0.0, 0.0), scale_modifier=0.7,
                         opacity="init"),
    "inside_sh1": dict(config="inside", view=1, extended=False, sh_degree=1, bg=(0.0, 0.0, 0.0)),
    # D4: per-Gaussian confidence != 1 multiplies the opacity; pinned against the reference kernels run
    # with opacities * confidence (and dL/dopacity scaled by the confidence)
    "tiny_sh3_ext_conf": dict(config="tiny", view=1, extended=True, sh_degree=3, bg=(0.1, 0.0, 0.3), confidence=True),
}
# cases the reference CUDA code can run directly (vanilla outputs); extended ones are cross-checked by channel packing
GRAD_KEYS = ("means3D", "means2D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D")


def case_inputs(name):
    c = CASES[name]
    sc = syn.make_config(c["config"], opacity=c.get("opacity", "trained"))
    cam = sc.cameras[c["view"]]
    inp = dict(name=name, cam=cam, means3D=sc.means3D, opacities=sc.opacities, bg=np.array(c["bg"], np.float32),
               sh_degree=c["sh_degree"], scale_modifier=c.get("scale_modifier", 1.0), extended=c["extended"],
               shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None, features=None,
               shs_language=None, confidence=None)
    rng = np.random.default_rng(12345)
    if c.get("colors_precomp"):
        inp["colors_precomp"] = rng.uniform(0, 1, size=(sc.P, 3)).astype(np.float32)
    else:
        inp["shs"] = sc.shs
    if c.get("cov3D_precomp"):
        q, s = sc.rotations.astype(np.float64), sc.scales.astype(np.float64)
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                      2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                      2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], axis=1).reshape(-1, 3, 3)
        Lm = R * s[:, None, :]
        cov = Lm @ np.transpose(Lm, (0, 2, 1))
        inp["cov3D_precomp"] = np.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2],
                                         cov[:, 2, 2]], axis=1).astype(np.float32)
    else:
        inp["scales"], inp["rotations"] = sc.scales, sc.rotations
    if c["extended"]:
        inp["features"] = sc.features
    if c.get("confidence"):
        inp["confidence"] = rng.uniform(0.3, 1.0, size=(sc.P, 1)).astype(np.float32)
    return inp


def case_cotangents(inp, seed=99):
    return syn.cotangents(inp["cam"], seed)


# ---------------------------------------------------------------- oracle (CPU)
def run_oracle(inp, backward=True, cot=None):
    from oracle import cpu_oracle as orc
    o = orc.forward(inp["means3D"], inp["opacities"], inp["cam"], inp["bg"], shs=inp["shs"],
                    colors_precomp=inp["colors_precomp"], scales=inp["scales"], rotations=inp["rotations"],
                    cov3D_precomp=inp["cov3D_precomp"], sh_degree=inp["sh_degree"], scale_modifier=inp["scale_modifier"],
                    extended=inp["extended"], features=inp["features"], confidence=inp["confidence"])
    res = dict(radii=o["radii"], depths=o["depths"], means2D=o["means2D"], conic_opacity=o["conic_opacity"],
               rgb=(inp["colors_precomp"] if inp["colors_precomp"] is not None else o["rgb"]), rect=o["rect"],
               tiles_touched=o["tiles_touched"], num_rendered=o["num_rendered"], point_list=o["point_list"],
               point_list_keys=o["point_list_keys"], ranges=o["ranges"], final_T=o["final_T"],
               n_contrib=o["n_contrib"], color=o["color"], clamped=o["clamped"], cov3D=o["cov3D"])
    if inp["extended"]:
        res.update(depth=o["depth"], alpha=o["alpha"], feature=o["feature"])
    if backward:
        cot = cot or case_cotangents(inp)
        g = orc.backward(o, *(cot if inp["extended"] else cot[:1]))
        res["grads"] = dict(means3D=g["means3D"], means2D=g["means2D"], opacities=g["opacities"], shs=g["shs"],
                            colors_precomp=g["colors_precomp"], scales=g["scales"], rotations=g["rotations"],
                            cov3D=g["cov3D"], features=g.get("features"), conic=g["conic"])
    return res


# ---------------------------------------------------------------- product (GPU)
def _settings(inp, dev):
    import torch
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    cam = inp["cam"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    kw = dict(image_height=cam.height, image_width=cam.width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=t(inp["bg"]),
              scale_modifier=inp["scale_modifier"], viewmatrix=t(cam.viewmatrix), projmatrix=t(cam.projmatrix),
              sh_degree=inp["sh_degree"], campos=t(cam.campos), prefiltered=False, debug=False)
    if inp["extended"]:
        P = inp["means3D"].shape[0]
        conf = t(inp["confidence"]) if inp["confidence"] is not None else torch.ones((P, 1), device=dev)
        kw.update(include_feature=True, confidence=conf)
    return GaussianRasterizationSettings(**kw)


def run_product(inp, backward=True, cot=None, dev="cuda"):
    """Through the public drop-in surface (GaussianRasterizer) for outputs/gradients, plus a decode of the
    saved workspaces (b200gs_*_layout) for the intermediates."""
    import torch
    from b200gs import _lib
    from b200gs import rasterizer as rz
    from diff_gaussian_rasterization import GaussianRasterizer

    t = lambda a, rg=True: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev).requires_grad_(rg)
    P = inp["means3D"].shape[0]
    T = dict(means3D=t(inp["means3D"]), opacities=t(inp["opacities"]), shs=t(inp["shs"]),
             colors_precomp=t(inp["colors_precomp"]), scales=t(inp["scales"]), rotations=t(inp["rotations"]),
             cov3D_precomp=t(inp["cov3D_precomp"]), features=t(inp["features"]), shs_language=t(inp["shs_language"]))
    means2D = torch.zeros((P, 3), device=dev, requires_grad=True)
    rs = _settings(inp, dev)
    rast = GaussianRasterizer(rs)
    kwargs = dict(means3D=T["means3D"], means2D=means2D, opacities=T["opacities"], shs=T["shs"],
                  colors_precomp=T["colors_precomp"], scales=T["scales"], rotations=T["rotations"],
                  cov3D_precomp=T["cov3D_precomp"])
    if inp["extended"]:
        kwargs.update(shs_language=T["shs_language"], language_feature_precomp=T["features"])
    outs = rast(**kwargs)
    res = {}
    if inp["extended"]:
        color, depth, alpha, feature, radii = outs
        res.update(depth=depth.detach().cpu().numpy(), alpha=alpha.detach().cpu().numpy(),
                   feature=feature.detach().cpu().numpy())
    else:
        color, radii = outs
    res["color"] = color.detach().cpu().numpy()
    res["radii"] = radii.cpu().numpy()

    # intermediates from the saved workspaces
    node = color.grad_fn
    saved = node.saved_tensors
    geom, binning, img = saved[-3].cpu().numpy(), saved[-2].cpu().numpy(), saved[-1].cpu().numpy()
    cam = inp["cam"]
    W, H = cam.width, cam.height
    off = (C.c_int64 * 7)()
    _lib.lib.b200gs_geom_layout(P, off)
    view = lambda buf, o, dt, n: np.frombuffer(buf, dtype=dt, count=n, offset=int(o)).copy()
    hdr = view(geom, off[0], np.uint64, 1)
    L = int(hdr[0])
    res["num_rendered"] = L
    vis = res["radii"] > 0
    depths = view(geom, off[1], np.float32, P)
    res["depths"] = np.where(vis, depths, 0).astype(np.float32)
    rect = view(geom, off[2], np.uint16, 4 * P).reshape(P, 4).astype(np.uint32)
    res["rect"] = rect
    res["tiles_touched"] = ((rect[:, 2] - rect[:, 0]) * (rect[:, 3] - rect[:, 1])).astype(np.uint32)
    rec = view(geom, off[3], np.float32, 16 * P).reshape(P, 16)
    rec = np.where(vis[:, None], rec, 0).astype(np.float32)
    res["means2D"] = rec[:, 0:2].copy()
    res["conic_opacity"] = np.stack([rec[:, 2], rec[:, 3], rec[:, 4], rec[:, 5]], axis=1)
    res["rgb"] = rec[:, 8:11].copy()
    res["rec"] = rec
    res["order"] = view(geom, off[5], np.uint32, P)
    off3 = (C.c_int64 * 3)()
    _lib.lib.b200gs_image_layout(W, H, off3)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    res["final_T"] = view(img, off3[0], np.float32, W * H).reshape(H, W)
    res["n_contrib"] = view(img, off3[1], np.uint32, W * H).reshape(H, W)
    res["ranges"] = view(img, off3[2], np.uint32, 2 * tiles).reshape(tiles, 2)
    off2 = (C.c_int64 * 2)()
    cap = int(getattr(node, "capacity", L))
    _lib.lib.b200gs_binning_layout(W, H, cap, off2)  # layout depends on the workspace capacity
    L = min(L, cap)  # instances beyond the capacity were dropped (overflow flag raised)
    res["point_list"] = view(binning, off2[0], np.uint32, L) if L else np.zeros(0, np.uint32)
    tile_ids = view(binning, off2[1], np.uint32, L) if L else np.zeros(0, np.uint32)
    dbits = depths.view(np.uint32)
    res["point_list_keys"] = (tile_ids.astype(np.uint64) << np.uint64(32)) | dbits[res["point_list"]].astype(np.uint64)

    if backward:
        cot = cot or case_cotangents(inp)
        tt = lambda a: torch.from_numpy(a).to(dev)
        if inp["extended"]:
            loss = (color * tt(cot[0])).sum() + (depth * tt(cot[1])).sum() + (alpha * tt(cot[2])).sum() + (feature * tt(cot[3])).sum()
        else:
            loss = (color * tt(cot[0])).sum()
        loss.backward()
        gnp = lambda x: None if (x is None or x.grad is None) else x.grad.detach().cpu().numpy()
        res["grads"] = dict(means3D=gnp(T["means3D"]), means2D=gnp(means2D), opacities=gnp(T["opacities"]),
                            shs=gnp(T["shs"]), colors_precomp=gnp(T["colors_precomp"]), scales=gnp(T["scales"]),
                            rotations=gnp(T["rotations"]), cov3D=gnp(T["cov3D_precomp"]), features=gnp(T["features"]),
                            shs_language=gnp(T["shs_language"]))
    return res


# ---------------------------------------------------------------- reference CUDA (GPU)
def run_reference(inp, backward=True, cot=None):
    """The unmodified reference kernels: vanilla outputs only (color, radii, intermediates, vanilla grads)."""
    from oracle import ref_cuda as ref
    conf = inp.get("confidence")
    if conf is not None:  # the reference has no confidence input: it multiplies the opacity (Appendix D), so feed the product
        inp = dict(inp, opacities=(inp["opacities"] * conf.reshape(inp["opacities"].shape)).astype(np.float32), confidence=None)
    s = ref.forward(inp["means3D"], inp["opacities"], inp["cam"], inp["bg"], shs=inp["shs"],
                    colors_precomp=inp["colors_precomp"], scales=inp["scales"], rotations=inp["rotations"],
                    cov3D_precomp=inp["cov3D_precomp"], sh_degree=inp["sh_degree"], scale_modifier=inp["scale_modifier"])
    radii = s.radii.cpu().numpy()
    vis = radii > 0
    m = lambda a: np.where(vis.reshape((-1,) + (1,) * (a.ndim - 1)), a, 0).astype(a.dtype)
    res = dict(radii=radii, depths=m(s.depths), means2D=m(s.means2D), conic_opacity=m(s.conic_opacity),
               rgb=(inp["colors_precomp"] if inp["colors_precomp"] is not None else m(s.rgb)),
               tiles_touched=s.tiles_touched, num_rendered=s.num_rendered, point_list=s.point_list,
               point_list_keys=s.point_list_keys, ranges=s.ranges, final_T=s.final_T, n_contrib=s.n_contrib,
               color=s.color.cpu().numpy(), clamped=m(s.clamped), cov3D=m(s.cov3D))
    if backward:
        cot = cot or case_cotangents(inp)
        g = ref.backward(s, cot[0])
        res["grads"] = dict(means3D=g["means3D"], means2D=g["means2D"], opacities=g["opacities"], shs=g["shs"],
                            colors_precomp=g["colors"], scales=g["scales"], rotations=g["rotations"], cov3D=g["cov3D"],
                            conic=g["conic"])
        if conf is not None:  # d(o * conf)/do
            res["grads"]["opacities"] = (res["grads"]["opacities"].reshape(-1) * conf.reshape(-1)).astype(np.float32).reshape(g["opacities"].shape)
    res["_state"] = s
    return res


# ---------------------------------------------------------------- comparisons
INT_KEYS = ("radii", "tiles_touched", "num_rendered", "point_list", "point_list_keys", "ranges", "n_contrib")


def rel_err(a, b):
    """Per-tensor relative error ||a-b|| / ||b|| (the 1e-3 gradient bar of BASELINE.md §5)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.linalg.norm(a - b)
    n = np.linalg.norm(b)
    return d / n if n > 0 else d


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)

"""PyTorch host side of the b200gs rasterizer: the autograd op behind GaussianRasterizer.

Mirrors `_RasterizeGaussians` of the reference (DGR/diff_gaussian_rasterization/__init__.py:44-155):
same argument order, same saved state (three opaque byte workspaces + num_rendered), same gradient
tuple order, the same debug-snapshot behaviour -- but the C++/pybind layer (DGR/rasterize_points.cu)
is replaced by ctypes calls into the C-ABI of include/b200gs.h, on torch's current CUDA stream.
PyTorch is only used for device memory, streams and autograd plumbing.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import Gaussians, GradOutputs, Grads, Outputs, View, Workspace, check, lib

# When not None: binning capacity (instances) used instead of the reference's blocking D2H read of
# num_rendered (DGR/cuda_rasterizer/rasterizer_impl.cu:281).  See set_binning_capacity().
_CAPACITY = None


_AUTO = {}      # (P, W, H, device) -> capacity learned from earlier forwards ("auto" mode)
_PENDING = []   # [(event, pinned header copy, key)] status read-backs of earlier no-sync forwards, checked lazily
_FREE_STATUS = []  # recycled (event, pinned 16-byte buffer) pairs


def set_binning_capacity(capacity):
    """How the binning workspace (one slot per (Gaussian, tile) instance) is sized.
    None (default): exactly, with one 8-byte D2H sync per forward, as the reference does (rasterizer_impl.cu:281).
    int: no host sync; instances beyond `capacity` are dropped and the overflow flag is raised on the device.
    "auto": the first forward of a given (P, W, H) runs in exact mode; later ones run without any host sync with
      1.3x the largest instance count seen so far.  Each forward's status word is copied back asynchronously and
      checked at the next call: an overflow (the scene changed so much that the margin was exceeded) raises a
      B200GSError one call late and the capacity is re-learned."""
    global _CAPACITY
    _CAPACITY = capacity if (capacity is None or capacity == "auto") else int(capacity)
    if capacity != "auto":
        _AUTO.clear()
        _PENDING.clear()


def _check_pending(block=False):
    """Inspect completed status read-backs of earlier "auto" forwards."""
    import numpy as np
    keep = []
    err = None
    for ev, host, key in _PENDING:
        if not block and not ev.query():
            keep.append((ev, host, key))
            continue
        if block:
            ev.synchronize()
        words = host.numpy().view(np.uint64)
        n = int(words[0])
        overflow = int(host.numpy().view(np.uint32)[2])
        if n * 1.15 > _AUTO.get(key, 0):
            _AUTO[key] = int(n * 1.3) + 4096
        if overflow & 1:
            err = _lib.B200GSError(f"binning capacity overflow in an earlier forward (num_rendered={n}); capacity re-learned")
        _FREE_STATUS.append((ev, host))
    _PENDING[:] = keep
    if err is not None:
        raise err


def _ptr(t):
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t, name):
    if t is None or t.numel() == 0:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (b200gs has no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _carve(flat, shapes, align=64):
    """Views of one flat allocation (one allocator call instead of one per output); `align` in elements."""
    out, off = [], 0
    for shp in shapes:
        n = 1
        for d in shp:
            n *= d
        out.append(flat[off:off + n].view(shp))
        off += (n + align - 1) // align * align
    return out


def _carve_size(shapes, align=64):
    tot = 0
    for shp in shapes:
        n = 1
        for d in shp:
            n *= d
        tot += (n + align - 1) // align * align
    return tot


def _stream(device=None):
    """The raw cudaStream_t torch is currently launching on (a C call; torch.cuda.current_stream() costs ~10x more)."""
    idx = torch.cuda.current_device() if device is None or device.index is None else device.index
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(idx))


def _cpu_deep_copy_tuple(args):
    return tuple(a.detach().cpu().clone() if isinstance(a, torch.Tensor) else a for a in args)


_WS_BYTES = {}
_BIN_BYTES = {}


def _ws_bytes(P, W, H):
    k = (P, W, H)
    v = _WS_BYTES.get(k)
    if v is None:
        v = _WS_BYTES[k] = (int(lib.b200gs_geom_bytes(P)), int(lib.b200gs_image_bytes(W, H)))
    return v


class _State:
    """What forward leaves behind for backward / inspection."""
    __slots__ = ("view", "gauss", "keep", "P", "W", "H", "capacity", "num_rendered", "extended")


def _build_view(rs, P, M, extended, keep):
    v = View()
    v.width, v.height = int(rs.image_width), int(rs.image_height)
    v.tan_fovx, v.tan_fovy = float(rs.tanfovx), float(rs.tanfovy)
    v.scale_modifier = float(rs.scale_modifier)
    v.sh_degree, v.sh_coeffs = int(rs.sh_degree), int(M)
    v.prefiltered, v.debug, v.extended = int(bool(rs.prefiltered)), int(bool(rs.debug)), int(extended)
    bg = _f32c(rs.bg, "bg")
    vm = _f32c(rs.viewmatrix, "viewmatrix")
    pm = _f32c(rs.projmatrix, "projmatrix")
    cp = _f32c(rs.campos, "campos")
    keep.extend([bg, vm, pm, cp])
    v.background, v.viewmatrix, v.projmatrix, v.campos = bg.data_ptr(), vm.data_ptr(), pm.data_ptr(), cp.data_ptr()
    return v


_VIEW_CACHE = {}  # (id(settings), M, extended) -> (settings, View, tensors kept alive)


def _cached_view(rs, P, M, extended, keep):
    """The View struct of a settings tuple is a pure function of it (pointers of its four tensors + scalars): built once per
    settings object instead of once per forward and once per backward (the tuple is immutable; its tensors may be updated in
    place, the pointers stay valid).  Not cached when a tensor had to be converted (the struct would point at a private copy)."""
    key = (id(rs), M, extended)
    hit = _VIEW_CACHE.get(key)
    if hit is not None and hit[0] is rs:
        return hit[1]
    mine = []
    v = _build_view(rs, P, M, extended, mine)
    if all(a is b for a, b in zip(mine, (rs.bg, rs.viewmatrix, rs.projmatrix, rs.campos))):
        if len(_VIEW_CACHE) > 512:
            _VIEW_CACHE.clear()
        _VIEW_CACHE[key] = (rs, v, mine)
    else:
        keep.extend(mine)
    return v


def _forward_impl(rs, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, shs_language,
                  language_feature_precomp, confidence, extended):
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:57-59
    dev = means3D.device
    P = means3D.shape[0]
    H, W = int(rs.image_height), int(rs.image_width)
    keep = []
    means3D = _f32c(means3D, "means3D")
    sh, colors_precomp = _f32c(sh, "shs"), _f32c(colors_precomp, "colors_precomp")
    opacities = _f32c(opacities, "opacities")
    scales, rotations = _f32c(scales, "scales"), _f32c(rotations, "rotations")
    cov3Ds_precomp = _f32c(cov3Ds_precomp, "cov3D_precomp")
    shs_language = _f32c(shs_language, "shs_language")
    language_feature_precomp = _f32c(language_feature_precomp, "language_feature_precomp")
    confidence = _f32c(confidence, "confidence")
    M = 0 if sh is None else int(sh.shape[1])  # rasterize_points.cu:83-87

    g = Gaussians()
    g.P = P
    g.means3D, g.shs, g.colors_precomp, g.opacities = _ptr(means3D), _ptr(sh), _ptr(colors_precomp), _ptr(opacities)
    g.scales, g.rotations, g.cov3D_precomp = _ptr(scales), _ptr(rotations), _ptr(cov3Ds_precomp)
    g.language_feature_precomp, g.shs_language, g.confidence = _ptr(language_feature_precomp), _ptr(shs_language), _ptr(confidence)
    v = _cached_view(rs, P, M, extended, keep)

    f32 = dict(dtype=torch.float32, device=dev)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)
    depth = alpha = feature = None
    o = Outputs()
    if extended:  # one allocation for the four maps: eight consecutive planes
        planes = torch.empty((8, H, W), **f32)
        color, depth, alpha, feature = planes[0:3], planes[3:4], planes[4:5], planes[5:8]
        o.depth, o.alpha, o.feature = depth.data_ptr(), alpha.data_ptr(), feature.data_ptr()
    else:
        color = torch.empty((3, H, W), **f32)
    o.color, o.radii = color.data_ptr(), _ptr(radii)

    gb, ib = _ws_bytes(P, W, H)
    wsbuf = torch.empty((gb + ib,), dtype=torch.uint8, device=dev)  # geom | image in one allocation (both 256-byte multiples)
    geom, img = wsbuf[:gb], wsbuf[gb:]
    ws = Workspace()
    ws.geom, ws.geom_bytes, ws.image, ws.image_bytes = geom.data_ptr(), gb, img.data_ptr(), ib
    stream = _stream(dev)
    if P == 0:  # rasterize_points.cu:81: nothing is launched for an empty scene, outputs are zeros
        color.zero_()
        if extended:
            depth.zero_(); alpha.zero_(); feature.zero_()
        binning = torch.empty((0,), dtype=torch.uint8, device=dev)
        return 0, 0, color, depth, alpha, feature, radii, geom, binning, img
    auto_key = None
    if _CAPACITY == "auto":
        _check_pending()
        auto_key = (P, W, H, str(dev))
    if _CAPACITY is None or (auto_key is not None and auto_key not in _AUTO):
        n = C.c_int64(0)
        check(lib.b200gs_forward_preprocess(C.byref(v), C.byref(g), C.byref(o), C.byref(ws), stream, C.byref(n)))
        num_rendered, capacity = int(n.value), int(n.value)
        if auto_key is not None:
            _AUTO[auto_key] = int(n.value * 1.3) + 4096
            auto_key = None  # exact this time, nothing to verify
        binning = torch.empty((lib.b200gs_binning_bytes(capacity, W, H),), dtype=torch.uint8, device=dev)
        ws.binning, ws.binning_bytes = binning.data_ptr(), binning.numel()
        check(lib.b200gs_forward_render(C.byref(v), C.byref(g), C.byref(o), C.byref(ws), C.c_int64(capacity), stream))
    else:  # capacity known up front: one call, no host synchronization
        num_rendered, capacity = -1, int(_AUTO[auto_key] if auto_key is not None else _CAPACITY)
        bb = _BIN_BYTES.get((capacity, W, H))
        if bb is None:
            bb = _BIN_BYTES[(capacity, W, H)] = int(lib.b200gs_binning_bytes(capacity, W, H))
        binning = torch.empty((bb,), dtype=torch.uint8, device=dev)
        ws.binning, ws.binning_bytes = binning.data_ptr(), bb
        check(lib.b200gs_forward(C.byref(v), C.byref(g), C.byref(o), C.byref(ws), C.c_int64(capacity), stream))
    if auto_key is not None:  # asynchronous read-back of {num_rendered, overflow}; looked at by a later call
        ev, host = _FREE_STATUS.pop() if _FREE_STATUS else (torch.cuda.Event(), torch.empty((16,), dtype=torch.uint8, pin_memory=True))
        host.copy_(geom[:16], non_blocking=True)
        ev.record()
        _PENDING.append((ev, host, auto_key))
    del keep
    return num_rendered, capacity, color, depth, alpha, feature, radii, geom, binning, img


def _backward_impl(rs, num_rendered, capacity, extended, means3D, sh, colors_precomp, opacities, scales, rotations,
                   cov3Ds_precomp, shs_language, language_feature_precomp, confidence, radii, geom, binning, img,
                   g_color, g_depth, g_alpha, g_feature):
    dev = means3D.device
    P = means3D.shape[0]
    keep = []
    M = 0 if (sh is None or sh.numel() == 0) else int(sh.shape[1])
    f32 = dict(dtype=torch.float32, device=dev)
    has = lambda t: t is not None and t.numel() != 0
    # every buffer is fully written by the kernels (zeros where radii == 0): no torch.zeros fills.  One allocation, carved.
    want = [("means3D", (P, 3), True), ("means2D", (P, 3), True), ("opac", (P, 1), True), ("sh", (P, M, 3), has(sh)),
            ("colors", (P, 3), has(colors_precomp)), ("scales", (P, 3), has(scales)), ("rots", (P, 4), has(rotations)),
            ("cov", (P, 6), has(cov3Ds_precomp)), ("feat", (P, 3), extended and has(language_feature_precomp)),
            ("shl", tuple(shs_language.shape) if has(shs_language) else (P, 3), extended and has(shs_language)),
            ("scratch", (P, 16), P > 0)]
    shapes = [shp for _, shp, on in want if on]
    views = iter(_carve(torch.empty((max(_carve_size(shapes), 1),), **f32), shapes))
    got = {name: (next(views) if on else None) for name, _, on in want}
    d_means3D, d_means2D, d_opac, d_sh, d_colors = got["means3D"], got["means2D"], got["opac"], got["sh"], got["colors"]
    d_scales, d_rots, d_cov, d_feat, d_shl, scratch = got["scales"], got["rots"], got["cov"], got["feat"], got["shl"], got["scratch"]
    if P == 0:
        return d_means3D, d_means2D, d_sh, d_colors, d_opac, d_scales, d_rots, d_cov, d_shl, d_feat

    g = Gaussians()
    g.P = P
    g.means3D, g.shs, g.colors_precomp, g.opacities = _ptr(means3D), _ptr(sh), _ptr(colors_precomp), _ptr(opacities)
    g.scales, g.rotations, g.cov3D_precomp = _ptr(scales), _ptr(rotations), _ptr(cov3Ds_precomp)
    g.language_feature_precomp, g.shs_language, g.confidence = _ptr(language_feature_precomp), _ptr(shs_language), _ptr(confidence)
    v = _cached_view(rs, P, M, extended, keep)
    ws = Workspace()
    ws.geom, ws.geom_bytes, ws.image, ws.image_bytes = geom.data_ptr(), geom.numel(), img.data_ptr(), img.numel()
    ws.binning, ws.binning_bytes = binning.data_ptr(), binning.numel()
    go = GradOutputs()
    gc = _f32c(g_color, "grad color")
    gd, ga, gf = _f32c(g_depth, "grad depth"), _f32c(g_alpha, "grad alpha"), _f32c(g_feature, "grad feature")
    go.dL_dcolor, go.dL_ddepth, go.dL_dalpha, go.dL_dfeature = _ptr(gc), _ptr(gd), _ptr(ga), _ptr(gf)
    gr = Grads()
    gr.dL_dmeans3D, gr.dL_dmeans2D, gr.dL_dshs, gr.dL_dcolors = _ptr(d_means3D), _ptr(d_means2D), _ptr(d_sh), _ptr(d_colors)
    gr.dL_dopacities, gr.dL_dscales, gr.dL_drotations, gr.dL_dcov3D = _ptr(d_opac), _ptr(d_scales), _ptr(d_rots), _ptr(d_cov)
    gr.dL_dfeatures, gr.dL_dshs_language, gr.scratch = _ptr(d_feat), _ptr(d_shl), scratch.data_ptr()
    check(lib.b200gs_backward(C.byref(v), C.byref(g), radii.data_ptr(), C.byref(ws), C.c_int64(capacity), C.byref(go),
                              C.byref(gr), _stream(dev)))
    del keep
    return d_means3D, d_means2D, d_sh, d_colors, d_opac, d_scales, d_rots, d_cov, d_shl, d_feat


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                shs_language, language_feature_precomp, raster_settings, extended):
        rs = raster_settings
        confidence = getattr(rs, "confidence", None)
        args = (rs, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, shs_language,
                language_feature_precomp, confidence, extended)
        if rs.debug:
            cpu_args = _cpu_deep_copy_tuple(args[1:])  # copy them before they can be corrupted
            try:
                res = _forward_impl(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            res = _forward_impl(*args)
        num_rendered, capacity, color, depth, alpha, feature, radii, geom, binning, img = res
        ctx.raster_settings = rs
        ctx.num_rendered, ctx.capacity, ctx.extended = num_rendered, capacity, extended
        ctx.confidence = confidence
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, opacities,
                              shs_language, language_feature_precomp, geom, binning, img)
        ctx.mark_non_differentiable(radii)
        if extended:
            return color, depth, alpha, feature, radii
        return color, radii

    @staticmethod
    def backward(ctx, *grad_outs):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, opacities, shs_language,
         language_feature_precomp, geom, binning, img) = ctx.saved_tensors
        if ctx.extended:
            g_color, g_depth, g_alpha, g_feature = grad_outs[0], grad_outs[1], grad_outs[2], grad_outs[3]
        else:
            g_color, g_depth, g_alpha, g_feature = grad_outs[0], None, None, None
        args = (rs, ctx.num_rendered, ctx.capacity, ctx.extended, means3D, sh, colors_precomp, opacities, scales,
                rotations, cov3Ds_precomp, shs_language, language_feature_precomp, ctx.confidence, radii, geom,
                binning, img, g_color, g_depth, g_alpha, g_feature)
        if rs.debug:
            cpu_args = _cpu_deep_copy_tuple(args[1:])
            try:
                grads = _backward_impl(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            grads = _backward_impl(*args)
        d_means3D, d_means2D, d_sh, d_colors, d_opac, d_scales, d_rots, d_cov, d_shl, d_feat = grads
        # order of forward()'s inputs (reference order + the two SDP-GS tensors + settings + mode flag)
        return (d_means3D, d_means2D, d_sh, d_colors, d_opac, d_scales, d_rots, d_cov, d_shl, d_feat, None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, shs_language=None, language_feature_precomp=None, extended=False):
    empty = lambda t: torch.Tensor([]) if t is None else t
    return _RasterizeGaussians.apply(means3D, means2D, empty(sh), empty(colors_precomp), opacities, empty(scales),
                                     empty(rotations), empty(cov3Ds_precomp), empty(shs_language),
                                     empty(language_feature_precomp), raster_settings, bool(extended))


def mark_visible(positions, viewmatrix, projmatrix):
    """bool[P] frustum mask (DGR/rasterize_points.cu:198-217)."""
    positions = _f32c(positions, "positions")
    P = 0 if positions is None else positions.shape[0]
    present = torch.zeros((P,), dtype=torch.bool, device=viewmatrix.device)
    if P:
        vm, pm = _f32c(viewmatrix, "viewmatrix"), _f32c(projmatrix, "projmatrix")
        check(lib.b200gs_mark_visible(P, positions.data_ptr(), vm.data_ptr(), pm.data_ptr(), present.data_ptr(), _stream()))
    return present


def launch_count():
    return int(lib.b200gs_launch_count())


_WARMED = set()  # (device index) whose kernels have been launched once outside a capture (lazy module loading)


def capture_graph(fn, device, warmup=None):
    """Capture `fn` into a CUDA graph on a side stream with the raw capture_begin / capture_end calls.  `torch.cuda.graph`
    would also run gc.collect() and torch.cuda.empty_cache() on entry -- milliseconds each, which is most of the cost of
    re-capturing after a densification.  Nothing inside `fn` may allocate through torch (the b200gs calls never do).
    `warmup` (default: fn) runs once per device before the first capture: lazy module loading must not happen while capturing."""
    key = torch.device(device).index
    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    if key not in _WARMED:
        with torch.cuda.stream(side):
            (warmup or fn)()
        torch.cuda.synchronize(device)
        _WARMED.add(key)
    import gc
    g = torch.cuda.CUDAGraph()
    gc_was_on = gc.isenabled()
    gc.disable()  # a collection inside the capture could destroy an old CUDA graph, which invalidates a global-mode capture
    try:
        with torch.cuda.stream(side):
            g.capture_begin(capture_error_mode="relaxed")
            try:
                fn()
            finally:
                g.capture_end()
    finally:
        if gc_was_on:
            gc.enable()
    torch.cuda.current_stream(device).wait_stream(side)
    return g


class RasterSession:
    """Pre-planned forward(+backward) over fixed device buffers: no allocation, no host sync, so a whole
    step can be captured once into a CUDA graph and replayed (`capture()` / `replay()`).

    This is the steady-state path for a training loop whose P and image size do not change between
    densification events: the binning workspace is sized for `capacity` instances; if a view ever needs
    more, the overflow flag is raised on the device (`status()` reads it back, which synchronizes)."""

    def __init__(self, raster_settings, *, means3D, opacities, shs=None, colors_precomp=None, scales=None,
                 rotations=None, cov3D_precomp=None, shs_language=None, language_feature_precomp=None,
                 extended=False, capacity, grads_out=None, with_backward=True, grad_scatter=None, live_count=None):
        """live_count: optional uint32/int32 device tensor of one element -- the number of live rows of capacity-sized
        parameter buffers (b200gs_gaussians_t.live_count); rows past it are culled, so the Gaussian count can change on the
        device without touching this session or its captured graph."""
        rs = raster_settings
        self.rs, self.extended, self.capacity = rs, bool(extended), int(capacity)
        dev = means3D.device
        self.dev = dev
        P = self.P = means3D.shape[0]
        H, W = int(rs.image_height), int(rs.image_width)
        self._keep = []
        f = lambda t, n: _f32c(t, n)
        self.inputs = dict(means3D=f(means3D, "means3D"), opacities=f(opacities, "opacities"), shs=f(shs, "shs"),
                           colors_precomp=f(colors_precomp, "colors_precomp"), scales=f(scales, "scales"),
                           rotations=f(rotations, "rotations"), cov3D_precomp=f(cov3D_precomp, "cov3D_precomp"),
                           shs_language=f(shs_language, "shs_language"),
                           language_feature_precomp=f(language_feature_precomp, "language_feature_precomp"),
                           confidence=f(getattr(rs, "confidence", None), "confidence"))
        i = self.inputs
        M = 0 if i["shs"] is None else int(i["shs"].shape[1])
        g = self.g = Gaussians()
        g.P = P
        g.means3D, g.shs, g.colors_precomp, g.opacities = _ptr(i["means3D"]), _ptr(i["shs"]), _ptr(i["colors_precomp"]), _ptr(i["opacities"])
        g.scales, g.rotations, g.cov3D_precomp = _ptr(i["scales"]), _ptr(i["rotations"]), _ptr(i["cov3D_precomp"])
        g.language_feature_precomp, g.shs_language, g.confidence = _ptr(i["language_feature_precomp"]), _ptr(i["shs_language"]), _ptr(i["confidence"])
        if live_count is not None:
            if live_count.device != dev or live_count.numel() != 1 or live_count.dtype not in (torch.int32, torch.uint32):
                raise RuntimeError("live_count must be a one-element int32/uint32 tensor on the session's device")
            self._keep.append(live_count)
            g.live_count = live_count.data_ptr()
        self.v = _build_view(rs, P, M, self.extended, self._keep)
        f32 = dict(dtype=torch.float32, device=dev)
        self.color = torch.empty((3, H, W), **f32)
        self.radii = torch.empty((P,), dtype=torch.int32, device=dev)
        o = self.o = Outputs()
        o.color, o.radii = self.color.data_ptr(), _ptr(self.radii)
        self.depth = self.alpha = self.feature = None
        if self.extended:
            self.depth, self.alpha, self.feature = torch.empty((1, H, W), **f32), torch.empty((1, H, W), **f32), torch.empty((3, H, W), **f32)
            o.depth, o.alpha, o.feature = self.depth.data_ptr(), self.alpha.data_ptr(), self.feature.data_ptr()
        u8 = dict(dtype=torch.uint8, device=dev)
        self.geom = torch.empty((lib.b200gs_geom_bytes(P),), **u8)
        self.img = torch.empty((lib.b200gs_image_bytes(W, H),), **u8)
        self.binning = torch.empty((lib.b200gs_binning_bytes(self.capacity, W, H),), **u8)
        ws = self.ws = Workspace()
        ws.geom, ws.geom_bytes, ws.image, ws.image_bytes = self.geom.data_ptr(), self.geom.numel(), self.img.data_ptr(), self.img.numel()
        ws.binning, ws.binning_bytes = self.binning.data_ptr(), self.binning.numel()
        self.graph = None
        self.with_backward = with_backward
        if not with_backward:
            ws.persistent = 1
            check(lib.b200gs_workspace_init(C.byref(ws), P, None, _stream(dev)))
        if with_backward:
            go = grads_out or {}
            has = lambda t: t is not None
            mk = lambda name, shape, cond=True: (go.get(name) if go.get(name) is not None else torch.empty(shape, **f32)) if cond else None
            self.grads = dict(
                means3D=mk("means3D", (P, 3)), means2D=mk("means2D", (P, 3)), opacities=mk("opacities", (P, 1)),
                shs=mk("shs", (P, M, 3), has(i["shs"])), colors_precomp=mk("colors_precomp", (P, 3), has(i["colors_precomp"])),
                scales=mk("scales", (P, 3), has(i["scales"])), rotations=mk("rotations", (P, 4), has(i["rotations"])),
                cov3D=mk("cov3D", (P, 6), has(i["cov3D_precomp"])),
                features=mk("features", (P, 3), self.extended and has(i["language_feature_precomp"])),
                shs_language=mk("shs_language", (P, 3), self.extended and has(i["shs_language"])))
            for k, t in self.grads.items():
                if t is not None and (not t.is_contiguous() or t.dtype != torch.float32):
                    raise RuntimeError(f"grads_out[{k}] must be a contiguous float32 tensor")
            self.scratch = torch.empty((lib.b200gs_scratch_bytes(P),), **u8)
            gr = self.gr = Grads()
            G = self.grads
            gr.dL_dmeans3D, gr.dL_dmeans2D, gr.dL_dshs, gr.dL_dcolors = _ptr(G["means3D"]), _ptr(G["means2D"]), _ptr(G["shs"]), _ptr(G["colors_precomp"])
            gr.dL_dopacities, gr.dL_dscales, gr.dL_drotations, gr.dL_dcov3D = _ptr(G["opacities"]), _ptr(G["scales"]), _ptr(G["rotations"]), _ptr(G["cov3D"])
            gr.dL_dfeatures, gr.dL_dshs_language, gr.scratch = _ptr(G["features"]), _ptr(G["shs_language"]), self.scratch.data_ptr()
            # persistent workspaces (b200gs_workspace_t.persistent): initialised once, left clean by the kernels, no memset per step
            ws.persistent = 1
            check(lib.b200gs_workspace_init(C.byref(ws), P, self.scratch.data_ptr(), _stream(dev)))
            if grad_scatter is not None:  # image-parallel training: parameter gradients are pushed to the owning ranks (parallel.FusedGradBuffer)
                gr.scatter_bases, gr.scatter_shard_rows, gr.scatter_rank, gr.scatter_world = grad_scatter
            self.cot = dict(color=torch.zeros((3, H, W), **f32))
            go_ = self.go = GradOutputs()
            go_.dL_dcolor = self.cot["color"].data_ptr()
            if self.extended:
                self.cot.update(depth=torch.zeros((1, H, W), **f32), alpha=torch.zeros((1, H, W), **f32),
                                feature=torch.zeros((3, H, W), **f32))
                go_.dL_ddepth, go_.dL_dalpha, go_.dL_dfeature = self.cot["depth"].data_ptr(), self.cot["alpha"].data_ptr(), self.cot["feature"].data_ptr()

    def forward(self):
        check(lib.b200gs_forward(C.byref(self.v), C.byref(self.g), C.byref(self.o), C.byref(self.ws),
                                 C.c_int64(self.capacity), _stream()))

    def backward(self):
        check(lib.b200gs_backward(C.byref(self.v), C.byref(self.g), self.radii.data_ptr(), C.byref(self.ws),
                                  C.c_int64(self.capacity), C.byref(self.go), C.byref(self.gr), _stream()))

    def step(self):
        """forward, then backward with the cotangents currently stored in `self.cot` (fill them between the two
        yourself by calling forward()/backward() separately when the loss depends on the outputs)."""
        self.forward()
        if self.with_backward:
            self.backward()

    def capture(self, fn=None):
        """Capture `fn` (default: self.step) into a CUDA graph on a side stream; returns self."""
        fn = fn or self.step
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            fn()  # warm-up outside capture (lazy module loading must not happen while capturing; `fn` may launch foreign kernels)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            fn()
        return self

    def replay(self):
        self.graph.replay()

    def status(self):
        """(num_rendered, overflow) -- synchronizes the current stream."""
        n, ov = C.c_int64(0), C.c_int32(0)
        rc = lib.b200gs_forward_status(C.byref(self.ws), _stream(), C.byref(n), C.byref(ov))
        if rc not in (0, -3):
            check(rc)
        return int(n.value), int(ov.value)

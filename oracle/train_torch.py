"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT.

Plain-PyTorch restatement of the torch-side operations of one SDP-GS training iteration, used (a) by tests/ as
the checker of the fused kernels of include/b200gs_train.h and (b) by `bench.py --impl reference` as the
reference's own (stock torch) code path around its CUDA rasterizer.  Each function cites what it follows:

  l1_loss, ssim            utils/loss_utils.py:119-163 (11x11 window, sigma 1.5, F.conv2d groups=C, zero padding)
  pearson_corrcoef         torchmetrics.functional.pearson_corrcoef as called at train.py:126-129
                           (torchmetrics is not in this image: restated from its published definition
                            r = cov(x, y) / sqrt(var(x) var(y)), clamped to [-1, 1])
  depth_loss               train.py:115-131
  activations              scene/gaussian_model.py:44-57 (exp, sigmoid, torch.nn.functional.normalize)
  make_optimizer           scene/gaussian_model.py:228-267 (torch.optim.Adam(lr=0, eps=1e-15), 7 groups)
  expon lr                 utils/general_utils.py:get_expon_lr_func
"""
from math import exp

import torch
import torch.nn.functional as F


def l1_loss(network_output, gt):
    return torch.abs((network_output - gt)).mean()


def _gaussian(window_size, sigma):
    gauss = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return gauss / gauss.sum()


def _create_window(window_size, channel):
    w1 = _gaussian(window_size, 1.5).unsqueeze(1)
    w2 = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous()


def ssim(img1, img2, window_size=11):
    channel = img1.size(-3)
    window = _create_window(window_size, channel).to(img1.device).type_as(img1)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def pearson_corrcoef(x, y):
    x, y = x.reshape(-1), y.reshape(-1)
    xm, ym = x - x.mean(), y - y.mean()
    r = (xm * ym).sum() / torch.sqrt((xm * xm).sum() * (ym * ym).sum())
    return torch.clamp(r, -1.0, 1.0)


def depth_loss(depth, depth_mono):
    d1, m1 = depth.reshape(-1, 1), depth_mono.reshape(-1, 1)
    return torch.min(1 - pearson_corrcoef(m1, d1), 1 - pearson_corrcoef(1 / (-m1 + 200), d1))


def total_loss(image, gt, depth, depth_mono, lambda_dssim, depth_weight):
    Ll1 = l1_loss(image, gt)
    S = ssim(image, gt)
    dl = depth_loss(depth, depth_mono)
    return (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - S) + depth_weight * dl, Ll1, S, depth_weight * dl


def activate(raw):
    """raw: dict of leaf tensors xyz, f_dc [P,1,3], f_rest [P,15,3], opacity, scaling, rotation, feature."""
    return dict(xyz=raw["xyz"], shs=torch.cat((raw["f_dc"], raw["f_rest"]), dim=1), opacity=torch.sigmoid(raw["opacity"]),
                scaling=torch.exp(raw["scaling"]), rotation=F.normalize(raw["rotation"]), feature=raw["feature"])


def make_optimizer(raw, hp):
    groups = [
        {"params": [raw["feature"]], "lr": hp["language_feature_lr"], "name": "language_feature"},
        {"params": [raw["f_dc"]], "lr": hp["feature_lr"], "name": "f_dc"},
        {"params": [raw["f_rest"]], "lr": hp["feature_lr"] / 20.0, "name": "f_rest"},
        {"params": [raw["xyz"]], "lr": hp["position_lr_init"] * hp["spatial_lr_scale"], "name": "xyz"},
        {"params": [raw["opacity"]], "lr": hp["opacity_lr"], "name": "opacity"},
        {"params": [raw["scaling"]], "lr": hp["scaling_lr"], "name": "scaling"},
        {"params": [raw["rotation"]], "lr": hp["rotation_lr"], "name": "rotation"},
    ]
    return torch.optim.Adam(groups, lr=0.0, eps=1e-15)


def set_xyz_lr(optimizer, lr):
    for g in optimizer.param_groups:
        if g["name"] == "xyz":
            g["lr"] = lr


# ---- utils/sh_utils.py:24-112 (constants and eval_sh, degrees 0-3) and the python-SH branch of the render glue
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def eval_sh(deg, sh, dirs):
    """sh [..., C, (deg+1)^2], dirs [..., 3] unit vectors -> [..., C]"""
    result = SH_C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        result = result - SH_C1 * y * sh[..., 1] + SH_C1 * z * sh[..., 2] - SH_C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            result = (result + SH_C2[0] * xy * sh[..., 4] + SH_C2[1] * yz * sh[..., 5] + SH_C2[2] * (2.0 * zz - xx - yy) * sh[..., 6]
                      + SH_C2[3] * xz * sh[..., 7] + SH_C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                result = (result + SH_C3[0] * y * (3 * xx - yy) * sh[..., 9] + SH_C3[1] * xy * z * sh[..., 10]
                          + SH_C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
                          + SH_C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + SH_C3[5] * z * (xx - yy) * sh[..., 14]
                          + SH_C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return result


def python_sh_colors(features, xyz, campos, active_sh_degree, max_sh_degree=3):
    """gaussian_renderer/__init__.py:269-274 (pipe.convert_SHs_python)"""
    shs_view = features.transpose(1, 2).view(-1, 3, (max_sh_degree + 1) ** 2)
    dir_pp = xyz - campos.repeat(features.shape[0], 1)
    dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
    return torch.clamp_min(eval_sh(active_sh_degree, shs_view, dir_pp_normalized) + 0.5, 0.0)


def python_language_feature(language_feature):
    """gaussian_renderer/__init__.py:281-287"""
    sh2language = eval_sh(0, language_feature.view(-1, 3, 1), None)
    return sh2language / (sh2language.norm(dim=-1, keepdim=True) + 1e-9)


# ---- scene/gaussian_model.py:400-608: densification on nn.Parameters + the torch optimizer state (restated; proximity()
# is left out because simple_knn is not vendored -- SURVEY.md F5)
class DensifyModel:
    """The slice of GaussianModel that densify_and_prune touches.  `raw`: dict of tensors xyz, f_dc [P,1,3], f_rest
    [P,15,3], opacity, scaling, rotation, feature; `state`: dict name -> (exp_avg, exp_avg_sq) or None."""
    NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation", "feature")

    def __init__(self, raw, hp, moments=None, percent_dense=0.01, prune_from_iter=500):
        from torch import nn
        self.p = {k: nn.Parameter(raw[k].clone().requires_grad_(True)) for k in self.NAMES}
        self.optimizer = make_optimizer(self.p, hp)
        self.group_of = {"language_feature": "feature"}
        if moments is not None:
            for g in self.optimizer.param_groups:
                k = self.group_of.get(g["name"], g["name"])
                self.optimizer.state[g["params"][0]] = {"step": torch.tensor(0.0), "exp_avg": moments[k][0].clone(),
                                                        "exp_avg_sq": moments[k][1].clone()}
        self.percent_dense, self.prune_from_iter = percent_dense, prune_from_iter
        P = raw["xyz"].shape[0]
        dev = raw["xyz"].device
        self.xyz_gradient_accum = torch.zeros((P, 1), device=dev)
        self.denom = torch.zeros((P, 1), device=dev)
        self.max_radii2D = torch.zeros((P,), device=dev)

    get_scaling = property(lambda self: torch.exp(self.p["scaling"]))
    get_opacity = property(lambda self: torch.sigmoid(self.p["opacity"]))

    def _key(self, group):
        return self.group_of.get(group["name"], group["name"])

    def _prune_optimizer(self, mask):
        from torch import nn
        for group in self.optimizer.param_groups:
            stored = self.optimizer.state.get(group["params"][0], None)
            if stored is not None:
                stored["exp_avg"] = stored["exp_avg"][mask]
                stored["exp_avg_sq"] = stored["exp_avg_sq"][mask]
                del self.optimizer.state[group["params"][0]]
                group["params"][0] = nn.Parameter(group["params"][0][mask].requires_grad_(True))
                self.optimizer.state[group["params"][0]] = stored
            else:
                group["params"][0] = nn.Parameter(group["params"][0][mask].requires_grad_(True))
            self.p[self._key(group)] = group["params"][0]

    def prune_points(self, mask, it):
        if it > self.prune_from_iter:
            valid = ~mask
            self._prune_optimizer(valid)
            self.xyz_gradient_accum = self.xyz_gradient_accum[valid]
            self.denom = self.denom[valid]
            self.max_radii2D = self.max_radii2D[valid]

    def cat_tensors_to_optimizer(self, d):
        from torch import nn
        for group in self.optimizer.param_groups:
            ext = d[self._key(group)]
            stored = self.optimizer.state.get(group["params"][0], None)
            if stored is not None:
                stored["exp_avg"] = torch.cat((stored["exp_avg"], torch.zeros_like(ext)), dim=0)
                stored["exp_avg_sq"] = torch.cat((stored["exp_avg_sq"], torch.zeros_like(ext)), dim=0)
                del self.optimizer.state[group["params"][0]]
                group["params"][0] = nn.Parameter(torch.cat((group["params"][0], ext), dim=0).requires_grad_(True))
                self.optimizer.state[group["params"][0]] = stored
            else:
                group["params"][0] = nn.Parameter(torch.cat((group["params"][0], ext), dim=0).requires_grad_(True))
            self.p[self._key(group)] = group["params"][0]

    def densification_postfix(self, d):
        self.cat_tensors_to_optimizer(d)
        P, dev = self.p["xyz"].shape[0], self.p["xyz"].device
        self.xyz_gradient_accum = torch.zeros((P, 1), device=dev)
        self.denom = torch.zeros((P, 1), device=dev)
        self.max_radii2D = torch.zeros((P,), device=dev)

    @staticmethod
    def build_rotation(r):
        norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
        q = r / norm[:, None]
        R = torch.zeros((q.size(0), 3, 3), device=r.device)
        r_, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r_ * z); R[:, 0, 2] = 2 * (x * z + r_ * y)
        R[:, 1, 0] = 2 * (x * y + r_ * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r_ * x)
        R[:, 2, 0] = 2 * (x * z - r_ * y); R[:, 2, 1] = 2 * (y * z + r_ * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
        return R

    def densify_and_split(self, grads, grad_threshold, scene_extent, it, N=2, generator=None):
        n_init = self.p["xyz"].shape[0]
        padded = torch.zeros((n_init,), device=grads.device)
        padded[:grads.shape[0]] = grads.squeeze()
        sel = torch.where(padded >= grad_threshold, True, False)
        sel = torch.logical_and(sel, torch.max(self.get_scaling, dim=1).values > self.percent_dense * scene_extent)
        stds = self.get_scaling[sel].repeat(N, 1)
        means = torch.zeros((stds.size(0), 3), device=grads.device)
        samples = torch.normal(mean=means, std=stds, generator=generator)
        rots = self.build_rotation(self.p["rotation"][sel]).repeat(N, 1, 1)
        d = dict(xyz=torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + self.p["xyz"][sel].repeat(N, 1),
                 scaling=torch.log(self.get_scaling[sel].repeat(N, 1) / (0.8 * N)), rotation=self.p["rotation"][sel].repeat(N, 1),
                 f_dc=self.p["f_dc"][sel].repeat(N, 1, 1), f_rest=self.p["f_rest"][sel].repeat(N, 1, 1),
                 opacity=self.p["opacity"][sel].repeat(N, 1), feature=self.p["feature"][sel].repeat(N, 1))
        self.densification_postfix(d)
        self.prune_points(torch.cat((sel, torch.zeros(N * int(sel.sum()), device=grads.device, dtype=bool))), it)

    def densify_and_clone(self, grads, grad_threshold, scene_extent):
        sel = torch.where(torch.norm(grads, dim=-1) >= grad_threshold, True, False)
        sel = torch.logical_and(sel, torch.max(self.get_scaling, dim=1).values <= self.percent_dense * scene_extent)
        self.densification_postfix({k: self.p[k][sel] for k in self.NAMES})

    @staticmethod
    def dist_knn3(xyz):
        """distCUDA2 of the SDP-GS simple_knn fork restated with torch: mean squared distance to the 3 nearest neighbours (self
        excluded) and their indices.  (simple_knn.cu of graphdeco-inria/simple-knn: boxMeanDist, K = 3; the fork adds the indices.)"""
        x = xyz.detach().double()
        P = x.shape[0]
        vals, idxs = [], []
        for a in range(0, P, 1024):  # row blocks: the full P x P matrix does not fit for P ~ 1e5
            d = ((x[a:a + 1024, None, :] - x[None, :, :]) ** 2).sum(-1)
            r = torch.arange(a, min(a + 1024, P), device=x.device)
            d[r - a, r] = float("inf")
            val, idx = torch.topk(d, 3, dim=1, largest=False, sorted=True)
            vals.append(val); idxs.append(idx)
        val, idx = torch.cat(vals), torch.cat(idxs)
        return val.mean(dim=1).float(), idx

    def reset_opacity(self):
        """scene/gaussian_model.py:351-355 + replace_tensor_to_optimizer (:400-413)"""
        from torch import nn
        new = torch.log(torch.min(self.get_opacity, torch.ones_like(self.get_opacity) * 0.01) / (1 - torch.min(self.get_opacity, torch.ones_like(self.get_opacity) * 0.01)))
        for group in self.optimizer.param_groups:
            if group["name"] == "opacity":
                stored = self.optimizer.state.get(group["params"][0], None)
                if stored is not None:
                    stored["exp_avg"] = torch.zeros_like(new)
                    stored["exp_avg_sq"] = torch.zeros_like(new)
                    del self.optimizer.state[group["params"][0]]
                group["params"][0] = nn.Parameter(new.detach().requires_grad_(True))
                if stored is not None:
                    self.optimizer.state[group["params"][0]] = stored
                self.p["opacity"] = group["params"][0]

    def proximity(self, scene_extent, N=3):
        """scene/gaussian_model.py:513-533"""
        dist, nearest = self.dist_knn3(self.p["xyz"])
        sel = torch.logical_and(dist > (5. * scene_extent), torch.max(self.get_scaling, dim=1).values > scene_extent)
        new_indices = nearest[sel].reshape(-1).long()
        source_xyz = self.p["xyz"][sel].repeat(1, N, 1).reshape(-1, 3)
        target_xyz = self.p["xyz"][new_indices]
        rot = torch.zeros_like(self.p["rotation"][new_indices])
        rot[:, 0] = 1
        self.densification_postfix(dict(xyz=(source_xyz + target_xyz) / 2, scaling=self.p["scaling"][new_indices], rotation=rot,
                                        f_dc=torch.zeros_like(self.p["f_dc"][new_indices]), f_rest=torch.zeros_like(self.p["f_rest"][new_indices]),
                                        opacity=self.p["opacity"][new_indices], feature=self.p["feature"][new_indices]))

    def densify_and_prune(self, max_grad, min_opacity, extent, max_screen_size, it, generator=None):
        grads = self.xyz_gradient_accum / self.denom
        grads[grads.isnan()] = 0.0
        self.densify_and_clone(grads, max_grad, extent)
        self.densify_and_split(grads, max_grad, extent, it, generator=generator)
        if it < 2000:
            self.proximity(extent)
        prune_mask = (self.get_opacity < min_opacity).squeeze()
        if max_screen_size:
            big_vs = self.max_radii2D > max_screen_size
            big_ws = self.get_scaling.max(dim=1).values > 0.1 * extent
            prune_mask = torch.logical_or(torch.logical_or(prune_mask, big_vs), big_ws)
        self.prune_points(prune_mask, it)

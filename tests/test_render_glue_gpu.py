"""`gaussian_renderer.render()` drop-in: the kernel-side SH evaluation / feature
normalisation against the reference's default python-SH branch (gaussian_renderer/__init__.py:268-292, restated in
oracle/train_torch.py) pushed through the same rasterizer as precomputed colours / features.
Tolerance: images 2e-5 max-abs, gradients 1e-3 relative (the two paths differ only in fp32 rounding)."""
import math
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import helpers


def _scene(dev, degree):
    from b200gs import synthetic as syn
    sc = syn.make_config("small")
    cam = sc.cameras[0]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    leaves = dict(xyz=t(sc.means3D), features=t(sc.shs), opacity=t(sc.opacities), scaling=t(sc.scales), rotation=t(sc.rotations),
                  language=t(sc.features))
    leaves = {k: v.requires_grad_(True) for k, v in leaves.items()}
    pc = SimpleNamespace(get_xyz=leaves["xyz"], get_features=leaves["features"], get_opacity=leaves["opacity"],
                         get_scaling=leaves["scaling"], get_rotation=leaves["rotation"], get_language_feature=leaves["language"],
                         active_sh_degree=degree, max_sh_degree=3, confidence=torch.ones((sc.P, 1), device=dev))
    view = SimpleNamespace(FoVx=2 * math.atan(cam.tanfovx), FoVy=2 * math.atan(cam.tanfovy), image_height=cam.height,
                           image_width=cam.width, world_view_transform=t(cam.viewmatrix), full_proj_transform=t(cam.projmatrix),
                           camera_center=t(cam.campos))
    return sc, cam, leaves, pc, view


def test_render_glue_imports_without_a_gpu():
    import inspect
    from gaussian_renderer import render
    assert list(inspect.signature(render).parameters) == ["viewpoint_camera", "pc", "pipe", "bg_color", "opt", "scaling_modifier",
                                                          "override_color", "override_language"]


@pytest.mark.gpu
@pytest.mark.parametrize("degree", [0, 3])
def test_render_matches_python_sh_branch(degree):
    from gaussian_renderer import render
    from oracle import train_torch as tt
    dev = torch.device("cuda", 0)
    sc, cam, leaves, pc, view = _scene(dev, degree)
    pipe = SimpleNamespace(convert_SHs_python=True, compute_cov3D_python=False, debug=False, use_confidence=False)
    opt = SimpleNamespace(include_feature=True)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    g = torch.Generator(device="cpu").manual_seed(3)
    cots = dict(render=torch.randn((3, cam.height, cam.width), generator=g).to(dev), depth=torch.randn((1, cam.height, cam.width), generator=g).to(dev),
                feature=torch.randn((3, cam.height, cam.width), generator=g).to(dev))

    pkg = render(view, pc, pipe, bg, opt)
    assert set(pkg) == {"render", "depth", "alpha", "opacity", "feature", "viewspace_points", "visibility_filter", "radii", "color"}
    loss = sum((pkg[k] * c).sum() for k, c in cots.items())
    loss.backward()
    mine = {k: v.grad.clone() for k, v in leaves.items()}
    mine_vs = pkg["viewspace_points"].grad.clone()
    for v in leaves.values():
        v.grad = None

    # the reference's branch: SH and feature normalisation in torch, handed over as precomputed values
    colors = tt.python_sh_colors(leaves["features"], leaves["xyz"], view.camera_center, degree)
    feats = tt.python_language_feature(leaves["language"])
    ref = render(view, pc, pipe, bg, opt, override_color=colors, override_language=feats)
    loss = sum((ref[k] * c).sum() for k, c in cots.items())
    loss.backward()
    for k in ("render", "depth", "alpha", "feature"):
        assert float((pkg[k] - ref[k]).abs().max()) <= 2e-5, k
    assert torch.equal(pkg["radii"], ref["radii"]) and torch.equal(pkg["visibility_filter"], ref["visibility_filter"])
    for k, v in leaves.items():
        assert helpers.rel_err(mine[k].cpu().numpy(), v.grad.cpu().numpy()) <= 1e-3, k
    assert helpers.rel_err(mine_vs.cpu().numpy(), ref["viewspace_points"].grad.cpu().numpy()) <= 1e-3

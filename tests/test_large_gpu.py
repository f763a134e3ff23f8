"""GPU: BASELINE.json's large shapes (configs[3], configs[4]) through size-independent properties, and against
the CPU oracle where it finishes in seconds (integer stages of the 3M-Gaussian forward)."""
import numpy as np
import pytest

import helpers
from helpers import bits, rel_err, run_oracle, run_product

pytestmark = pytest.mark.gpu


def _inputs(config, view, extended, P=None):
    from b200gs import synthetic as syn
    sc = syn.make_config(config, P=P, views=max(view + 1, 2))
    return dict(name=config, cam=sc.cameras[view], means3D=sc.means3D, opacities=sc.opacities, bg=np.zeros(3, np.float32),
                sh_degree=3, scale_modifier=1.0, extended=extended, shs=sc.shs, colors_precomp=None, scales=sc.scales,
                rotations=sc.rotations, cov3D_precomp=None, features=sc.features if extended else None, shs_language=None,
                confidence=None)


def _check_sorted(p):
    keys = p["point_list_keys"]
    assert (keys[1:] >= keys[:-1]).all()
    same = keys[1:] == keys[:-1]
    assert (p["point_list"][1:][same] > p["point_list"][:-1][same]).all()
    r = p["ranges"].astype(np.int64)
    nz = r[:, 1] > r[:, 0]
    assert (r[nz, 1] - r[nz, 0]).sum() == p["num_rendered"] == int(p["tiles_touched"].astype(np.int64).sum())
    # every tile's range holds exactly the instances keyed with that tile
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    counts = np.bincount(tiles, minlength=r.shape[0])
    np.testing.assert_array_equal(counts, r[:, 1] - r[:, 0])


def test_mip360_render_shape_forward():
    """configs[3]: 3M Gaussians at 1297x840, forward only (13-bit tile ids -> two tile-sort passes, large-tile sort path)."""
    import torch
    assert torch.cuda.is_available()
    inp = _inputs("mip360_render", 0, extended=False)
    p = run_product(inp, backward=False)
    _check_sorted(p)
    o = run_oracle(inp, backward=False)
    assert p["num_rendered"] == o["num_rendered"]
    for k in ("radii", "tiles_touched", "point_list", "point_list_keys", "ranges"):
        np.testing.assert_array_equal(p[k], o[k], err_msg=k)
    np.testing.assert_array_equal(bits(p["depths"]), bits(o["depths"]))
    assert (p["n_contrib"] != o["n_contrib"]).mean() <= 1e-3
    assert np.abs(p["color"] - o["color"]).max() <= 1e-4


def test_stress_train_shape_forward_backward():
    """configs[4]: 6M Gaussians at 1920x1080 with the SDP-GS outputs, forward + backward."""
    import torch
    assert torch.cuda.is_available()
    inp = _inputs("stress_train", 1, extended=True)
    cot = helpers.case_cotangents(inp, seed=3)
    p = run_product(inp, True, cot)
    _check_sorted(p)
    assert np.abs(p["alpha"][0] - (1.0 - p["final_T"])).max() <= 1e-5
    for k in ("means3D", "opacities", "scales", "rotations", "shs", "features"):
        g = np.asarray(p["grads"][k])
        assert np.isfinite(g).all(), k
        assert not g[p["radii"] == 0].any(), k
    # linearity of the backward in the cotangents
    p2 = run_product(inp, True, tuple(-0.5 * c for c in cot))
    for k in ("means3D", "opacities", "shs"):
        assert rel_err(p2["grads"][k], -0.5 * np.asarray(p["grads"][k])) <= 1e-4, k
    np.testing.assert_array_equal(bits(p2["color"]), bits(p["color"]))


def test_dtu_scan_shape_forward_backward():
    """configs[2]: ~300k Gaussians at 400x300 (25x19 = 475 tiles, 9-bit tile ids) with the rendered-depth output that feeds
    depthfusion.py: every integer stage bit-exact against the CPU oracle, maps within 1e-4, gradients within 1e-3."""
    import torch
    assert torch.cuda.is_available()
    inp = _inputs("dtu_scan_3view", 2, extended=True)
    cot = helpers.case_cotangents(inp, seed=5)
    p = run_product(inp, True, cot)
    _check_sorted(p)
    o = run_oracle(inp, True, cot)
    assert p["num_rendered"] == o["num_rendered"]
    for k in ("radii", "tiles_touched", "point_list", "point_list_keys", "ranges"):
        np.testing.assert_array_equal(p[k], o[k], err_msg=k)
    np.testing.assert_array_equal(bits(p["depths"]), bits(o["depths"]))
    assert (p["n_contrib"] != o["n_contrib"]).mean() <= 1e-3
    for k in ("color", "depth", "alpha", "feature"):
        assert np.abs(p[k] - o[k]).max() <= 1e-4, k
    for k in ("means3D", "opacities", "scales", "rotations", "shs", "features"):
        assert rel_err(p["grads"][k], o["grads"][k]) <= 1e-3, k

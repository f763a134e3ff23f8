"""Round-2 session features against the eager drop-in path (which the parity tests pin to the oracle / reference):
  * persistent workspaces (b200gs_workspace_t.persistent): no memset per step -- every replayed step must still produce the
    eager path's image bits, integer state and gradients, also after a forward-only call and across views;
  * b200gs_grads_t.accumulate: several views per optimizer step summed inside the backward kernel;
  * the pseudo-view iteration (train.py:138-153): second render + single-correlation Pearson loss, gradients of both
    renders accumulated, against torch autograd over the same rasterizer.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import helpers
from helpers import bits, case_cotangents, case_inputs, rel_err, run_product

gpu = pytest.mark.gpu


def _session(inp, dev, capacity, **kw):
    from b200gs import rasterizer as rz
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    rs = helpers._settings(inp, dev)
    return rz.RasterSession(rs, means3D=t(inp["means3D"]), opacities=t(inp["opacities"]), shs=t(inp["shs"]), scales=t(inp["scales"]),
                            rotations=t(inp["rotations"]), language_feature_precomp=t(inp["features"]), extended=inp["extended"],
                            capacity=capacity, **kw)


@gpu
@pytest.mark.parametrize("name", ["tiny_sh3_ext", "small_sh3"])
def test_persistent_session_matches_eager_path_step_after_step(name):
    dev = torch.device("cuda", 0)
    inp = case_inputs(name)
    cot = case_cotangents(inp)
    ref = run_product(inp, True, cot)  # eager: fresh workspaces + memsets every call
    s = _session(inp, dev, capacity=int(ref["num_rendered"] * 1.5) + 64)
    s.cot["color"].copy_(torch.from_numpy(cot[0]).to(dev))
    if inp["extended"]:
        s.cot["depth"].copy_(torch.from_numpy(cot[1]).to(dev)); s.cot["alpha"].copy_(torch.from_numpy(cot[2]).to(dev))
        s.cot["feature"].copy_(torch.from_numpy(cot[3]).to(dev))
    s.capture()
    for it in range(4):
        if it == 2:
            s.forward()  # a forward without its backward in between must not disturb the self-cleaning state
        s.replay()
        torch.cuda.synchronize()
        n, ov = s.status()
        assert (n, ov) == (ref["num_rendered"], 0)
        np.testing.assert_array_equal(bits(s.color.cpu().numpy()), bits(ref["color"]))
        np.testing.assert_array_equal(s.radii.cpu().numpy(), ref["radii"])
        for k in ("means3D", "opacities", "shs", "scales", "rotations"):
            assert rel_err(s.grads[k].cpu().numpy().reshape(ref["grads"][k].shape), ref["grads"][k]) <= 1e-5, (it, k)


@gpu
def test_accumulate_sums_views_inside_the_backward_kernel():
    """view A written, view B accumulated on top == gradient(A) + gradient(B) from two independent sessions."""
    dev = torch.device("cuda", 0)
    inps = [case_inputs("tiny_sh3_ext"), case_inputs("tiny_sh3_ext_conf")]  # same scene, two cameras (the confidence is dropped below)
    for i in inps:
        i["confidence"] = None
    cots = [case_cotangents(i, seed=7 + k) for k, i in enumerate(inps)]
    P = inps[0]["means3D"].shape[0]
    f32 = dict(dtype=torch.float32, device=dev)
    shared = dict(means3D=torch.empty((P, 3), **f32), shs=torch.empty((P, 16, 3), **f32), opacities=torch.empty((P, 1), **f32),
                  scales=torch.empty((P, 3), **f32), rotations=torch.empty((P, 4), **f32), features=torch.empty((P, 3), **f32))
    solo, acc = [], []
    for inp, cot in zip(inps, cots):
        for store, kw in ((solo, {}), (acc, dict(grads_out=shared))):
            s = _session(inp, dev, capacity=100_000, **kw)
            for k, c in zip(("color", "depth", "alpha", "feature"), cot):
                s.cot[k].copy_(torch.from_numpy(c).to(dev))
            store.append(s)
    for s in solo:
        s.step()
    acc[0].gr.accumulate = 0
    acc[0].step()
    acc[1].gr.accumulate = 1
    acc[1].step()
    torch.cuda.synchronize()
    for k in shared:
        want = solo[0].grads[k] + solo[1].grads[k]
        assert rel_err(shared[k].cpu().numpy(), want.cpu().numpy()) <= 1e-6, k
    # means2D (the densification input) is per view: never accumulated
    # (float REDs in a different order on every run: elementwise last bits differ, so the check is on the tensor's scale)
    assert rel_err(acc[1].grads["means2D"].cpu().numpy(), solo[1].grads["means2D"].cpu().numpy()) <= 1e-6


@gpu
def test_pseudo_view_iteration_matches_autograd():
    """GaussianTrainer.step_pair: training view (L1+SSIM + min-form Pearson) + pseudo view (single-correlation Pearson with its
    own weight), one Adam step -- against torch autograd over the same rasterizer, torch losses and torch.optim.Adam."""
    from b200gs.trainer import GaussianTrainer, DEFAULTS
    from diff_gaussian_rasterization import GaussianRasterizer
    from oracle import train_torch as tt
    from test_train_gpu import _trainer_inputs
    dev = torch.device("cuda", 0)
    sc, cams, gts, monos, raw = _trainer_inputs("small", dev)
    P = sc.P
    rng = np.random.default_rng(3)
    pseudo_ref = rng.uniform(1, 8, size=(1, cams[0].height, cams[0].width)).astype(np.float32)
    tr = GaussianTrainer(cameras=cams[:1], gt_images=gts[:1], depth_mono=monos[:1], device=dev, capacity=400_000, **raw)
    tr.add_pseudo_views(cams[1:2], [pseudo_ref])
    w = 0.35
    tr.set_pseudo_weight(w)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    leaf = dict(xyz=t(raw["xyz"]), f_dc=t(raw["shs"][:, :1]), f_rest=t(raw["shs"][:, 1:]), opacity=t(raw["opacity_raw"]).reshape(P, 1),
                scaling=t(raw["scaling_raw"]), rotation=t(raw["rotation_raw"]), feature=t(raw["feature"]))
    leaf = {k: v.requires_grad_(True) for k, v in leaf.items()}
    hp = dict(DEFAULTS)
    opt = tt.make_optimizer(leaf, hp)
    for it in range(1, 4):
        tr.step_pair(0, 0, use_graph=(it > 1))
        mine = tr.loss_values()
        a = tt.activate(leaf)

        def render(cam):
            m2 = torch.zeros((P, 3), device=dev, requires_grad=True)
            return GaussianRasterizer(tr._default_settings(cam))(means3D=a["xyz"], means2D=m2, opacities=a["opacity"], shs=a["shs"],
                                                                scales=a["scaling"], rotations=a["rotation"], shs_language=a["feature"])
        color, depth, _, _, _ = render(cams[0])
        total = tt.total_loss(color, t(gts[0]), depth, t(monos[0]), hp["lambda_dssim"], hp["depth_weight"])[0]
        dp = render(cams[1])[1]
        total = total + w * (1.0 - tt.pearson_corrcoef(dp.reshape(-1), t(pseudo_ref).reshape(-1)))
        total.backward()
        opt.step(); opt.zero_grad(set_to_none=True)
        assert abs(mine[0] - float(total)) <= 2e-4 * abs(float(total)) + 1e-7, (it, mine, float(total))
    torch.cuda.synchronize()
    ref = dict(xyz=leaf["xyz"], shs=torch.cat((leaf["f_dc"], leaf["f_rest"]), 1).reshape(P, 48), opacity=leaf["opacity"],
               scaling=leaf["scaling"], rotation=leaf["rotation"], feature=leaf["feature"])
    lrs = dict(xyz=hp["position_lr_init"], shs=hp["feature_lr"], opacity=hp["opacity_lr"], scaling=hp["scaling_lr"],
               rotation=hp["rotation_lr"], feature=hp["language_feature_lr"])
    for k in tr.raw:  # Adam moves an element by ~lr per step whatever the gradient's size: gradients that are rounding noise may
        err = (tr.raw[k] - ref[k].detach()).abs()  # flip sign between the pipelines, so the bar is on the bulk (as in test_train_gpu)
        frac = float((err <= 0.05 * lrs[k] * 3 + 1e-7).float().mean())
        assert frac >= 0.99, (k, frac, float(err.max()))

"""CPU check of the densification row logic (sdp-gs_b200/b200gs/densify_logic.py: pure torch, no native library) against the
restated reference procedure (oracle/train_torch.py::DensifyModel, scene/gaussian_model.py:400-608) on the same inputs and
the same RNG seed: identical rows, row order and Adam moments, with and without proximity(), with and without pruning."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sdp-gs_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

WIDTHS = dict(xyz=3, shs=48, opacity=1, scaling=3, rotation=4, feature=3)


def _inputs(P, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    raw = dict(xyz=r(P, 3) * 2.0, shs=r(P, 48) * 0.3, opacity=r(P, 1) * 2.0, scaling=r(P, 3) * 0.8 - 2.0, rotation=r(P, 4), feature=r(P, 3))
    raw["scaling"][: P // 10] += 3.5  # a few large ones: split and proximity candidates
    m = {k: r(*t.shape) * 1e-3 for k, t in raw.items()}
    v = {k: (r(*t.shape) * 1e-3) ** 2 for k, t in raw.items()}
    denom = torch.randint(0, 6, (P, 1), generator=g).float()   # zeros -> NaN gradients, as for never-visible Gaussians
    accum = torch.rand(P, 1, generator=g) * denom * 2e-4
    return raw, m, v, accum, denom


@pytest.mark.parametrize("iteration,extent,max_screen_size,min_opacity", [(1000, 0.05, None, 0.05), (1000, 3.0, 20, 0.3), (3000, 3.0, 20, 0.3),
                                                                          (400, 3.0, None, 0.3)])
def test_row_logic_matches_the_restated_reference_procedure(iteration, extent, max_screen_size, min_opacity):
    from b200gs.densify_logic import densify_rows
    from b200gs.schedule import DEFAULTS
    from oracle import train_torch as tt
    P = 600
    raw, m, v, accum, denom = _inputs(P, 11)
    thr = float(torch.quantile((accum / denom.clamp_min(1)).squeeze(), 0.85))

    sh = lambda t: t.view(P, 16, 3)
    model = tt.DensifyModel(dict(xyz=raw["xyz"], f_dc=sh(raw["shs"])[:, :1], f_rest=sh(raw["shs"])[:, 1:], opacity=raw["opacity"],
                                 scaling=raw["scaling"], rotation=raw["rotation"], feature=raw["feature"]), dict(DEFAULTS),
                            moments=dict(xyz=(m["xyz"], v["xyz"]), f_dc=(sh(m["shs"])[:, :1], sh(v["shs"])[:, :1]),
                                         f_rest=(sh(m["shs"])[:, 1:], sh(v["shs"])[:, 1:]), opacity=(m["opacity"], v["opacity"]),
                                         scaling=(m["scaling"], v["scaling"]), rotation=(m["rotation"], v["rotation"]),
                                         feature=(m["feature"], v["feature"])))
    model.xyz_gradient_accum, model.denom = accum.clone(), denom.clone()
    model.densify_and_prune(thr, min_opacity, extent, max_screen_size, iteration, generator=torch.Generator().manual_seed(77))

    knn = lambda xyz: tuple(t if i == 0 else t.to(torch.int32) for i, t in enumerate(tt.DensifyModel.dist_knn3(xyz.contiguous())))
    nr, nm, nv = densify_rows(raw, m, v, accum.clone(), denom.clone(), widths=WIDTHS, max_grad=thr, min_opacity=min_opacity, extent=extent,
                              max_screen_size=max_screen_size, iteration=iteration, knn3=knn, generator=torch.Generator().manual_seed(77))
    newP = model.p["xyz"].shape[0]
    assert nr["xyz"].shape[0] == newP and newP != P
    ref = dict(xyz=model.p["xyz"], shs=torch.cat((model.p["f_dc"], model.p["f_rest"]), 1).reshape(newP, 48), opacity=model.p["opacity"],
               scaling=model.p["scaling"], rotation=model.p["rotation"], feature=model.p["feature"])
    for k in WIDTHS:
        assert torch.equal(nr[k], ref[k].detach()), k
    st = {model._key(g): model.optimizer.state[g["params"][0]] for g in model.optimizer.param_groups}
    for key, mine in (("exp_avg", nm), ("exp_avg_sq", nv)):
        want = dict(xyz=st["xyz"][key], shs=torch.cat((st["f_dc"][key], st["f_rest"][key]), 1).reshape(newP, 48), opacity=st["opacity"][key],
                    scaling=st["scaling"][key], rotation=st["rotation"][key], feature=st["feature"][key])
        for k in WIDTHS:
            assert torch.equal(mine[k], want[k]), (key, k)


def test_module_needs_no_native_library():
    import importlib
    src = open(os.path.join(ROOT, "sdp-gs_b200", "b200gs", "densify_logic.py")).read()
    assert "_lib" not in src and "ctypes" not in src and "oracle" not in src
    importlib.import_module("b200gs.densify_logic")

"""Row logic of densify_and_prune (scene/gaussian_model.py:513-608: densify_and_clone, densify_and_split, proximity, prune_points
with cat_tensors_to_optimizer / _prune_optimizer / densification_postfix) as one pure-torch function over the trainer's tensors.
No native library is involved: it runs on whatever device its inputs live on, which is how tests/test_densify_logic.py checks
it row for row against the restated reference procedure on the CPU.

The logic works on three packed matrices [rows, 62] (parameters, first and second Adam moments; columns in the order of
`widths`) instead of 18 per-group tensors, and on row indices taken once per mask: the same gathers and concatenations the
reference performs per tensor (same rows, same order, same torch.normal draws), in a third of the torch calls and with one
host read per mask instead of one per indexed tensor."""
from __future__ import annotations

import torch


def build_rotation(r):
    """utils/general_utils.py:88-109"""
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    R = torch.zeros((q.size(0), 3, 3), device=r.device)
    r_, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r_ * z); R[:, 0, 2] = 2 * (x * z + r_ * y)
    R[:, 1, 0] = 2 * (x * y + r_ * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r_ * x)
    R[:, 2, 0] = 2 * (x * z - r_ * y); R[:, 2, 1] = 2 * (y * z + r_ * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def densify_rows(raw, m, v, accum, denom, *, widths, max_grad, min_opacity, extent, max_screen_size, iteration, knn3,
                 percent_dense=0.01, prune_from_iter=500, N=2, generator=None, proximity_until_iter=2000):
    """raw / m / v: dicts name -> [P, w] (parameters before activation, Adam moments), `widths`: ordered dict name -> w with
    the names xyz, shs, opacity, scaling, rotation, feature; accum, denom: [P, 1] densification statistics; knn3(xyz) ->
    (mean squared distance to the 3 nearest neighbours [P], their indices [P, 3]).  Returns the new (raw, m, v) as dicts of
    column views of three packed matrices; new rows carry zero moments."""
    dev = raw["xyz"].device
    off, c = {}, 0
    for k, w in widths.items():
        off[k] = (c, c + w); c += w
    R = torch.cat([raw[k] for k in widths], dim=1)
    M = torch.cat([m[k] for k in widths], dim=1)
    V = torch.cat([v[k] for k in widths], dim=1)
    col = lambda T, k: T[:, off[k][0]:off[k][1]]
    grads = accum / denom
    grads = torch.where(grads.isnan(), torch.zeros_like(grads), grads)  # grads[grads.isnan()] = 0.0 without the host read
    get_scaling = lambda: torch.exp(col(R, "scaling"))
    rows_of = lambda mask: mask.nonzero().squeeze(1)

    def cat(newR):  # cat_tensors_to_optimizer + densification_postfix: new rows start with zero moments
        nonlocal R, M, V
        z = torch.zeros_like(newR)
        R, M, V = torch.cat((R, newR), dim=0), torch.cat((M, z), dim=0), torch.cat((V, z), dim=0)

    def prune(mask):  # prune_points: keep ~mask
        nonlocal R, M, V
        if iteration > prune_from_iter:
            keep = rows_of(~mask)
            R, M, V = R.index_select(0, keep), M.index_select(0, keep), V.index_select(0, keep)

    # densify_and_clone
    sel = torch.where(torch.norm(grads, dim=-1) >= max_grad, True, False)
    sel = torch.logical_and(sel, torch.max(get_scaling(), dim=1).values <= percent_dense * extent)
    cat(R.index_select(0, rows_of(sel)))
    # densify_and_split
    n_init = R.shape[0]
    padded = torch.zeros((n_init,), device=dev)
    padded[:grads.shape[0]] = grads.squeeze()
    sel = torch.where(padded >= max_grad, True, False)
    sel = torch.logical_and(sel, torch.max(get_scaling(), dim=1).values > percent_dense * extent)
    idx = rows_of(sel)
    parents = R.index_select(0, idx)
    scal = torch.exp(col(parents, "scaling"))
    stds = scal.repeat(N, 1)
    means = torch.zeros((stds.size(0), 3), device=dev)
    samples = torch.normal(mean=means, std=stds, generator=generator)
    rots = build_rotation(col(parents, "rotation")).repeat(N, 1, 1)
    newR = parents.repeat(N, 1)
    col(newR, "xyz").copy_(torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + col(parents, "xyz").repeat(N, 1))
    col(newR, "scaling").copy_(torch.log(scal.repeat(N, 1) / (0.8 * N)))
    cat(newR)
    prune(torch.cat((sel, torch.zeros(N * int(idx.numel()), device=dev, dtype=bool))))
    # proximity
    if iteration < proximity_until_iter and R.shape[0] >= 4:
        dist, nn = knn3(col(R, "xyz"))
        sel = torch.logical_and(dist > 5.0 * extent, torch.max(get_scaling(), dim=1).values > extent)
        src = rows_of(sel)
        if int(src.numel()) > 0:
            idx = nn.index_select(0, src).reshape(-1).long()
            source = col(R, "xyz").index_select(0, src).repeat(1, 3, 1).reshape(-1, 3)  # the reference's own pairing (sources tiled, targets grouped)
            newR = R.index_select(0, idx)  # opacity, scaling and feature of the neighbour
            col(newR, "xyz").copy_((source + col(newR, "xyz")) / 2)
            col(newR, "shs").zero_()
            col(newR, "rotation").zero_()
            newR[:, off["rotation"][0]] = 1
            cat(newR)
    # prune (max_radii2D was reset by densification_postfix, so the screen-size test sees zeros, as in the reference)
    mask = (torch.sigmoid(col(R, "opacity")) < min_opacity).squeeze(1)
    if max_screen_size:
        big_vs = torch.zeros((R.shape[0],), device=dev) > max_screen_size
        big_ws = get_scaling().max(dim=1).values > 0.1 * extent
        mask = torch.logical_or(torch.logical_or(mask, big_vs), big_ws)
    prune(mask)
    return ({k: col(R, k) for k in widths}, {k: col(M, k) for k in widths}, {k: col(V, k) for k in widths})

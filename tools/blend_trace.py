"""Developer tool (GPU box): per-unit timeline of the blend kernels from a -DB200GS_BLEND_TRACE build.
    B200GS_LIB=variants/libb200gs_trace.so python tools/blend_trace.py [--workload llff_fern_3view]
Prints, for the forward and the backward kernel of one view: kernel span, unit-duration statistics, how many warps are
still busy as the kernel drains, and the cost model fit (duration vs rounds / survivors)."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sdp-gs_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from b200gs import _lib  # noqa: E402
from b200gs import rasterizer as rz  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="llff_fern_3view")
ap.add_argument("--out", default=None)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
wl = bench.Workload(a.workload, "extended", dev, None)
cam = wl.cams[0]
P = wl.scene.P
rs = wl.settings(cam, P)
res = rz._forward_impl(rs, wl.devt["means3D"], wl.devt["shs"], None, wl.devt["opacities"], wl.devt["scales"], wl.devt["rotations"],
                       None, None, wl.devt.get("features"), getattr(rs, "confidence", None), True)
cap = int(res[0] * 1.25) + 1024
s = rz.RasterSession(wl.settings(cam, P), means3D=wl.devt["means3D"], opacities=wl.devt["opacities"], shs=wl.devt["shs"],
                     scales=wl.devt["scales"], rotations=wl.devt["rotations"], language_feature_precomp=wl.devt.get("features"),
                     extended=True, capacity=cap, with_backward=True)
s.cot["color"].copy_(wl.cot[0][0]); s.cot["depth"].copy_(wl.cot[0][1]); s.cot["alpha"].copy_(wl.cot[0][2]); s.cot["feature"].copy_(wl.cot[0][3])
for _ in range(5):
    s.step()
torch.cuda.synchronize()
fn = _lib.lib.b200gs_debug_blend_trace
fn.argtypes = [C.c_void_p, C.c_int]
out = {}
for d, name in ((0, "forward"), (1, "backward")):
    buf = np.zeros((1 << 16, 8), dtype=np.uint32)
    assert fn(buf.ctypes.data, d) == 0
    r = buf[buf[:, 3] != 0]
    t0 = r[:, 2].astype(np.int64); t1 = r[:, 3].astype(np.int64)
    base = t0.min()
    t0 -= base; t1 -= base
    t1[t1 < t0] += 1 << 32
    dur = (t1 - t0) / 1000.0
    span = t1.max() / 1000.0
    n, rounds, surv, cyc = r[:, 4].astype(float), r[:, 5].astype(float), r[:, 6].astype(float), r[:, 7].astype(float)
    ends = np.sort(t1) / 1000.0
    busy = lambda t: int(((t0 / 1000.0 <= t) & (t1 / 1000.0 > t)).sum())
    A = np.stack([np.ones_like(rounds), rounds, surv], 1)
    coef, *_ = np.linalg.lstsq(A, dur, rcond=None)
    sm = r[:, 1]
    per_sm_end = np.array([t1[sm == i].max() / 1000.0 for i in np.unique(sm)])
    top = np.argsort(-dur)[:8]
    out[name] = dict(units=int(len(r)), span_us=round(span, 1), sum_unit_us=round(float(dur.sum()), 0),
                     unit_us=dict(mean=round(float(dur.mean()), 2), p50=round(float(np.median(dur)), 2), p90=round(float(np.percentile(dur, 90)), 2),
                                  p99=round(float(np.percentile(dur, 99)), 2), max=round(float(dur.max()), 2)),
                     done_at_us={q: round(float(ends[int(len(ends) * q / 100) - 1]), 1) for q in (50, 75, 90, 95, 99, 100)},
                     busy_warps_at={int(f * 100): busy(span * f) for f in (0.1, 0.25, 0.5, 0.6, 0.7, 0.8, 0.9, 0.95)},
                     sm_finish_us=dict(min=round(float(per_sm_end.min()), 1), mean=round(float(per_sm_end.mean()), 1), max=round(float(per_sm_end.max()), 1)),
                     fit_us=dict(const=round(float(coef[0]), 3), per_round=round(float(coef[1]), 3), per_survivor=round(float(coef[2]), 4)),
                     rounds=dict(mean=round(float(rounds.mean()), 1), max=float(rounds.max())), survivors=dict(mean=round(float(surv.mean()), 1), max=float(surv.max())),
                     list_len=dict(mean=round(float(n.mean()), 1), max=float(n.max())),
                     heaviest=[dict(unit=int(r[i, 0]), start=round(float(t0[i]) / 1000, 1), us=round(float(dur[i]), 1), rounds=int(rounds[i]), surv=int(surv[i]), n=int(n[i]),
                                    cycles=int(cyc[i])) for i in top])
if hasattr(_lib.lib, "b200gs_debug_blend_trace2"):
    fn2 = _lib.lib.b200gs_debug_blend_trace2
    fn2.argtypes = [C.c_void_p]
    b2 = np.zeros((1 << 16, 4), dtype=np.uint32)
    assert fn2(b2.ctypes.data) == 0
    for h in out["forward"]["heaviest"]:
        h["cycles_process_wait_stage_batches"] = [int(x) for x in b2[h["unit"]]]
print(json.dumps(out))
if a.out:
    open(a.out, "w").write(json.dumps(out, indent=1))

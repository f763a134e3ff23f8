"""Developer tool (GPU box): per-stage CUDA-event times and the graph-replay step time of one workload.
    python tools/stage_times.py [--workload llff_fern_3view] [--mode extended] [--steps 20] [--fwd-only] [--views 3]
Prints one JSON line.  Honours B200GS_LIB (A/B builds)."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sdp-gs_b200"))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="llff_fern_3view")
ap.add_argument("--mode", default="extended")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--fwd-only", action="store_true")
ap.add_argument("--views", type=int, default=3)
ap.add_argument("--P", type=int, default=None)
ap.add_argument("--no-graph", action="store_true", help="eager launches only (for ncu)")
a = ap.parse_args()
from b200gs import _lib  # noqa: E402
from b200gs import rasterizer as rz  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
wl = bench.Workload(a.workload, a.mode, dev, a.P)
wl.cams = wl.cams[:a.views]
P, ext = wl.scene.P, wl.extended
flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
flush = lambda: flush_buf.zero_()
Ls = []
for cam in wl.cams:
    rs = wl.settings(cam, P)
    res = rz._forward_impl(rs, wl.devt["means3D"], wl.devt["shs"], None, wl.devt["opacities"], wl.devt["scales"], wl.devt["rotations"],
                           None, None, wl.devt.get("features"), getattr(rs, "confidence", None), ext)
    Ls.append(res[0])
    del res
cap = int(max(Ls) * 1.25) + 1024
sessions = []
for vi, cam in enumerate(wl.cams):
    s = rz.RasterSession(wl.settings(cam, P), means3D=wl.devt["means3D"], opacities=wl.devt["opacities"], shs=wl.devt["shs"],
                         scales=wl.devt["scales"], rotations=wl.devt["rotations"], language_feature_precomp=wl.devt.get("features"),
                         extended=ext, capacity=cap, with_backward=not a.fwd_only)
    if not a.fwd_only:
        s.cot["color"].copy_(wl.cot[vi][0])
        if ext:
            s.cot["depth"].copy_(wl.cot[vi][1]); s.cot["alpha"].copy_(wl.cot[vi][2]); s.cot["feature"].copy_(wl.cot[vi][3])
    sessions.append(s if a.no_graph else s.capture())
n = len(sessions)
ms, _ = bench.event_loop(a.steps, 3, lambda i: (sessions[i % n].step() if a.no_graph else sessions[i % n].replay()), flush, 1)
_lib.lib.b200gs_profile_enable(1)
for i in range(a.steps + 2):
    if i == 2:
        _lib.lib.b200gs_profile_read(None, None, 1)
    flush()
    sessions[i % n].step()
arr = (C.c_double * 10)(); cnt = (C.c_int64 * 10)()
_lib.lib.b200gs_profile_read(arr, cnt, 1)
_lib.lib.b200gs_profile_enable(0)
for s in sessions:
    assert s.status()[1] == 0
print(json.dumps(dict(lib=os.path.basename(_lib.LIB_PATH), workload=a.workload, mode=a.mode, L=int(sum(Ls) / len(Ls)), step_ms=round(ms / a.steps, 4),
                      stages_us={bench.STAGES[i]: round(1000 * arr[i] / a.steps, 1) for i in range(10)})))

"""torchrun --nproc-per-node N tools/check_exchange.py [--P 100000] : the fused gradient exchange of the image-parallel
step (reduce-scatter pushed by the preprocess-backward kernel + b200gs_gather_reduce_f32) against the plain path
(local gradients, NCCL all-reduce) on real rasterizer gradients: max error, bit-identical on all ranks, and the time of
one step [forward + backward + exchange] both ways (CUDA-graph replays, max over ranks)."""
import argparse
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sdp-gs_b200"))
import numpy as np
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--P", type=int, default=100_000)
ap.add_argument("--workload", default="llff_fern_3view")
ap.add_argument("--iters", type=int, default=30)
a = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import bench
from b200gs import parallel
from b200gs import rasterizer as rz

rank, n = parallel.world()
dev = torch.device("cuda", local)
wl = bench.Workload(a.workload, "extended", dev, a.P)
P = wl.scene.P
vi = rank % len(wl.cams)
cam = wl.cams[vi]
res = rz._forward_impl(wl.settings(cam, P), wl.devt["means3D"], wl.devt["shs"], None, wl.devt["opacities"], wl.devt["scales"],
                       wl.devt["rotations"], None, None, wl.devt["features"], None, True)
cap = int(res[0] * 1.3) + 1024
del res


def make(bucket, scatter):
    go = dict(means3D=bucket.segment("xyz"), shs=bucket.segment("shs"), opacities=bucket.segment("opacity"), scales=bucket.segment("scaling"),
              rotations=bucket.segment("rotation"), features=bucket.segment("language_feature"))
    s = rz.RasterSession(wl.settings(cam, P), means3D=wl.devt["means3D"], opacities=wl.devt["opacities"], shs=wl.devt["shs"],
                         scales=wl.devt["scales"], rotations=wl.devt["rotations"], language_feature_precomp=wl.devt["features"],
                         extended=True, capacity=cap, grads_out=go, grad_scatter=bucket.scatter_descriptor() if scatter else None)
    s.cot["color"].copy_(wl.cot[vi][0]); s.cot["depth"].copy_(wl.cot[vi][1]); s.cot["alpha"].copy_(wl.cot[vi][2]); s.cot["feature"].copy_(wl.cot[vi][3])
    return s


os.environ["B200GS_ALLREDUCE"] = "auto"
fused = parallel.FusedGradBuffer(P, dev)
assert fused.fused_exchange, "symmetric memory unavailable"
plain = parallel.FusedGradBuffer(P, dev, symmetric=False)
sf, sp = make(fused, True), make(plain, False)
fused.flat.fill_(float("nan"))  # every word of the result must be written by the exchange
sf.step(); fused.gather_reduce()
sp.step()
ref = plain.grads_flat.clone()
dist.all_reduce(ref)
torch.cuda.synchronize()
names = ("xyz", "shs", "opacity", "scaling", "rotation", "language_feature")
errs = {}
Pp_f, Pp_p = fused.Pp, plain.Pp
c = 0
worst = 0.0
for k, w in parallel.SLOTS[:6]:
    mine = fused.segment(k).reshape(P, -1)
    want = ref[c * Pp_p:c * Pp_p + w * P].view(P, w)
    d = float((mine - want).abs().max()); sc = float(want.abs().max())
    errs[k] = d / max(sc, 1e-30)
    worst = max(worst, errs[k])
    c += w
nan_left = bool(torch.isnan(fused.grads_flat).any())
chk = fused.grads_flat.double().nan_to_num().sum().reshape(1).clone()
lst = [torch.zeros_like(chk) for _ in range(n)]
dist.all_gather(lst, chk)
ident = all(float(x) == float(lst[0]) for x in lst)


def timeit(fn, iters=a.iters):
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize(); dist.barrier()
    with torch.cuda.graph(g):
        fn()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


os.environ["B200GS_ALLREDUCE"] = "p2p"
twoshot = parallel.FusedGradBuffer(P, dev)
st = make(twoshot, False)
t_alone = timeit(lambda: sp.step())
t_fused = timeit(lambda: (sf.step(), fused.gather_reduce()))
t_two = timeit(lambda: (st.step(), twoshot.all_reduce()))
t_scatter_only = timeit(lambda: sf.step())
t_gather_only = timeit(lambda: fused.gather_reduce(chained=False))
t_two_only = timeit(lambda: twoshot.all_reduce())
if rank == 0:
    print(f"EXCHANGE world={n} P={P} max rel err vs NCCL sum={worst:.3e} " + " ".join(f"{k}={v:.1e}" for k, v in errs.items())
          + f" nan_left={nan_left} identical on all ranks={ident} | step alone {t_alone:.1f} us, + fused exchange {t_fused:.1f} us, "
          f"+ two-shot all-reduce {t_two:.1f} us | step with scatter only {t_scatter_only:.1f} us, gather kernel alone {t_gather_only:.1f} us, "
          f"two-shot kernel alone {t_two_only:.1f} us", flush=True)
dist.barrier()
dist.destroy_process_group()

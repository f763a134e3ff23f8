"""Host-side profile (cProfile) of the public autograd path: where the Python time of one forward+backward goes."""
import cProfile, os, pstats, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sdp-gs_b200"))
import numpy as np, torch
import bench
from b200gs import rasterizer as rz
from b200gs.hostio import PinnedFeeder
from diff_gaussian_rasterization import GaussianRasterizer

dev = torch.device("cuda", 0)
wl = bench.Workload("llff_fern_3view", "extended", dev)
P = wl.scene.P
sets = [wl.settings(c, P) for c in wl.cams]
feeder = PinnedFeeder(wl.host, dev)
rz.set_binning_capacity("auto")

def step(i):
    vi = i % 3
    t = {k: v.requires_grad_(True) for k, v in feeder.next().items()}
    means2D = torch.zeros((P, 3), device=dev, requires_grad=True)
    outs = GaussianRasterizer(sets[vi])(means3D=t["means3D"], means2D=means2D, opacities=t["opacities"], shs=t["shs"], scales=t["scales"],
                                        rotations=t["rotations"], language_feature_precomp=t["features"])
    torch.autograd.backward(list(outs[:4]), list(wl.cot[vi]))
    feeder.done()
    return float(outs[0].sum().item())

for i in range(10): step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(200): step(i)
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) * 5)
# host-only cost: no sync inside the loop
def step_nosync(i):
    vi = i % 3
    t = {k: v.requires_grad_(True) for k, v in feeder.next().items()}
    means2D = torch.zeros((P, 3), device=dev, requires_grad=True)
    outs = GaussianRasterizer(sets[vi])(means3D=t["means3D"], means2D=means2D, opacities=t["opacities"], shs=t["shs"], scales=t["scales"],
                                        rotations=t["rotations"], language_feature_precomp=t["features"])
    torch.autograd.backward(list(outs[:4]), list(wl.cot[vi]))
    feeder.done()
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(200): step_nosync(i)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue ms/step", (t1 - t0) * 5, "incl. drain", (time.perf_counter() - t0) * 5)
pr = cProfile.Profile(); pr.enable()
for i in range(300): step_nosync(i)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)

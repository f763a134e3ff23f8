"""Developer tool (GPU box, under compute-sanitizer): one eager forward+backward and a few persistent-session steps on the tiny case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sdp-gs_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import helpers
from helpers import case_inputs, case_cotangents, run_product
from test_session_gpu import _session
for name in ("tiny_sh3_ext", "tiny_sh0_mod"):
    inp = case_inputs(name); cot = case_cotangents(inp)
    p = run_product(inp, True, cot)
    s = _session(inp, torch.device("cuda", 0), capacity=int(p["num_rendered"] * 1.5) + 64)
    for _ in range(3):
        s.step()
    torch.cuda.synchronize()
    print(name, "ok", p["num_rendered"], s.status())

"""Developer tool (GPU box): product vs CPU oracle vs reference CUDA on every parity case, one report."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import helpers
from helpers import CASES, case_inputs, case_cotangents, run_oracle, run_product, run_reference, rel_err, INT_KEYS
from oracle import ref_cuda

def cmp(tag, a, b, inp):
    bad = []
    for k in INT_KEYS:
        if k in a and k in b:
            same = np.array_equal(np.asarray(a[k]), np.asarray(b[k]))
            if not same:
                x, y = np.asarray(a[k]), np.asarray(b[k])
                nd = int((x != y).sum()) if x.shape == y.shape else -1
                bad.append(f"{k}:{nd}")
    for k in ("depths", "means2D"):
        if k in a and k in b and not np.array_equal(helpers.bits(a[k]), helpers.bits(b[k])):
            bad.append(f"{k}bits:{int((helpers.bits(a[k]) != helpers.bits(b[k])).sum())}")
    fl = {}
    for k in ("color", "depth", "alpha", "feature", "final_T", "conic_opacity", "rgb"):
        if k in a and k in b and a[k] is not None and b[k] is not None:
            fl[k] = float(np.abs(np.asarray(a[k], np.float64) - np.asarray(b[k], np.float64)).max())
    gr = {}
    if "grads" in a and "grads" in b:
        for k, v in a["grads"].items():
            w = b["grads"].get(k)
            if v is not None and w is not None and np.asarray(v).size:
                gr[k] = rel_err(np.asarray(v).reshape(-1), np.asarray(w).reshape(-1))
    print(f"  {tag}: int-mismatch={bad or 'none'}")
    print("     maxabs " + " ".join(f"{k}={v:.2e}" for k, v in fl.items()))
    if gr: print("     grad rel " + " ".join(f"{k}={v:.2e}" for k, v in gr.items()))

names = sys.argv[1:] or list(CASES)
for name in names:
    inp = case_inputs(name); cot = case_cotangents(inp)
    print(f"== {name} P={inp['means3D'].shape[0]} {inp['cam'].width}x{inp['cam'].height} ext={inp['extended']}")
    t = time.time(); o = run_oracle(inp, True, cot); to = time.time() - t
    t = time.time(); p = run_product(inp, True, cot); tp = time.time() - t
    print(f"  L={o['num_rendered']} visible={(o['radii']>0).sum()} oracle {to:.2f}s product {tp:.2f}s")
    cmp("product vs oracle", p, o, inp)
    if ref_cuda.available():
        r = run_reference(inp, True, cot)
        cmp("product vs reference", p, r, inp)
        cmp("oracle vs reference", o, r, inp)

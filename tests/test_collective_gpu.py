"""The NVLink all-reduce kernel (include/b200gs_collective.h) against NCCL on >= 2 GPUs: bit-identical sums on every rank,
statistics segment untouched.  Skipped on a single-GPU box (the symbol/argument checks below still run on CPU)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

import helpers


def test_collective_header_symbols_are_exported():
    from b200gs import _lib
    src = open(os.path.join(helpers.ROOT, "include", "b200gs_collective.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(b200gs_[a-z_0-9]+)\s*\(", src)))
    assert sorted(_lib.COLLECTIVE_EXPORTS) == names
    raw = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n)
    assert _lib.lib.b200gs_allreduce_flag_words(8) >= 2 * 64 * 8
    # argument validation happens before any CUDA call
    assert _lib.lib.b200gs_allreduce_sum_f32(None, None, None, 0, 16, 0, 2, None) == -1
    assert _lib.lib.b200gs_allreduce_sum_f32(C.c_void_p(8), C.c_void_p(8), None, 0, 6, 0, 2, None) == -1  # n not a multiple of 4
    assert _lib.lib.b200gs_allreduce_sum_f32(C.c_void_p(8), C.c_void_p(8), None, 0, 16, 0, 1, None) == 0  # world 1: nothing to do


@pytest.mark.gpu
def test_allreduce_matches_nccl_on_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", os.path.join(helpers.ROOT, "tools", "check_allreduce.py")],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env).stdout
    lines = [l for l in out.splitlines() if l.startswith("P=")]
    assert len(lines) >= 2, out[-2000:]
    for l in lines:
        assert "max|err|=0.000e+00" in l and "stats untouched=True" in l and "identical on all ranks=True" in l, l


@pytest.mark.gpu
def test_fused_gradient_exchange_on_two_gpus():
    """reduce-scatter inside the preprocess-backward kernel + b200gs_gather_reduce_f32 == NCCL sum of the ranks' local gradients
    (fp32 reassociation only), every word of the result written, bit-identical on both ranks."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29534", os.path.join(helpers.ROOT, "tools", "check_exchange.py"), "--P", "30000", "--iters", "5"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env).stdout
    lines = [l for l in out.splitlines() if l.startswith("EXCHANGE")]
    assert len(lines) == 1, out[-3000:]
    m = re.search(r"max rel err vs NCCL sum=([0-9.e+-]+)", lines[0])
    assert m and float(m.group(1)) <= 1e-5, lines[0]
    assert "nan_left=False" in lines[0] and "identical on all ranks=True" in lines[0], lines[0]

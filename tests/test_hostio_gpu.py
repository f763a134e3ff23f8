"""PinnedFeeder (b200gs/hostio.py): every step's tensors are the host arrays as they were when that step's upload was
enqueued (the upload of step i+1 is enqueued by next() of step i), double buffering never hands out a buffer that is
being overwritten, and the leaves come back with .grad cleared."""
import numpy as np
import pytest
import torch

import helpers  # noqa: F401


@pytest.mark.gpu
def test_feeder_delivers_each_steps_host_data_in_order():
    from b200gs.hostio import PinnedFeeder
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    arrays = dict(a=rng.normal(size=(1000, 3)).astype(np.float32), b=rng.normal(size=(1000, 16, 3)).astype(np.float32),
                  c=rng.normal(size=(1000, 1)).astype(np.float32))
    f = PinnedFeeder(arrays, dev)
    assert f.nbytes == sum(v.nbytes for v in arrays.values())
    busy = torch.empty(64 << 20, device=dev)
    seen = []
    for step in range(6):
        t = f.next()  # enqueues the upload for step + 1 from the CURRENT host contents
        assert all(v.requires_grad and v.grad is None and v.is_leaf for v in t.values())
        busy.normal_()  # keep the compute stream busy while the next upload runs
        seen.append(float(t["c"].detach().sum().item()))
        (t["a"].sum() + t["b"].sum()).backward()
        assert t["a"].grad is not None
        f.done()
        f.host_view("c").fill_(float(step + 1))  # visible from step + 2 on (step + 1's upload is already in flight / done)
    torch.cuda.synchronize()
    base = float(arrays["c"].sum())
    assert abs(seen[0] - base) < 1e-2 and abs(seen[1] - base) < 1e-2
    for step in range(2, 6):
        assert seen[step] == 1000.0 * (step - 1), seen

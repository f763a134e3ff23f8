"""torchrun --nproc-per-node N tools/check_allreduce.py : the custom NVLink all-reduce against NCCL (result + time)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sdp-gs_b200"))
import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from b200gs import parallel
rank, n = parallel.world()
for P in (100_000, 6_000_000 if "--big" in sys.argv else 300_000):
    for mode in ("p2p", "auto"):
        os.environ["B200GS_ALLREDUCE"] = mode
        b = parallel.FusedGradBuffer(P, torch.device("cuda", local))
        g = torch.Generator(device="cuda").manual_seed(100 + rank)
        src = torch.randn(b.flat.shape, generator=g, device="cuda")
        b.flat.copy_(src)
        nred = b.grads_flat.numel()
        ref = src[:nred].clone()
        dist.all_reduce(ref)
        b.all_reduce()
        torch.cuda.synchronize()
        err = float((b.grads_flat - ref).abs().max())
        same_stats = bool(torch.equal(b.stats_flat, src[nred:]))
        # every rank must hold bit-identical sums
        chk = b.grads_flat.double().sum().reshape(1).clone()
        lst = [torch.zeros_like(chk) for _ in range(n)]
        dist.all_gather(lst, chk)
        ident = all(float(x) == float(lst[0]) for x in lst)
        def timeit(fn, iters=30):
            for _ in range(5): fn()
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters): fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters * 1e3
        t_mine = timeit(lambda: b.all_reduce())
        t_nccl = timeit(lambda: dist.all_reduce(b.grads_flat))
        if rank == 0:
            mb = 62 * P * 4 / 1e6
            print(f"P={P} ({mb:.1f} MB) mode={mode} multicast={bool(b._symm and b._symm.multicast)} max|err|={err:.3e} stats untouched={same_stats} "
                  f"identical on all ranks={ident}  custom {t_mine:.1f} us  nccl {t_nccl:.1f} us", flush=True)
        del b
dist.barrier()
dist.destroy_process_group()

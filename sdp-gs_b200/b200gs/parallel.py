"""Multi-GPU drivers for the two ways the rasterizer path shards on one 8xB200 NVSwitch box
(SURVEY.md §8e).  One process per GPU, `torch.distributed` for the plumbing (NCCL on GPUs; the
same code runs under gloo on CPU for the host-logic tests).

* View-sharded rendering (model: render.py:38-39): parameters replicated, rank r renders views
  r, r+n, ...; no data-path collective.
* Image-parallel training (train.py:93-231 with one view per rank): each rank renders and
  back-propagates its own views; the per-Gaussian gradients of all ranks are combined with ONE
  all-reduce over a fused buffer of P*64 f32 (62 parameter-gradient floats per Gaussian for the Adam
  groups of scene/gaussian_model.py:228-237 -- xyz 3, f_dc 3 + f_rest 45, opacity 1, scaling 3,
  rotation 4, language_feature 3 -- plus the two densification statistics xyz_gradient_accum and
  denom, scene/gaussian_model.py:610-612), and max_radii2D with one MAX all-reduce, so that the Adam
  step and densify/prune are applied identically on every rank.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

# segments of the fused buffer, floats per Gaussian
SLOTS = (("xyz", 3), ("shs", 48), ("opacity", 1), ("scaling", 3), ("rotation", 4), ("language_feature", 3),
         ("xyz_gradient_accum", 1), ("denom", 1))
FUSED_WIDTH = sum(w for _, w in SLOTS)  # 64 floats = 256 B per Gaussian
assert FUSED_WIDTH == 64


def shard_views(n_views, rank, world_size):
    """Indices of the views rank `rank` owns (round-robin, like `views[r::n]`)."""
    return list(range(rank, n_views, world_size))


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class _SymmetricBuffer:
    """A float32 buffer allocated at the same offset of a peer-mapped allocation on every rank
    (torch.distributed._symmetric_memory supplies the allocation, the handle exchange and the NVLS multicast mapping;
    the reduction itself is b200gs_allreduce_sum_f32)."""

    @classmethod
    def create(cls, n_floats, device, mode):
        try:
            import torch.distributed._symmetric_memory as symm
            from ._lib import lib
            self = cls()
            rank, n = world()
            self.rank, self.world = rank, n
            self.tensor = symm.empty(n_floats, dtype=torch.float32, device=device)
            self.tensor.zero_()
            self.hdl = symm.rendezvous(self.tensor, dist.group.WORLD)
            words = int(lib.b200gs_allreduce_flag_words(n))
            self.flags = symm.empty(words, dtype=torch.int32, device=device)
            self.flags.zero_()
            self.fhdl = symm.rendezvous(self.flags, dist.group.WORLD)
            mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
            # NVLS multimem measured slower than plain P2P at 2 GPUs (85 vs 58 us for 24.8 MB): opt-in only
            self.multicast = mc if mode == "multimem" else 0
            if mode == "multimem" and not mc:
                raise RuntimeError("no NVLS multicast mapping for the symmetric buffer")
            torch.cuda.synchronize(device)
            dist.barrier()
            return self
        except Exception as ex:  # not fatal: NCCL does the same sum
            if os.environ.get("B200GS_ALLREDUCE", "auto") in ("p2p", "multimem"):
                raise
            import warnings
            warnings.warn(f"symmetric gradient buffer unavailable ({type(ex).__name__}: {ex}); using NCCL all-reduce")
            return None

    def all_reduce(self, offset_floats, n_floats):
        import ctypes as C
        from ._lib import check, lib
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib.b200gs_allreduce_sum_f32(C.c_void_p(int(self.hdl.buffer_ptrs_dev)), C.c_void_p(int(self.fhdl.buffer_ptrs_dev)),
                                           C.c_void_p(self.multicast) if self.multicast else None, C.c_int64(offset_floats),
                                           C.c_int64(n_floats), self.rank, self.world, stream))

    def add_staging(self, n_floats, device):
        """Second symmetric allocation: the per-source staging block the backward kernels of all ranks push into."""
        import torch.distributed._symmetric_memory as symm
        self.staging = symm.empty(n_floats, dtype=torch.float32, device=device)
        self.staging.zero_()
        self.shdl = symm.rendezvous(self.staging, dist.group.WORLD)
        torch.cuda.synchronize(device)
        dist.barrier()

    def gather_reduce(self, shard_rows, chained):
        import ctypes as C
        from ._lib import check, lib
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib.b200gs_gather_reduce_f32(C.c_void_p(int(self.shdl.buffer_ptrs_dev)), C.c_void_p(int(self.hdl.buffer_ptrs_dev)),
                                           C.c_void_p(int(self.fhdl.buffer_ptrs_dev)), C.c_int64(shard_rows), self.rank, self.world,
                                           1 if chained else 0, stream))


class FusedGradBuffer:
    """One flat f32 allocation of P*64 floats holding every per-Gaussian quantity of an image-parallel
    step as back-to-back contiguous segments ([P,3] xyz, [P,16,3] shs = f_dc + f_rest, [P,1] opacity, ...).
    The rasterizer backward can write straight into the segments (they are ordinary contiguous tensors), so
    combining all ranks' gradients is ONE all-reduce with no packing pass."""

    def __init__(self, P, device, sh_coeffs=16, symmetric=None):
        assert sh_coeffs == 16, "segment table assumes max_sh_degree = 3 (arguments/__init__.py:49)"
        self.P = P
        rank, n = world()
        # shard of the fused exchange: rank o owns Gaussians [o * Ps, (o+1) * Ps), Ps a multiple of the 128-Gaussian tile
        self.shard_rows = Ps = ((P + n - 1) // n + 127) // 128 * 128
        self.flat = None
        self._symm = None
        self.fused_exchange = False
        mode = os.environ.get("B200GS_ALLREDUCE", "auto")
        if symmetric is None:
            symmetric = n > 1 and torch.device(device).type == "cuda" and mode != "nccl"
        # segment stride: every segment starts 16-byte aligned whatever P is; n whole shards when the exchange is sharded
        Pp = self.Pp = n * Ps if (symmetric and mode in ("auto", "fused")) else (P + 3) // 4 * 4
        if symmetric:
            self._symm = _SymmetricBuffer.create(Pp * FUSED_WIDTH, device, mode)
            if self._symm is not None:
                self.flat = self._symm.tensor
                if mode in ("auto", "fused"):
                    # default exchange: reduce-scatter pushed by the backward kernel itself, gather by b200gs_gather_reduce_f32
                    self._symm.add_staging(Pp * FUSED_WIDTH, device)
                    self.fused_exchange = True
        if self.flat is None:
            self.flat = torch.zeros((Pp * FUSED_WIDTH,), dtype=torch.float32, device=device)
        self.max_radii2D = torch.zeros((P,), dtype=torch.int32, device=device)
        self.seg = {}
        c = 0
        for name, w in SLOTS:
            self.seg[name] = self.flat[c * Pp:c * Pp + w * P].view(P, w)
            c += w
        self.seg["shs"] = self.seg["shs"].view(P, 16, 3)

    def segment(self, name):
        return self.seg[name]

    def zero_(self):
        self.flat.zero_()
        self.max_radii2D.zero_()

    def scatter_descriptor(self):
        """(device array of staging base pointers, shard rows, rank, world) for RasterSession(grad_scatter=...), or None
        when the fused exchange is not available (single GPU, no symmetric memory, B200GS_ALLREDUCE=p2p|nccl)."""
        if not self.fused_exchange:
            return None
        return (int(self._symm.shdl.buffer_ptrs_dev), self.shard_rows, self._symm.rank, self._symm.world)

    def gather_reduce(self, chained=True):
        """Second half of the fused exchange (the first half ran inside the rasterizer backward): afterwards every rank's
        parameter-gradient segments hold the sum over ranks, bit-identical everywhere."""
        self._symm.gather_reduce(self.shard_rows, chained)

    def add_statistics(self, viewspace_grad, radii):
        """Densification statistics of one rendered view (train.py:218-221, scene/gaussian_model.py:610-612)."""
        vis = radii > 0
        self.seg["xyz_gradient_accum"][vis] += torch.norm(viewspace_grad[vis, :2], dim=-1, keepdim=True)
        self.seg["denom"][vis] += 1.0
        self.max_radii2D[vis] = torch.maximum(self.max_radii2D[vis], radii[vis].to(torch.int32))

    def accumulate_view(self, *, d_xyz, d_shs, d_opacity, d_scaling, d_rotation, d_feature, viewspace_grad, radii):
        """Add one rendered view's gradients and densification statistics."""
        self.seg["xyz"].add_(d_xyz)
        if d_shs is not None:
            self.seg["shs"].add_(d_shs)
        self.seg["opacity"].add_(d_opacity.reshape(-1, 1))
        self.seg["scaling"].add_(d_scaling)
        self.seg["rotation"].add_(d_rotation)
        if d_feature is not None:
            self.seg["language_feature"].add_(d_feature)
        self.add_statistics(viewspace_grad, radii)

    @property
    def grads_flat(self):
        """The 62 parameter-gradient floats per Gaussian (everything before the two statistics segments)."""
        return self.flat[: 62 * self.Pp]

    @property
    def stats_flat(self):
        return self.flat[62 * self.Pp:]

    def all_reduce(self, async_op=False, with_statistics=False):
        """SUM over ranks of the parameter gradients: ONE collective per training step.  The densification
        statistics are sums / maxima over iterations as well as over ranks, so they only need combining when
        densify_and_prune is about to read them (every `densification_interval` steps, train.py:223-225):
        pass with_statistics=True then, or call all_reduce_statistics()."""
        rank, n = world()
        if n == 1:
            return None
        if self._symm is not None and not async_op:
            # our own kernel over NVLink peer memory (include/b200gs_collective.h), on the current stream
            self._symm.all_reduce(0, (64 if with_statistics else 62) * self.Pp)
            if with_statistics:
                dist.all_reduce(self.max_radii2D, op=dist.ReduceOp.MAX)
            return None
        works = [dist.all_reduce(self.flat if with_statistics else self.grads_flat, op=dist.ReduceOp.SUM, async_op=True)]
        if with_statistics:
            works.append(dist.all_reduce(self.max_radii2D, op=dist.ReduceOp.MAX, async_op=True))
        if async_op:
            return works
        for w in works:
            w.wait()
        return None

    def all_reduce_statistics(self):
        rank, n = world()
        if n == 1:
            return
        w1 = dist.all_reduce(self.stats_flat, op=dist.ReduceOp.SUM, async_op=True)
        w2 = dist.all_reduce(self.max_radii2D, op=dist.ReduceOp.MAX, async_op=True)
        w1.wait()
        w2.wait()


def render_views_sharded(render_fn, n_views, gather=False):
    """Each rank calls `render_fn(view_index)` for the views it owns.  Returns {index: result}; with
    gather=True rank 0 additionally receives every rank's {index: result} (results must be picklable)."""
    rank, n = world()
    mine = {i: render_fn(i) for i in shard_views(n_views, rank, n)}
    if gather and n > 1:
        out = [None] * n if rank == 0 else None
        dist.gather_object(mine, out, dst=0)
        if rank == 0:
            merged = {}
            for d in out:
                merged.update(d)
            return merged
    return mine


def image_parallel_step(fwd_bwd_fn, view_indices, bucket):
    """One image-parallel training step: this rank runs `fwd_bwd_fn(view_index)` (which must return the
    kwargs of FusedGradBuffer.accumulate_view) for its share of `view_indices`, then the fused all-reduce.
    After return `bucket` holds the same totals on every rank."""
    rank, n = world()
    bucket.zero_()
    for k in shard_views(len(view_indices), rank, n):
        bucket.accumulate_view(**fwd_bwd_fn(view_indices[k]))
    bucket.all_reduce(with_statistics=True)
    return bucket

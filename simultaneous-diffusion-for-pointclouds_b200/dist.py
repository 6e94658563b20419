"""View sharding across GPUs (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

SURVEY.md 8(e): the score network is per-view (InstanceNorm statistics are per sample), so views are
independent for >99% of the FLOPs.  The cross-view block needs (1) every view of the same group and
(2) one scalar: the tooHigh gate is a max over ALL views of the call (KITTISampling.py:162).

Rank r owns the contiguous block of views [r*B/n, (r+1)*B/n).  Per step, when a group spans ranks (CUDA):
    update own views                                  (sdpc_langevin_update, tgt range = own block)
    pack: own updated planes + the max|x0| word       (sdpc_shard_pack -> the rank's slot of the gather buffer)
    ONE NCCL all-gather of the slots, in place        (512 KiB per view + 128 B; the max rides in the payload)
    unpack: other ranks' planes -> x, maxima folded   (sdpc_shard_unpack)
    z-buffers + correction for own target views       (sdpc_crossview_share, tgt range = own block)
When every group lives on one rank only the max is exchanged (1-float all-reduce).  The gloo / CPU path of the tests
keeps two plain collectives (all-reduce of the max, all-gather of the planes): same results, bit for bit.
"""
import torch
import torch.distributed as dist


class ViewShard:
    def __init__(self, n_views, group_size, process_group=None, replicated_noise=True):
        self.pg = process_group
        self.rank = dist.get_rank(process_group)
        self.world = dist.get_world_size(process_group)
        if n_views % self.world != 0:
            raise ValueError(f"n_views={n_views} must be divisible by world size {self.world}")
        self.per = n_views // self.world
        self.lo, self.hi = self.rank * self.per, (self.rank + 1) * self.per
        # a gather is needed only if some group spans more than one rank
        self.needs_gather = not (self.per % group_size == 0)
        self.replicated_noise = replicated_noise
        self._gbuf = None

    def exchange_description(self):
        if not self.needs_gather:
            return "1-float all-reduce(MAX) per step (tooHigh gate); every group lives on one rank"
        return (f"per step: ONE NCCL all-gather over NVLink of each rank's updated x planes ({self.per} view(s) x 512 KiB) "
                f"with its max|x0| word (tooHigh gate) appended; pack / unpack kernels on either side")

    def attach(self, run, x):
        run.tgt_first, run.tgt_count = self.lo, self.per

    def local(self, t):
        return t[self.lo:self.hi]

    def score(self, scorenet, x, labels, grad_full):
        grad_full[self.lo:self.hi] = scorenet(x[self.lo:self.hi].contiguous(), labels[self.lo:self.hi])
        return grad_full

    def step(self, run, p, b, x):
        run.update_only(p, b)
        if not p.share:
            return
        if self.needs_gather and x.is_cuda and hasattr(run.lib, "sdpc_shard_pack"):
            self._exchange_fused(run, x)
        else:
            mx = run.local_max()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.pg)
            run.merge_max(mx)
            if self.needs_gather:
                self._all_gather(x, x[self.lo:self.hi])
        run.share_only(p, b)

    def _exchange_fused(self, run, x):
        """pack -> one all-gather -> unpack (the max word travels with the planes)."""
        import ctypes as C
        from . import cabi
        H, W = x.shape[2], x.shape[3]
        slot = int(run.lib.sdpc_shard_slot_floats(self.per, H, W))
        if self._gbuf is None or self._gbuf.numel() != slot * self.world or self._gbuf.device != x.device:
            self._gbuf = torch.empty(slot * self.world, dtype=torch.float32, device=x.device)
        ptr = lambda t: C.c_void_p(t.data_ptr())
        mine = self._gbuf[self.rank * slot:(self.rank + 1) * slot]
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        cabi.check(run.lib, run.lib.sdpc_shard_pack(ptr(run.workspace), ptr(x[self.lo:self.hi]), ptr(mine), self.per, H, W, stream),
                   "sdpc_shard_pack")
        if dist.get_backend(self.pg) == "nccl":
            dist.all_gather_into_tensor(self._gbuf, mine, group=self.pg)
        else:
            # gloo moves CUDA tensors only through broadcast / all-reduce: gather = SUM of slots that are zero everywhere
            # but at the owner (x + 0 is exact; test-only path, tests/test_gpu_dist.py::test_sharded_two_ranks_on_one_gpu)
            for r in range(self.world):
                if r != self.rank:
                    self._gbuf[r * slot:(r + 1) * slot].zero_()
            dist.all_reduce(self._gbuf, op=dist.ReduceOp.SUM, group=self.pg)
        cabi.check(run.lib, run.lib.sdpc_shard_unpack(ptr(run.workspace), ptr(x), ptr(self._gbuf), self.world, self.rank,
                                                      self.per, H, W, stream), "sdpc_shard_unpack")

    def _all_gather(self, out, mine):
        if out.is_cuda and dist.get_backend(self.pg) == "nccl":
            dist.all_gather_into_tensor(out, mine, group=self.pg)        # in place: `mine` is out's own block
        elif out.is_cuda:                                                # gloo with CUDA tensors (one-GPU test): see _exchange_fused
            mine = mine.clone()
            out.zero_()
            out[self.rank * self.per:(self.rank + 1) * self.per] = mine
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.pg)
        else:                                                            # gloo (CPU tests)
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine.contiguous(), group=self.pg)
            for r, t in enumerate(parts):
                out[r * self.per:(r + 1) * self.per] = t

    def gather_result(self, x):
        out = x.clone()
        self._all_gather(out, x[self.lo:self.hi].contiguous())
        return out

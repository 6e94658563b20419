"""View sharding across GPUs (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

SURVEY.md 8(e): the score network is per-view (InstanceNorm statistics are per sample), so views are
independent for >99% of the FLOPs.  The cross-view block needs (1) every view of the same group and
(2) one scalar: the tooHigh gate is a max over ALL views of the call (KITTISampling.py:162).

Rank r owns the contiguous block of views [r*B/n, (r+1)*B/n).  Per step:
    update own views                                  (sdpc_langevin_update, tgt range = own block)
    all-reduce(MAX) of max|x0|                        (1 float)
    all-gather of the updated x planes, in place      (512 KiB per view; skipped when every group
                                                       lives entirely on one rank)
    z-buffers + correction for own target views       (sdpc_crossview_share, tgt range = own block)
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
        mx = run.local_max()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.pg)
        run.merge_max(mx)
        if self.needs_gather:
            self._all_gather(x, x[self.lo:self.hi])
        run.share_only(p, b)

    def _all_gather(self, out, mine):
        if out.is_cuda:
            dist.all_gather_into_tensor(out, mine, group=self.pg)        # in place: `mine` is out's own block
        else:                                                            # gloo (CPU tests)
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine.contiguous(), group=self.pg)
            for r, t in enumerate(parts):
                out[r * self.per:(r + 1) * self.per] = t

    def gather_result(self, x):
        out = x.clone()
        self._all_gather(out, x[self.lo:self.hi].contiguous())
        return out

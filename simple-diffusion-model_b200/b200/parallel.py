"""Batch-sharded data parallelism for training (one process per GPU, torch.distributed / NCCL over NVLink).

The reference has no distributed code at all; the only exchange step data parallelism needs is a sum-all-reduce of the
gradients.  Gradients already live in one flat fp32 buffer in the order in which backward completes them
(GradLayout), so the engine reports `[lo, hi)` ranges as they become final and this class launches bucketed
asynchronous all-reduces on them (NCCL runs on its own stream, overlapping the rest of backward).  Parameters that
never get a gradient (y_shift, attention norm) are not in the buffer, so nothing is ever waited on in vain.
Sampling shards by image and needs no collective at all (`shard_range`).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """[start, end) of `total` independent units owned by `rank` (contiguous, sizes differ by at most one)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class DataParallel:
    def __init__(self, net, process_group=None, bucket_bytes=64 << 20, device=None):
        import os
        if os.environ.get("SDM_B200_BUCKET_MB"):
            bucket_bytes = int(os.environ["SDM_B200_BUCKET_MB"]) << 20
        self.net = net
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.pending = []           # async work handles
        self.open = None            # [lo, hi) being coalesced
        self.launched = []          # ranges already handed to the collective (for tests / accounting)
        eng = net.engine()
        device = device if device is not None else next(net.parameters()).device
        self.layout = eng.grad_layout(device)
        flat = self.layout.flatten_params()
        if self.world > 1:
            dist.broadcast(flat, src=0, group=self.group)       # identical replicas (reference: single process)
            for n_, p in net.named_parameters():                # parameters outside the flat buffer (never trained)
                if id(p) not in self.layout.offsets:
                    dist.broadcast(p.data, src=0, group=self.group)
        # SDM_B200_DP_RESERVE_SMS=<r>: under data parallelism the persistent tensor-core grids leave r SMs to NCCL's CTAs
        # (0 = off, the default; see DESIGN.md section 6 for the measurement)
        reserve = int(os.environ.get("SDM_B200_DP_RESERVE_SMS", "0"))
        if self.world > 1 and reserve > 0:
            import b200
            sms = torch.cuda.get_device_properties(device).multi_processor_count
            b200.set_option("sm_limit", max(1, sms - reserve))
        eng.on_grads_ready = self.ready
        eng.post_backward = self.finish
        self.opt = None              # FusedAdam attached with attach_optimizer(): bucket-wise updates under the backward pass
        self.opt_stream = None
        self._stepping = False

    def attach_optimizer(self, optimizer):
        """Each bucket is updated (Adam, memory-bound) on a second stream as soon as its all-reduce has landed, underneath the
        tensor-core kernels of the remaining backward pass, instead of one 17 GB pass after it.  `optimizer.step()` still has
        to be called after backward: it finishes the ranges no bucket covered and does the bookkeeping."""
        self.opt = optimizer
        return self

    @property
    def grad_scale(self):
        """Multiply summed gradients by this (1 / world) -- FusedAdam folds it into its kernel."""
        return 1.0 / self.world

    def _launch(self, lay, lo, hi):
        self.launched.append((lo, hi))
        work = None
        if self.world > 1:
            work = dist.all_reduce(lay.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        if self.opt is None or not hasattr(self.opt, "step_range"):
            if work is not None:
                self.pending.append(work)
            return
        dev = lay.flat.device
        main = torch.cuda.current_stream(dev)
        if self.opt_stream is None:
            self.opt_stream = torch.cuda.Stream(device=dev)
        if not self._stepping:
            self._stepping = self.opt.begin_step(lay)      # on the main stream: ordered before every range update
            if not self._stepping:
                if work is not None:
                    self.pending.append(work)
                return
        self.opt_stream.wait_stream(main)                  # the bucket's gradients are final on the main stream
        with torch.cuda.stream(self.opt_stream):
            if work is not None:
                work.wait()                                # ... and summed over the ranks
            self.opt.step_range(lay, lo, hi)

    def ready(self, lay, lo, hi):
        if self.open is not None and self.open[1] == lo:
            self.open[1] = hi                                   # adjacent in completion order: coalesce
        else:
            if self.open is not None:
                self._launch(lay, *self.open)
            self.open = [lo, hi]
        if self.open[1] - self.open[0] >= self.bucket_elems:
            self._launch(lay, *self.open)
            self.open = None

    def finish(self, lay):
        if self.open is not None:
            self._launch(lay, *self.open)
            self.open = None
        for w in self.pending:
            w.wait()
        self.pending = []
        if self._stepping:
            torch.cuda.current_stream(lay.flat.device).wait_stream(self.opt_stream)
            self._stepping = False

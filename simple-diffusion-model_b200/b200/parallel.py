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
    def __init__(self, net, process_group=None, bucket_bytes=64 << 20, device=None, grad_dtype=None, _force_transport=False):
        """grad_dtype: "fp32" (default) or "bf16" (also SDM_B200_DP_GRAD_DTYPE): a bucket is cast to bf16, summed by NCCL in
        bf16 and consumed by the fused Adam as bf16 -- half the bytes on the wire; moments, weights and the optimiser
        arithmetic stay fp32.  `_force_transport` exercises the cast path at world size 1 (tests)."""
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
        self.grad_dtype = grad_dtype or os.environ.get("SDM_B200_DP_GRAD_DTYPE", "fp32")
        if self.grad_dtype not in ("fp32", "bf16"):
            raise ValueError("grad_dtype must be 'fp32' or 'bf16'")
        self.g16 = None              # bf16 transport buffer, same offsets as the flat gradient buffer
        if self.grad_dtype == "bf16" and (self.world > 1 or _force_transport):
            self.g16 = torch.zeros(self.layout.total, dtype=torch.bfloat16, device=device)
        self._cast_back = []         # buckets to expand to fp32 when no fused optimiser consumes the bf16 sums
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
        from ._lib import call, ptr, stream
        self.launched.append((lo, hi))
        work = None
        g16 = None
        if self.g16 is not None:
            g16 = self.g16[lo:hi]
            call("b2_cast_f32_bf16", ptr(lay.flat[lo:hi]), ptr(g16), hi - lo, 0, stream())
        if self.world > 1:
            work = dist.all_reduce(g16 if g16 is not None else lay.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        if self.opt is None or not hasattr(self.opt, "step_range"):
            if work is not None:
                self.pending.append(work)
            if g16 is not None:
                self._cast_back.append((lo, hi))
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
            self.opt.step_range(lay, lo, hi, grad16=g16)

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
        if self._cast_back:                                     # no fused optimiser: hand the bf16 sums back as fp32 gradients
            from ._lib import call, ptr, stream
            for lo, hi in self._cast_back:
                call("b2_cast_f32_bf16", ptr(lay.flat[lo:hi]), ptr(self.g16[lo:hi]), hi - lo, 1, stream())
            self._cast_back = []
        if self._stepping:
            torch.cuda.current_stream(lay.flat.device).wait_stream(self.opt_stream)
            self._stepping = False

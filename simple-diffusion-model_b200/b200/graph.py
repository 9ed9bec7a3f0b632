"""CUDA-graph execution of the hot loops.

The reference's train step issues ~4 k ATen launches from Python (train_diffusion.py:333-364) and each sampler step
~1.2 k (diffusion_sampling_algorithms.py:34-55); at small per-GPU batches both are bound by the host.  Every kernel of
this library is allocation-free, sync-free and takes its tensors by pointer, so a whole optimisation step
(q-sample -> U-Net forward -> MSE + gradient -> backward -> gradient all-reduce -> Adam) and a whole U-Net evaluation
can be captured once and replayed with static input buffers.  Step-dependent scalars (Adam bias corrections, learning
rate) live in device memory (b2_adam_flat_graph), so the captured kernels never change.
"""
import torch

from . import _lib
from ._lib import B200Error, call, ptr, stream


PDL_MAX_PIXELS = 65536      # N * H * W up to which programmatic dependent launch pays (measured: 32 Ki pixels +4 %, 512 Ki -1 %)


def _auto_pdl(n, h, w):
    """Programmatic dependent launch for the graph about to be captured: on for small workloads (launch gaps dominate), off for
    large ones, unless SDM_B200_PDL forces it.  The attribute is baked into the captured launches."""
    import os
    if os.environ.get("SDM_B200_PDL") in ("0", "1"):
        return
    import b200
    b200.set_option("pdl", 1 if n * h * w <= PDL_MAX_PIXELS else 0)


class GraphedTrainStep:
    """One trainer step as a replayable graph.  kind: "eps" (train_diffusion.py:336-350, train_doodle_diffusion.py:304-315;
    target = eps), "x0" (train_noise_cold_diffusion.py:330-340; target = x0) or "target" (train_SR_diffusion.py:366-372;
    explicit target tensor).  The optimiser must be a FusedAdam(capturable=True) over `net.parameters()`."""

    def __init__(self, net, degrader, optimizer, kind="eps", warmup=2, philox_seed=None, philox_first_elem=0, cond_t=None):
        """philox_seed: draw eps INSIDE the q-sample kernel (and re-draw it inside the MSE kernel for kind "eps") instead of
        taking an `eps` tensor: call the step with eps=None.  The draw number is the fused optimiser's device-side step count,
        so every replay of the graph sees fresh noise without an RNG launch.  philox_first_elem: global index of this rank's
        first image element (rank * N * C * H * W).  cond_t: kind "target" in Philox mode -- `cond_img` is then the CLEAN
        low-resolution image, noised here at cond_t with the SAME eps (train_SR_diffusion.py:358-366)."""
        if kind not in ("eps", "x0", "target"):
            raise ValueError("kind must be 'eps', 'x0' or 'target'")
        if not getattr(optimizer, "capturable", False):
            raise B200Error("GraphedTrainStep needs FusedAdam(..., capturable=True)")
        self.net, self.degrader, self.opt, self.kind, self.warmup = net, degrader, optimizer, kind, warmup
        self.philox_seed, self.philox_first_elem, self.cond_t = philox_seed, int(philox_first_elem), cond_t
        self.graph = None
        self.key = None
        self.replays = 0
        self.launches_per_step = 0

    # the step body, written against the engine directly (no autograd graph is built)
    def _body(self):
        s = self.static
        noise = None
        cond_img = s["cond_img"]
        if self.philox_seed is not None:
            from degraders import PhiloxNoise
            eng0 = self.net.engine()
            lay = eng0.grad_layout(s["x0"].device)
            dev_state = self.opt.device_state(lay, self.opt._flat_group(lay) or self.opt.param_groups[0])
            noise = PhiloxNoise(self.philox_seed, 0, dev_state[0:1], self.philox_first_elem)
            x_t = self.degrader.forward_philox(s["x0"], s["t"], noise)
            if self.cond_t is not None and cond_img is not None:
                cond_img = self.degrader.forward_philox(cond_img, s["cond_t"], noise)
        else:
            x_t = self.degrader(s["x0"], s["t"], s["eps"])
        inp = torch.cat((x_t, cond_img), dim=1) if cond_img is not None else x_t
        eng = self.net.engine()
        with torch.no_grad():
            pred, tape = eng._forward_tape(inp, s["t"], s["labels"])
            if noise is not None and self.kind == "eps":
                call("b2_mse_loss_grad_philox", ptr(pred), ptr(s["dpred"]), ptr(s["loss"]), pred.numel(), 1.0, noise.seed,
                     noise.offset, ptr(noise.offset_dev), noise.first_elem, stream())
            else:
                target = {"eps": s["eps"], "x0": s["x0"], "target": s["target"]}[self.kind]
                call("b2_mse_loss_grad", ptr(pred), ptr(target), ptr(s["dpred"]), ptr(s["loss"]), pred.numel(), 1.0, stream())
            eng._backward_tape(tape, s["dpred"])
            self.opt.step()

    def _capture(self, x0, t, eps, labels, cond_img, target):
        dev = x0.device
        clone = lambda v: None if v is None else v.detach().clone().contiguous()
        self.static = {"x0": clone(x0.float()), "t": clone(t.to(torch.int64)), "eps": clone(eps.float()) if eps is not None else None,
                       "cond_t": torch.tensor([int(self.cond_t)], dtype=torch.int64, device=dev) if self.cond_t is not None else None,
                       "labels": clone(labels),
                       "cond_img": clone(cond_img), "target": clone(target), "loss": torch.zeros((), dtype=torch.float32, device=dev)}
        n, _, h, w = x0.shape
        _auto_pdl(n, h, w)
        # (the 128-channel conv variants are auto-tuned for inference graphs only -- GraphedUNet._capture: in the train step the
        # swapped form won the micro-benchmark but not the step: 83.3 vs 82.9 ms, profiles/r02z13_autotune_ab.log)
        from .autotune import reset_conv128
        reset_conv128()
        import os
        if os.environ.get("SDM_B200_OVERLAP_WGRAD") not in ("0", "1"):
            # small workloads are bound by the length of ~1.3 k short dependent kernels: weight gradients (off the critical path)
            # go to a second stream (measured at 64x64 batch 8 together with PDL: -5.6 %; neutral or worse at 128x128 batch 32)
            self.net.engine().overlap_wgrad = n * h * w <= PDL_MAX_PIXELS
        # Deferring a module's weight gradients into ONE grouped launch (b2_conv2d_wgrad_batch, SDM_B200_GROUP_WGRAD=1) was measured
        # at 64x64 batch 8 (profiles/r02z6_grouped_wgrad_ab.log): 17.97 ms grouped vs 18.76 ms per layer on one stream, but 18.26 ms
        # vs 17.50 ms when the weight gradients run on the side stream -- fine-grained interleaving beats fewer launches -- so it
        # stays opt-in.
        out_ch = self.net.out_layers[1].conv_layer[0].weight.shape[0]
        self.static["dpred"] = torch.empty((n, out_ch, h, w), dtype=torch.float32, device=dev)
        lay = self.net.engine().grad_layout(dev)
        lay.flatten_params()
        # The warm-up iterations (allocator, NCCL communicators, weight cache) run the real step; parameters, moments
        # and step counters are restored afterwards so that capturing has no training side effect.
        opt = self.opt
        had_moments = id(lay) in opt._flat
        saved = [lay.params_flat.clone()] + ([m.clone() for m in opt._flat[id(lay)]] if had_moments else [])
        steps_before = 0.0
        for group in opt.param_groups:
            for p in group["params"]:
                st = opt.state.get(p)
                if st and "step" in st:
                    steps_before = float(st["step"])
                    break
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        l0 = _lib.LAUNCHES
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self._body()
        self.launches_per_step = _lib.LAUNCHES - l0      # C-ABI calls recorded in the graph = launched by every replay
        lay.params_flat.copy_(saved[0])
        m_flat, v_flat = opt._flat[id(lay)]
        if had_moments:
            m_flat.copy_(saved[1])
            v_flat.copy_(saved[2])
        else:
            m_flat.zero_()
            v_flat.zero_()
        for group in opt.param_groups:
            dev_state = opt.device_state(lay, group)
            dev_state[0:1].fill_(steps_before)
            for p in group["params"]:
                st = opt.state.get(p)
                if st and "step" in st:
                    st["step"].fill_(steps_before)
        lay.epoch += 1
        if lay.shadow is not None:              # the bf16 copy must describe the restored weights
            lay.shadow.copy_(lay.params_flat)
            lay.stepped(True)
        del saved

    def __call__(self, x0, t, eps=None, labels=None, cond_img=None, target=None):
        if not x0.is_cuda:
            raise B200Error("training needs CUDA tensors: this build has no CPU path")
        if (eps is None) != (self.philox_seed is not None):
            raise B200Error("GraphedTrainStep: pass eps=None exactly when the step was built with philox_seed")
        key = (tuple(x0.shape), tuple(t.shape), None if labels is None else tuple(labels.shape),
               None if cond_img is None else tuple(cond_img.shape), None if target is None else tuple(target.shape),
               self.net.precision)
        if self.graph is None:
            self.key = key
            self._capture(x0, t, eps, labels, cond_img, target)
        elif key != self.key:
            # a differently shaped batch (the short last batch of an epoch): run the same kernel sequence eagerly
            captured, self.static = self.static, {
                "x0": x0.contiguous().float(), "t": t.to(torch.int64), "eps": eps.contiguous().float() if eps is not None else None,
                "cond_t": captured.get("cond_t"), "labels": labels,
                "cond_img": cond_img, "target": target, "loss": torch.zeros((), dtype=torch.float32, device=x0.device),
                "dpred": torch.empty((x0.shape[0], captured["dpred"].shape[1]) + tuple(x0.shape[2:]), dtype=torch.float32,
                                     device=x0.device)}
            try:
                self.opt.sync_lr()
                self._body()
                return self.static["loss"]
            finally:
                self.static = captured
        s = self.static
        s["x0"].copy_(x0, non_blocking=True)
        s["t"].copy_(t, non_blocking=True)
        if eps is not None:
            s["eps"].copy_(eps, non_blocking=True)
        for name, v in (("labels", labels), ("cond_img", cond_img), ("target", target)):
            if v is not None:
                s[name].copy_(v, non_blocking=True)
        self.opt.sync_lr()
        self.graph.replay()
        self.replays += 1
        _lib.LAUNCHES += self.launches_per_step
        self.opt.note_replayed()
        lay = self.net.engine().layout
        lay.stepped(lay.shadow is not None)      # eager users of the weight cache (eval / sampling) must re-pack
        return s["loss"]


class GraphedUNet:
    """Inference-mode U_Net evaluation as a replayable graph with the call signature the samplers use
    (`diffusion_net(x, t, labels)`, diffusion_sampling_algorithms.py:34-37).  One graph per input signature."""

    def __init__(self, net, warmup=1):
        self.net, self.warmup = net, warmup
        self.graphs = {}

    def eval(self):
        self.net.eval()
        return self

    def train(self, mode=True):
        self.net.train(mode)
        return self

    def __getattr__(self, name):
        return getattr(self.__dict__["net"], name)

    def _weights_key(self):
        lay = getattr(self.net.engine(), "layout", None)
        return (self.net.precision, lay.epoch if lay is not None else 0,
                tuple(p._version for p in self.net.parameters()))

    def __call__(self, x, t=None, cond=None):
        if not x.is_cuda:
            raise B200Error("U_Net.forward needs CUDA tensors: this build has no CPU path")
        key = (tuple(x.shape), None if t is None else tuple(t.shape), None if cond is None else tuple(cond.shape))
        wkey = self._weights_key()
        entry = self.graphs.get(key)
        if entry is None or entry["wkey"] != wkey:
            entry = self._capture(x, t, cond, wkey)
            self.graphs[key] = entry
        entry["x"].copy_(x, non_blocking=True)
        if t is not None:
            entry["t"].copy_(t, non_blocking=True)
        if cond is not None:
            entry["cond"].copy_(cond, non_blocking=True)
        entry["graph"].replay()
        _lib.LAUNCHES += entry["launches"]
        return entry["out"]

    def _capture(self, x, t, cond, wkey):
        dev = x.device
        eng = self.net.engine()
        _auto_pdl(x.shape[0], x.shape[2], x.shape[3])
        from .autotune import tune_conv128
        tune_conv128(self.net, x.shape[0], x.shape[2], x.shape[3], dev)
        st = {"x": x.detach().clone().contiguous().float(), "t": None if t is None else t.detach().clone().to(torch.int64),
              "cond": None if cond is None else cond.detach().clone().float(), "wkey": wkey}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):         # packs the kernel-layout weights outside the graph
                eng.forward(st["x"], st["t"], st["cond"])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        l0 = _lib.LAUNCHES
        with torch.cuda.graph(g, capture_error_mode="thread_local"), torch.no_grad():
            st["out"] = eng.forward(st["x"], st["t"], st["cond"])
        st["launches"] = _lib.LAUNCHES - l0
        st["graph"] = g
        return st

"""FusedAdam: torch.optim.Adam semantics (the reference's optimiser, train_diffusion.py:214-218) executed as ONE
sm_100a kernel over the flat parameter / gradient / moment buffers of a U_Net (b200.train_engine.GradLayout), or one
launch per tensor for parameters that are not part of a flat layout.  `state_dict()` keeps torch's per-parameter
keys ("step", "exp_avg", "exp_avg_sq": views of the flat moments) so optimiser checkpoints interchange."""
import math

import torch

from ._lib import call, ptr, stream


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, capturable=False):
        # the extra keys are torch.optim.Adam's remaining per-group options at their defaults, so that a state_dict written
        # here loads into the reference's torch.optim.Adam (train_diffusion.py:214-227) and vice versa
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self.grad_scale = grad_scale          # e.g. 1 / world_size when gradients were sum-all-reduced
        self._flat = {}                       # id(layout) -> (m_flat, v_flat)
        # capturable: step count / lr / grad scale live in device memory (b2_adam_flat_graph) so that a CUDA graph
        # holding step() replays correctly; b200.graph.GraphedTrainStep keeps the host-side counters in sync.
        self.capturable = capturable
        self.bf16_shadow = True               # emit the bf16 copy of the weights the tensor-core kernels read (flat layouts)
        self._dev_state = {}                  # id(layout) -> device float[8]
        self._dev_lr = {}
        self._adopted = set()                 # id(param) whose state moments are views of the flat buffers

    def device_state(self, lay, group, steps=0.0):
        st = self._dev_state.get(id(lay))
        if st is None:
            st = torch.tensor([steps, group["lr"], self.grad_scale, 0, 0, 0, 0, 0], dtype=torch.float32, device=lay.flat.device)
            self._dev_state[id(lay)] = st
            self._dev_lr[id(lay)] = group["lr"]
        return st

    def sync_lr(self):
        """Pushes a changed learning rate (the reference halves it every lr_steps, train_diffusion.py:368-371) to the device."""
        for group in self.param_groups:
            for p in group["params"]:
                lay = getattr(p, "_b2_layout", None)
                if lay is not None and id(lay) in self._dev_state and self._dev_lr[id(lay)] != group["lr"]:
                    self._dev_state[id(lay)][1:2].fill_(group["lr"])
                    self._dev_lr[id(lay)] = group["lr"]
                break

    def note_replayed(self):
        """A captured step() ran on the device: advance the host-side per-parameter step counters to match."""
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state.get(p)
                if st and "step" in st:
                    st["step"] += 1

    def _moments(self, lay):
        if id(lay) not in self._flat:
            self._flat[id(lay)] = (torch.zeros_like(lay.flat), torch.zeros_like(lay.flat))
        return self._flat[id(lay)]

    def _adopt_state(self, lay, p, st):
        """Makes state[p]'s moments views of the flat moment buffers the kernels update.  Fresh state starts at zero; moments
        that arrived as standalone tensors (load_state_dict of a reference / earlier checkpoint) are copied in first."""
        if id(p) in self._adopted and st:
            return
        self._adopted.add(id(p))
        m_flat, v_flat = self._moments(lay)
        m_view, v_view = lay._shaped(m_flat, p), lay._shaped(v_flat, p)
        if not st:
            st["step"] = torch.tensor(0.0)
        for key, view in (("exp_avg", m_view), ("exp_avg_sq", v_view)):
            have = st.get(key)
            if have is not None and have.data_ptr() != view.data_ptr():
                view.copy_(have.to(view.device, torch.float32))
            st[key] = view

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        """torch's loader installs standalone `exp_avg` / `exp_avg_sq` tensors; the kernels only see the flat buffers, so
        the loaded moments are copied into them (and the device-side step count / learning rate re-seeded).  Resuming
        with `load_diffusion_optim` (train_diffusion.py:219-227) therefore continues the reference's Adam trajectory."""
        super().load_state_dict(state_dict)
        self._adopted.clear()
        for group in self.param_groups:
            group["capturable"] = False          # a foreign checkpoint must not move `step` handling to torch's capturable path
            seen = {}
            for p in group["params"]:
                lay = getattr(p, "_b2_layout", None)
                st = self.state.get(p)
                if lay is None or not st or lay.params_flat is None or id(p) not in lay.offsets:
                    continue
                if not torch.is_tensor(st.get("step")):
                    st["step"] = torch.tensor(float(st.get("step", 0.0)))
                st["step"] = st["step"].detach().to("cpu", torch.float32).reshape(())
                self._adopt_state(lay, p, st)
                seen[id(lay)] = (lay, float(st["step"]))
            for lay, steps in seen.values():
                if self.capturable:
                    dev = self.device_state(lay, group, steps)
                    dev[0:1].fill_(steps)
                    dev[1:2].fill_(group["lr"])
                    dev[2:3].fill_(self.grad_scale)
                    self._dev_lr[id(lay)] = group["lr"]

    @staticmethod
    def _launch(p, g, m, v, n, group, step, grad_scale, shadow=None):
        b1, b2 = group["betas"]
        step_size = group["lr"] / (1.0 - b1 ** step)
        inv_bc2_sqrt = 1.0 / math.sqrt(1.0 - b2 ** step)
        call("b2_adam_flat", ptr(p), ptr(g), ptr(m), ptr(v), n, b1, b2, group["eps"], step_size, inv_bc2_sqrt, grad_scale,
             ptr(shadow), stream())

    # ---- bucket-wise stepping (b200/parallel.py): begin_step -> step_range* -> step() finishes whatever is left ----------
    def _flat_group(self, lay):
        for group in self.param_groups:
            for p in group["params"]:
                if getattr(p, "_b2_layout", None) is lay:
                    return group
        return None

    @torch.no_grad()
    def begin_step(self, lay):
        """Fixes this step's bias corrections (device-side advance in capturable mode) so that ranges of the flat buffers
        can be updated one by one, in any order, possibly on another stream."""
        group = self._flat_group(lay)
        if group is None or lay.params_flat is None:
            return False
        self._moments(lay)
        any_state = next((self.state[p] for p in group["params"] if self.state.get(p)), None)
        steps_so_far = float(any_state["step"]) if any_state else 0.0
        b1, b2 = group["betas"]
        if self.capturable:
            dev = self.device_state(lay, group, steps_so_far)
            call("b2_adam_advance", ptr(dev), float(b1), float(b2), stream())
        self._partial = {"lay": lay, "group": group, "done": [], "step": int(steps_so_far) + 1,
                         "shadow": lay.ensure_shadow() if self.bf16_shadow else None}
        return True

    @torch.no_grad()
    def step_range(self, lay, lo, hi, grad16=None):
        """grad16: the range's gradients as bf16 (data-parallel transport buffer) instead of lay.flat[lo:hi]."""
        part = getattr(self, "_partial", None)
        if part is None or part["lay"] is not lay or hi <= lo:
            return
        group, shadow = part["group"], part["shadow"]
        m_flat, v_flat = self._flat[id(lay)]
        b1, b2 = group["betas"]
        sh = shadow[lo:hi] if shadow is not None else None
        if grad16 is not None:
            if self.capturable:
                call("b2_adam_flat_g16", ptr(lay.params_flat[lo:hi]), ptr(grad16), ptr(m_flat[lo:hi]), ptr(v_flat[lo:hi]), hi - lo,
                     float(b1), float(b2), group["eps"], 0.0, 0.0, 0.0, ptr(self._dev_state[id(lay)]), ptr(sh), stream())
            else:
                step = part["step"]
                call("b2_adam_flat_g16", ptr(lay.params_flat[lo:hi]), ptr(grad16), ptr(m_flat[lo:hi]), ptr(v_flat[lo:hi]), hi - lo,
                     float(b1), float(b2), group["eps"], group["lr"] / (1.0 - b1 ** step), 1.0 / math.sqrt(1.0 - b2 ** step),
                     self.grad_scale, None, ptr(sh), stream())
            part["done"].append((lo, hi))
            return
        if self.capturable:
            call("b2_adam_flat_graph", ptr(lay.params_flat[lo:hi]), ptr(lay.flat[lo:hi]), ptr(m_flat[lo:hi]), ptr(v_flat[lo:hi]),
                 hi - lo, float(b1), float(b2), group["eps"], ptr(self._dev_state[id(lay)]), ptr(sh), 0, stream())
        else:
            self._launch(lay.params_flat[lo:hi], lay.flat[lo:hi], m_flat[lo:hi], v_flat[lo:hi], hi - lo, group, part["step"],
                         self.grad_scale, sh)
        part["done"].append((lo, hi))

    def _finish_partial(self):
        """Updates the ranges no bucket covered, then does the per-step bookkeeping of step()."""
        part = self._partial
        lay, group = part["lay"], part["group"]
        pos = 0
        for lo, hi in sorted(part["done"]):
            if lo > pos:
                self.step_range(lay, pos, lo)
            pos = max(pos, hi)
        if pos < lay.total:
            self.step_range(lay, pos, lay.total)
        m_flat, v_flat = self._flat[id(lay)]
        for p in group["params"]:
            if getattr(p, "_b2_layout", None) is not lay or id(p) not in lay.offsets:
                continue
            st = self.state[p]
            self._adopt_state(lay, p, st)
            st["step"] += 1
        lay.stepped(part["shadow"] is not None)
        self._partial = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if getattr(self, "_partial", None) is not None:
            self._finish_partial()
            return loss
        for group in self.param_groups:
            done_layouts = set()
            for p in group["params"]:
                if p.grad is None:
                    continue
                lay = getattr(p, "_b2_layout", None)
                flat_ok = lay is not None and p.grad.data_ptr() == lay.view(p).data_ptr() and lay.params_flat is not None \
                    and p.data_ptr() == lay.param_view(p).data_ptr()
                st = self.state[p]
                if flat_ok:
                    m_flat, v_flat = self._moments(lay)
                    if id(lay) not in done_layouts:
                        # every parameter of the layout is adopted BEFORE the single flat launch below (same, possibly
                        # channels-last, view as the parameter; standalone moments loaded from a checkpoint are copied in)
                        for q in group["params"]:
                            if getattr(q, "_b2_layout", None) is lay and id(q) in lay.offsets and q.grad is not None:
                                self._adopt_state(lay, q, self.state[q])
                    if self.capturable:
                        self.device_state(lay, group, float(st["step"]))      # created once, from the pre-step count
                    st["step"] += 1
                    if id(lay) not in done_layouts:
                        done_layouts.add(id(lay))
                        shadow = lay.ensure_shadow() if self.bf16_shadow else None
                        if self.capturable:
                            b1, b2 = group["betas"]
                            dev = self.device_state(lay, group)
                            call("b2_adam_flat_graph", ptr(lay.params_flat), ptr(lay.flat), ptr(m_flat), ptr(v_flat), lay.total,
                                 float(b1), float(b2), group["eps"], ptr(dev), ptr(shadow), 1, stream())
                        else:
                            self._launch(lay.params_flat, lay.flat, m_flat, v_flat, lay.total, group, int(st["step"]),
                                         self.grad_scale, shadow)
                        lay.stepped(shadow is not None)          # cached kernel-layout weights are stale now
                    continue
                if not st:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                g = p.grad.contiguous()
                if (p.data_ptr() | g.data_ptr()) % 16 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam needs 16-byte aligned contiguous fp32 parameters")
                self._launch(p, g, st["exp_avg"], st["exp_avg_sq"], p.numel(), group, int(st["step"]), self.grad_scale)
                if lay is not None:
                    lay.epoch += 1
        return loss

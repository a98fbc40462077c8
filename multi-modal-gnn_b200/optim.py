"""torch.optim.Adam as the reference builds it (train.py:255-260: lr, weight_decay; betas (0.9, 0.999), eps 1e-8, L2
weight decay, no amsgrad) with the whole step in ONE libb2g launch (csrc/optim.cu).

A regular ``torch.optim.Optimizer``: ``ReduceLROnPlateau`` / ``StepLR`` drive it through ``param_groups`` and
``state_dict()`` has torch.optim.Adam's layout (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``), so checkpoints move
between the two.  Parameters whose gradient is None are skipped -- no state, no step -- exactly like torch.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import _lib

_DESC = np.dtype([("param", "<u8"), ("grad", "<u8"), ("exp_avg", "<u8"), ("exp_avg_sq", "<u8"), ("numel", "<i8")])   # b2g_adam_tensor_t


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameter")
        # torch.optim.Adam's own defaults (so a state_dict saved here loads there), with our values on top
        defaults = dict(torch.optim.Adam([torch.zeros(1)], lr=lr, betas=betas, eps=eps, weight_decay=weight_decay).defaults)
        super().__init__(params, defaults)
        self._tables: Dict[tuple, tuple] = {}
        self.launches = 0

    def _table(self, ps: List[torch.Tensor]):
        """Device copies of the descriptor / chunk tables for this exact set of tensors (cached: in CUDA-graph mode the
        gradients keep their addresses, so the tables are built once)."""
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
        hit = self._tables.get(key)
        if hit is not None:
            return hit
        lib = _lib.load()
        chunk = int(lib.b2g_adam_chunk_elems())
        desc = np.zeros(len(ps), dtype=_DESC)
        chunks = []
        for i, p in enumerate(ps):
            st = self.state[p]
            desc[i] = (p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel())
            chunks.extend((i, off) for off in range(0, p.numel(), chunk))
        dev = ps[0].device
        d_desc = torch.from_numpy(desc.view(np.uint8).copy()).to(dev)
        d_chunks = torch.tensor(chunks, dtype=torch.int32).reshape(-1, 2).contiguous().to(dev)
        if len(self._tables) > 8:
            self._tables.clear()
        self._tables[key] = (d_desc, d_chunks, len(chunks))
        return self._tables[key]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            cohorts: Dict[int, List[torch.Tensor]] = {}        # parameters sharing one step counter are updated together
            fresh = None
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise _lib.B2GError("FusedAdam handles dense float32 CUDA parameters only (there is no CPU path)")
                st = self.state[p]
                if len(st) == 0:
                    if fresh is None:
                        fresh = torch.tensor(0.0, dtype=torch.float32)          # like torch: a CPU scalar tensor
                    st["step"] = fresh
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                cohorts.setdefault(id(st["step"]), []).append(p)
            for ps in cohorts.values():
                step_t = self.state[ps[0]]["step"]
                step_t += 1                                    # one shared tensor per cohort
                t = float(step_t)
                bc1 = 1.0 - beta1 ** t
                bc2 = 1.0 - beta2 ** t
                d_desc, d_chunks, n_chunks = self._table(ps)
                _lib.check(lib.b2g_adam_step(d_desc.data_ptr(), len(ps), d_chunks.data_ptr(), n_chunks, group["lr"] / bc1, bc2 ** 0.5,
                                             1.0 - beta1, beta2, 1.0 - beta2, group["eps"], group["weight_decay"],
                                             torch.cuda.current_stream().cuda_stream), "b2g_adam_step")
                # the kernel wrote through raw pointers: move the tensors' version counters like an in-place op would (caches keyed
                # on parameter versions, e.g. HeteroRGCN's eval-mode embedding cache, must see the update)
                torch.autograd.graph.increment_version(ps)
                self.launches += 1
        return loss

    def state_dict(self):
        """torch.optim.Adam's layout.  The step counters are shared between parameters internally; every parameter gets
        its own copy here, because torch's foreach Adam increments each listed step tensor once per parameter."""
        sd = super().state_dict()
        for st in sd["state"].values():
            if "step" in st:
                st["step"] = st["step"].clone()
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables.clear()
        # torch stores one step tensor per parameter; re-share them between parameters that are at the same step
        by_value: Dict[float, torch.Tensor] = {}
        for st in self.state.values():
            if "step" in st:
                v = float(st["step"])
                st["step"] = by_value.setdefault(v, torch.tensor(v, dtype=torch.float32))

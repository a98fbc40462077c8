"""Host side of the training step: the hot part of the reference's ``src/train.py`` (EdgeMasker,
Trainer._compute_lab_weights / train_epoch / validate) with the same names, argument meaning and error
behaviour, driving the CUDA model.  The reference's own ``train.py`` can also drive the model unchanged
(INTEGRATION.md); this module exists because the reference is absent on the GPU box and because its
loss/selection tail is fused here into one kernel (``ops.weighted_loss``).

Integer / boolean work (splits, supervision masks) stays on the host with torch's CPU generator so it
is bit-exact with the reference.
"""
from __future__ import annotations

import time
from typing import Dict, Optional, Tuple

import torch

from . import _lib, ops
from .dist import DistContext, allreduce_gradients, gradient_pattern
from .model import compute_regression_loss


class EdgeMasker:
    """train.py:37-176.  70/15/15 edge split by seeded CPU randperm; per-epoch 20 % supervision mask."""

    def __init__(self, data, train_split: float = 0.7, val_split: float = 0.15, test_split: float = 0.15,
                 mask_fraction: float = 0.2, seed: int = 42):
        assert abs(train_split + val_split + test_split - 1.0) < 1e-6, "Splits must sum to 1.0"
        self.data = data
        self.train_split, self.val_split, self.test_split = train_split, val_split, test_split
        self.mask_fraction, self.seed = mask_fraction, seed
        self.edge_type = ("patient", "has_lab", "lab")
        self.edge_index = data[self.edge_type].edge_index
        self.edge_attr = data[self.edge_type].edge_attr
        self.num_edges = int(self.edge_index.shape[1])
        self.train_mask, self.val_mask, self.test_mask = self._create_splits()
        self._split_cache: Dict[str, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = {}

    def _create_splits(self):
        """train.py:98-129 (host, CPU generator: bit-exact with the reference for the same seed)."""
        torch.manual_seed(self.seed)
        perm = torch.randperm(self.num_edges)
        n_train = int(self.train_split * self.num_edges)
        n_val = int(self.val_split * self.num_edges)
        masks = [torch.zeros(self.num_edges, dtype=torch.bool) for _ in range(3)]
        masks[0][perm[:n_train]] = True
        masks[1][perm[n_train:n_train + n_val]] = True
        masks[2][perm[n_train + n_val:]] = True
        return masks

    def split_mask(self, split: str) -> torch.Tensor:
        if split == "train":
            return self.train_mask
        if split == "val":
            return self.val_mask
        if split == "test":
            return self.test_mask
        raise ValueError(f"Unknown split: {split}")

    def supervision_mask(self, split: str, seed: Optional[int] = None) -> torch.Tensor:
        """train.py:150-165 (host).  ``seed=None`` reproduces the reference's wall-clock reseeding."""
        n = int(self.split_mask(split).sum())
        if split == "train" and self.mask_fraction > 0:
            torch.manual_seed(int(time.time()) if seed is None else int(seed))
            return torch.rand(n) < self.mask_fraction
        return torch.ones(n, dtype=torch.bool)

    def split_edges(self, split: str):
        """(edge_indices [2, n], edge_values [n]) of a split, on the graph's device (cached: static)."""
        if split not in self._split_cache:
            mask = self.split_mask(split).to(self.edge_index.device)
            ei = self.edge_index[:, mask].contiguous()
            self._split_cache[split] = (ei, self.edge_attr[mask].squeeze(-1).contiguous(), ei[0], ei[1])
        return self._split_cache[split][:2]

    def split_rows(self, split: str):
        """(patient_indices, lab_indices) views of split_edges(split)[0]; the same tensor objects every call."""
        self.split_edges(split)
        return self._split_cache[split][2:]

    def get_masked_data(self, split: str = "train", seed: Optional[int] = None):
        """train.py:131-176: (edge_indices, edge_values, mask, supervision_mask)."""
        mask = self.split_mask(split)
        sup = self.supervision_mask(split, seed)
        ei, ev = self.split_edges(split)
        return ei, ev, mask, sup


def compute_lab_weights(lab_indices: torch.Tensor, edge_values: torch.Tensor, num_labs: int,
                        dctx: Optional[DistContext] = None) -> torch.Tensor:
    """Trainer._compute_lab_weights (train.py:295-330): 1 / (unbiased variance + 1e-6) per lab over the
    train split (variance := 1 for labs with < 2 samples), rescaled to sum to num_labs.  One-off, O(E)."""
    v = edge_values.double()
    cnt = torch.zeros(num_labs, dtype=torch.float64, device=v.device).index_add_(0, lab_indices, torch.ones_like(v))
    s1 = torch.zeros(num_labs, dtype=torch.float64, device=v.device).index_add_(0, lab_indices, v)
    if dctx is not None:               # patient-partitioned: the variance is over the train pairs of ALL ranks
        dctx.all_reduce_(cnt)
        dctx.all_reduce_(s1)
    mean = s1 / cnt.clamp(min=1)
    s2 = torch.zeros(num_labs, dtype=torch.float64, device=v.device).index_add_(0, lab_indices, (v - mean[lab_indices]) ** 2)
    if dctx is not None:
        dctx.all_reduce_(s2)
    var = torch.where(cnt > 1, s2 / (cnt - 1).clamp(min=1), torch.ones_like(s2))
    w = 1.0 / (var.float() + 1e-6)
    return w * num_labs / w.sum()


class Trainer:
    """train.py:183-431 (optimizer/scheduler construction, lab weights, train_epoch, validate)."""

    def __init__(self, model, data, masker: EdgeMasker, config: Dict, device: torch.device, dist_ctx: Optional[DistContext] = None):
        """``dist_ctx``: patient-partitioned multi-GPU mode -- ``data`` / ``masker`` are this rank's partition (dist.py)."""
        self.dist = dist_ctx
        if dist_ctx is not None:
            model.set_distributed(dist_ctx)
        self.model = model.to(device)
        self.data = data.to(device)
        self.masker = masker
        masker.edge_index = self.data[masker.edge_type].edge_index
        masker.edge_attr = self.data[masker.edge_type].edge_attr
        masker._split_cache.clear()
        self.config, self.device = config, device
        tc = config["train"]
        self.optimizer = self._build_optimizer(tc["optimizer"])
        self.scheduler = self._build_scheduler(tc.get("lr_scheduler", {}))
        self.loss_fn = tc["loss"]
        self.epochs = tc["epochs"]
        self.early_stopping_patience = tc["early_stopping_patience"]
        self.best_val_loss, self.patience_counter = float("inf"), 0
        self.train_losses, self.val_losses = [], []
        self.lab_weights = self._compute_lab_weights()
        # called between backward and optimizer.step: the multi-GPU gradient all-reduce
        # (the sharded node type's embedding rows are rank-local: their gradients are never exchanged)
        self._grad_pattern = None       # which exchanged parameters have a gradient on some rank (static; one host sync)
        self.grad_hook = self._exchange_gradients if self.dist is not None else None
        self._use_graph, self._graph = False, None
        self.graph_kernel_nodes = 0

    def _build_optimizer(self, oc):
        kind = oc.get("type", "adam").lower()
        if kind == "adam":          # train.py:255-260; one-launch multi-tensor step (optim.py) instead of torch's foreach kernels
            from .optim import FusedAdam
            return FusedAdam(self.model.parameters(), lr=oc["lr"], weight_decay=oc["weight_decay"])
        if kind == "sgd":
            return torch.optim.SGD(self.model.parameters(), lr=oc["lr"], weight_decay=oc["weight_decay"],
                                   momentum=oc.get("momentum", 0.9))
        raise ValueError(f"Unknown optimizer: {kind}")

    def _build_scheduler(self, sc):
        if not sc.get("enabled", False):
            return None
        kind = sc.get("type", "reduce_on_plateau")
        if kind == "reduce_on_plateau":
            return torch.optim.lr_scheduler.ReduceLROnPlateau(self.optimizer, mode="min", factor=sc.get("factor", 0.5),
                                                              patience=sc.get("patience", 10))
        if kind == "step":
            return torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=sc.get("step_size", 30), gamma=sc.get("gamma", 0.1))
        raise ValueError(f"Unknown scheduler: {kind}")

    def _compute_lab_weights(self) -> torch.Tensor:
        ei, ev = self.masker.split_edges("train")
        return compute_lab_weights(ei[1], ev, int(self.data["lab"].num_nodes), self.dist)

    def _exchanged_params(self):
        return [p for n, p in self.model.named_parameters() if not n.startswith(f"embeddings.{self.dist.sharded_type}.")]

    def _exchange_gradients(self):
        params = self._exchanged_params()
        if self._grad_pattern is None or len(self._grad_pattern) != len(params):
            self._grad_pattern = gradient_pattern(params, self.dist)
        allreduce_gradients(params, self.dist, self._grad_pattern)

    # ---- CUDA graph mode ------------------------------------------------------------------------------------------
    def enable_cuda_graph(self, enabled: bool = True):
        """Capture forward + loss + backward of the training step into one CUDA graph (the step is ~270 short launches
        and host-bound when issued one by one).  The graph is captured on the first train_step for a given set of
        index / target tensors and replayed afterwards; the supervision mask and the dropout seed are refreshed in
        static device buffers before every replay; the optimizer step stays eager."""
        self._use_graph = bool(enabled)
        self._graph = None
        if not enabled:
            self.model._seed_buffer = None        # eager steps draw a fresh host seed per call again

    def _loss_of(self, pred, edge_values, lab_indices, sup):
        if self.loss_fn in ("mae", "mse"):
            loss = ops.weighted_loss(pred, edge_values, lab_indices, self.lab_weights, sup, self.loss_fn)
        else:
            loss = ops.weighted_loss(pred, edge_values, None, None, sup, self.loss_fn)   # train.py:378-383
        if self.dist is not None:
            # global mean over the supervised pairs of all ranks = sum_r (n_r / n) * local mean_r
            n_local = sup.sum().to(torch.float32)
            cnt = torch.zeros(4, dtype=torch.float32, device=loss.device)      # 16 bytes: eligible for the peer-memory path
            cnt[0] = n_local
            n_total = self.dist.all_reduce_(cnt)[0]
            loss = torch.nan_to_num(loss) * (n_local / n_total)
        return loss

    def global_loss(self, loss: torch.Tensor) -> torch.Tensor:
        """Loss of the whole (partitioned) batch: the per-rank terms of _loss_of summed over ranks."""
        if self.dist is None:
            return loss
        buf = torch.zeros(4, dtype=torch.float32, device=loss.device)
        buf[0] = loss.detach()
        return self.dist.all_reduce_(buf)[0]

    def _drop_unoptimized_grads(self):
        """Graph mode only: parameters the optimizer does not hold (note N2: the lazily created tables, which the reference
        never trains and whose .grad `optimizer.zero_grad()` never resets, so that it accumulates over all steps) start the
        capture without a .grad -- the captured AccumulateGrad then stores the step's gradient instead of adding it to the
        running sum (a 3-pass add over the [N_patient, d] table per step for a value nobody reads).  After a replay their
        .grad is the LAST step's gradient, not the sum over steps; eager mode keeps the reference's behaviour."""
        held = {id(p) for g in self.optimizer.param_groups for p in g["params"]}
        for p in self.model.parameters():
            if id(p) not in held:
                p.grad = None

    def _capture(self, pi, li, ev, sup):
        model, dev = self.model, self.device
        model._seed_buffer = torch.zeros(1, dtype=torch.int64, device=dev)
        sup_static = torch.empty_like(sup)
        sup_static.copy_(sup)
        model._seed_buffer.fill_(int(torch.randint(0, 2 ** 62, (1,)).item()))
        buffers = {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up: allocations, caches, lazy attributes
            for _ in range(2):
                self.optimizer.zero_grad(set_to_none=True)
                self._drop_unoptimized_grads()
                self._loss_of(model.predict_lab_values(self.data, pi, li), ev, li, sup_static).backward()
            if self.grad_hook is not None:                  # learn the (static) gradient pattern outside the capture
                self._grad_pattern = gradient_pattern(self._exchanged_params(), self.dist)
        torch.cuda.current_stream().wait_stream(side)
        with torch.no_grad():                               # the warm-up must not count as training steps
            sd = model.state_dict()
            for k, v in buffers.items():
                sd[k].copy_(v)
        self.optimizer.zero_grad(set_to_none=True)
        self._drop_unoptimized_grads()
        graph = torch.cuda.CUDAGraph()
        lib = _lib.load()
        before = int(lib.b2g_launch_count())
        coll0 = (self.dist.n_collectives, self.dist.n_peer) if self.dist is not None else None
        with torch.cuda.graph(graph):
            loss = self._loss_of(model.predict_lab_values(self.data, pi, li), ev, li, sup_static)
            loss.backward()
            if self.grad_hook is not None:                  # the gradient exchange is part of the captured step
                self.grad_hook()
        self.graph_kernel_nodes = int(lib.b2g_launch_count()) - before      # libb2g kernels replayed per step
        if self.dist is not None:
            self.dist.n_collectives_per_step = self.dist.n_collectives - coll0[0]
            self.dist.n_peer_per_step = self.dist.n_peer - coll0[1]
        with torch.no_grad():                               # capture does not execute; make sure buffers are unchanged
            for k, v in buffers.items():
                sd[k].copy_(v)
        self._graph = {"graph": graph, "loss": loss, "sup": sup_static, "key": (id(pi), id(li), id(ev)), "refs": (pi, li, ev)}

    def _graph_step(self, pi, li, ev, sup):
        if self._graph is None or self._graph["key"] != (id(pi), id(li), id(ev)):
            self._capture(pi, li, ev, sup)
        g = self._graph
        g["sup"].copy_(sup, non_blocking=True)
        self.model._seed_buffer.fill_(int(torch.randint(0, 2 ** 62, (1,)).item()))
        g["graph"].replay()
        self.optimizer.step()
        self.model.bump_generation()              # the replay wrote parameters / buffers through raw pointers
        return g["loss"]

    def train_step(self, patient_indices, lab_indices, edge_values, supervision_mask) -> torch.Tensor:
        """Body of train_epoch (train.py:356-390) on device tensors; returns the loss tensor (no host sync)."""
        if getattr(self, "_use_graph", False) and self.model.training:
            return self._graph_step(patient_indices, lab_indices, edge_values, supervision_mask)
        self.optimizer.zero_grad()
        pred = self.model.predict_lab_values(self.data, patient_indices, lab_indices)
        loss = self._loss_of(pred, edge_values, lab_indices, supervision_mask)
        loss.backward()
        if self.grad_hook is not None:
            self.grad_hook()
        self.optimizer.step()
        self.model.bump_generation()
        return loss

    def _check_exchange(self):
        """After a host synchronisation: a timed-out peer rendezvous (csrc/peer.cuh) means every exchange since summed garbage --
        stop instead of training on."""
        if self.dist is not None and getattr(self.dist, "peer", None) is not None:
            self.dist.peer.check()

    def train_epoch(self, seed: Optional[int] = None) -> float:
        """train.py:332-392."""
        self.model.train()
        _, ev, _, sup = self.masker.get_masked_data("train", seed)
        pi, li = self.masker.split_rows("train")
        sup_dev = sup.to(self.device, non_blocking=True)
        loss = float(self.train_step(pi, li, ev, sup_dev).item())
        self._check_exchange()
        return loss

    @torch.no_grad()
    def validate(self, split: str = "val") -> float:
        """train.py:394-431."""
        self.model.eval()
        _, ev, _, _ = self.masker.get_masked_data(split)
        pi, li = self.masker.split_rows(split)
        pred = self.model.predict_lab_values(self.data, pi, li)
        out = float(compute_regression_loss(pred, ev, loss_type=self.loss_fn).item())
        self._check_exchange()
        return out

#!/usr/bin/env python
"""Benchmark of the hetero-GNN training hot path (BASELINE.json metric: train edges/sec, fwd+bwd, per
HeteroConv step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|torch_gpu] [--workload C4s8|C2|C3|C1]

Default workload, at every N: "C4s8" = one GPU's share (1/8: 1.25 M patients, 12.5 M lab edges) of BASELINE.json
configs[3], the patient-partitioned 10 M-patient / 100 M-edge graph the north_star's roofline and scaling targets are
stated on; with --gpus 8 the eight shards ARE that graph (weak scaling: per-GPU work fixed).  configs[1] (the
MIMIC-III-shaped graph "C2", 46,520 patients / 5 M lab edges, largely L2-resident) is `--workload C2`.

One "step" = one full training step of the reference's Trainer.train_epoch (train.py:332-392) on the
full graph: encode -> 2 x HeteroConv/BN/ReLU -> gated decoder over all train-split pairs -> weighted
loss over the 20 % supervised pairs -> backward -> Adam.  The unit of work is a *directed edge of one
HeteroConv layer application*: L * 2(E_l+E_d+E_m) per step.

value : device-timed (CUDA events), step inputs already resident in HBM.
e2e   : the same step through the public Trainer API with HOST inputs: the step's pair indices, targets and
        supervision mask are copied from pinned host memory every step and the loss is read back.
roofline : ONE HeteroConv layer forward + backward against the layer's compulsory HBM traffic bytes_min (the
           north_star's own yardstick); `roofline.dominant_kernel` keeps the per-launch numbers of the dominant libb2g kernel
           of one instrumented step (CUDA events around every library call).
cpu_baseline : the oracle port (oracle/hetero_rgcn_ref.py, torch CPU, all host threads): the full graph for C1 / C2
           (BASELINE.md section 3), a bounded patient sample for the larger workloads.
torch_eager_gpu : the same oracle port run eagerly on this B200 (PyTorch ATen kernels: index_select / index_add_ /
           cuBLAS) -- the "PyTorch-eager on the same GPU" comparator of SURVEY.md 2.1; `--impl torch_gpu` prints it alone.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "multi-modal-gnn_b200"
METRIC = "train_edges_per_sec_fwd_bwd_per_heteroconv_step"
UNIT = "directed-edge-layer traversals/s"
NUM_LAYERS = 2
DEFAULT_WORKLOAD = "C4s8"
# CPU oracle: full graph when it fits the box's memory / a few minutes, else 1/shrink of the patients (same density)
CPU_SHRINK = {"tiny": 1, "C1": 1, "C2": 1, "C3": 16, "C4s8": 16, "C4": 128, "C5": 128}
REF_MAX_STEPS = 5            # the CPU arm is bounded: at most this many timed oracle steps (a step takes seconds)


def _cfg(dropout=0.2, loss="mse"):
    return {"model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": NUM_LAYERS, "dropout": dropout,
                      "use_batch_norm": True, "activation": "relu"},
            "train": {"mask_fraction": 0.2, "train_split": 0.7, "val_split": 0.15, "test_split": 0.15, "loss": loss,
                      "epochs": 100, "early_stopping_patience": 15,
                      "optimizer": {"type": "adam", "lr": 1e-3, "weight_decay": 1e-5},
                      "lr_scheduler": {"enabled": True, "type": "reduce_on_plateau", "factor": 0.5, "patience": 10},
                      "seed": 42}}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm": float(p["hbm_gbs"]), "tensor_burst": float(p["bf16_tflops"]),
                "tensor_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on a bounded sample of the workload
# ------------------------------------------------------------------------------------------------------------------
def cpu_baseline_run(workload: str, steps: int, warmup: int, shrink: int = None, loss: str = "mse", device: str = "cpu"):
    """Times oracle train steps (forward + weighted loss + backward + Adam, dropout 0.2 like the shipped config) on the
    workload (or on a 1/shrink patient sample of it), with every host thread -- or, with device='cuda', eagerly on the GPU
    (the PyTorch-eager comparator).  Returns (edges_per_s, ms, sample, cores)."""
    import torch
    pkg = importlib.import_module(PKG)
    from oracle import hetero_rgcn_ref as R
    spec = pkg.synth.SPECS[workload]
    if shrink is None:
        shrink = CPU_SHRINK.get(workload, 16) if device == "cpu" else 1
    small = spec if shrink == 1 else pkg.synth.GraphSpec(
        spec.name + f"/{shrink}", max(spec.n_patient // shrink, 64), spec.n_lab, spec.n_dx, spec.n_med, spec.e_lab // shrink,
        spec.e_dx // shrink, spec.e_med // shrink, spec.low_degree_frac)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dev = torch.device(device)
    big = small.n_patient >= 500_000
    g = pkg.synth.make_graph(small, seed=42, device=("cuda" if (big and torch.cuda.is_available()) else "cpu")).to(dev)
    counts = {nt: int(g[nt].num_nodes) for nt in g.node_types}
    ets = list(g.edge_types)
    sd = {k: v.to(dev) for k, v in R.init_state(counts, ets, seed=0).items()}
    ei = g["patient", "has_lab", "lab"].edge_index
    attr = g["patient", "has_lab", "lab"].edge_attr
    tr = R.split_masks(ei.shape[1])[0].to(dev)
    pi, li, tgt = ei[0][tr], ei[1][tr], attr[tr].squeeze(-1)
    w = R.lab_weights(li.cpu(), tgt.cpu(), counts["lab"]).to(dev)
    keys = R.trainable_keys(sd)
    params = [sd[k].requires_grad_(True) for k in keys]
    opt = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-5)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if device != "cpu":
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        sup = R.supervision_mask(int(tr.sum()), 0.2, 1000 + it).to(dev)
        opt.zero_grad()
        pred = R.predict_lab_values(sd, counts, ets, g.edge_index_dict, pi, li, True, p_drop=0.2)
        lossv = R.weighted_loss(pred, tgt, li, w, sup, loss)
        lossv.backward()
        opt.step()
        float(lossv.detach())            # (a device -> host read: also the synchronisation point of the GPU variant)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    edges = NUM_LAYERS * small.directed_edges_per_layer
    sample = ((f"{workload} (full graph): " if shrink == 1 else f"{workload} shrunk 1/{shrink}: ")
              + f"{small.n_patient} patients, {small.e_lab}/{small.e_dx}/{small.e_med} edges, median of {steps} oracle train steps "
              + f"after {warmup} warm-up" + ("" if device == "cpu" else ", PyTorch eager on the GPU"))
    return edges / t, t * 1e3, sample, cores


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  /root/reference is absent on the GPU
    box and PyG is not installable, so this is the oracle port (kind 'port')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, REF_MAX_STEPS))
    v, ms, sample, cores = cpu_baseline_run(args.workload, steps, 1)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": args.workload, "sample": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_torch_gpu(args):
    """--impl torch_gpu: the oracle port run eagerly on the B200 (SURVEY.md 2.1 / 8d "secondary baseline": what PyTorch + PyG-style
    index_select / index_add_ / cuBLAS would do for the same step on the same GPU).  Informative; not the reference arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    line = {"impl": "torch_gpu", "metric": METRIC, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (ATen / cuBLAS eager)", "data": "synthetic",
            "gpu_launches": 0}
    line.update(torch_eager_gpu(args.workload, args.steps, max(args.warmup, 1)))
    line["config"] = {"workload": args.workload, "sample": line.pop("sample", None)}
    print(json.dumps(line), flush=True)


def torch_eager_gpu(workload: str, steps: int, warmup: int):
    import torch
    try:
        v, ms, sample, _ = cpu_baseline_run(workload, steps, warmup, shrink=1, device="cuda")
        out = {"value": v, "ms_per_step": ms, "sample": sample, "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
    except torch.OutOfMemoryError as exc:        # PyG-style message materialisation ([E, 128] fp32 per relation) does not fit
        out = {"value": None, "ms_per_step": None, "unavailable": f"CUDA out of memory on the full {workload} graph: {str(exc)[:120]}"}
    torch.cuda.empty_cache()
    return out


# C-ABI call -> the kernel that does its work.  Several entry points share one kernel (the tcgen05 linear kernel serves
# nn.Linear forward, its input gradient and the dense-adjacency products; the tcgen05 MN-major kernel serves weight
# gradients and the transposed adjacency products): the roofline is reported for the dominant KERNEL.
KERNEL_OF = {
    "b2g_linear_fwd_tc": "k_linear_tf32", "b2g_linear_bwd_input_tc": "k_linear_tf32", "b2g_adjacency_mma_fwd": "k_linear_tf32",
    "b2g_linear_bwd_weight_tc": "k_wgrad_tf32+k_wgrad_tc_reduce", "b2g_adjacency_mma_bwd": "k_wgrad_tf32+k_wgrad_tc_reduce",
    "b2g_bn_stats": "k_col_reduce<0>", "b2g_col_sums": "k_col_reduce<0>", "b2g_bn_stats_sync": "k_col_reduce<0>",
    "b2g_bn_local_sums": "k_col_reduce<0>",
    "b2g_bn_bwd": "k_col_reduce<1>+k_bn_bwd_apply", "b2g_bn_bwd_sync": "k_col_reduce<1>+k_bn_bwd_apply",
    "b2g_linear_fwd": "k_sgemm_small", "b2g_linear_bwd_input": "k_sgemm_small", "b2g_linear_bwd_weight": "k_sgemm_small",
    "b2g_decoder_fwd_tc": "k_decoder_fwd_tc", "b2g_decoder_bwd_tc": "k_decoder_bwd_tc", "b2g_bn_apply": "k_bn_apply",
    "b2g_layer_fwd_tc": "k_layer_tf32", "b2g_layer_adjT_tc": "k_adjT_tf32+k_adjT_reduce",
}


def _measured_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu captures (profiles/r2_traffic.json, else r1_traffic.json), or None."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            if kernel in t:
                return t[kernel]
        except Exception:
            pass
    return None


# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = os.path.join("/tmp", f"b2g_clocks_{os.getpid()}.csv")

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                parts = [p.strip() for p in ln.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1])); mx.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            busy = [c for c, m in zip(sm, mx) if c >= 0.4 * m] or sm          # samples taken while the GPU was clocked up
            out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm), "samples_under_load": len(busy)}
        return out


def layer_microbench(model, trainer, spec, dev, flush, peak_gbs, reps=5):
    """ONE HeteroConv layer (all six relations incl. their SAGE linears, no BatchNorm / decoder),
    forward + backward, against the layer's compulsory HBM traffic bytes_min = 5 N_p d 4 + 2 E 4 + 6 (N_p + 1) 4 (read x_p,
    write out_p, CSR once; backward: read grad_out_p and x_p, write grad_x_p, CSR again).  Each repetition is enqueued behind
    a spin kernel, so the event pair brackets back-to-back device work.  Single GPU only."""
    import torch
    gi = model._graph_index(trainer.data)
    d = model.hidden_dim
    gen = torch.Generator(device=dev).manual_seed(7)
    counts = {nt: int(trainer.data[nt].num_nodes) for nt in trainer.data.node_types}
    x = {nt: torch.randn(n, d, device=dev, generator=gen).requires_grad_(True) for nt, n in counts.items()}
    gout = {nt: torch.randn(n, d, device=dev, generator=gen) for nt, n in counts.items()}
    params = [p for p in model.convs[0].parameters()]

    def once():
        for t in list(x.values()) + params:
            t.grad = None
        out = model._layer(0, x, gi)
        torch.autograd.backward([out[nt] for nt in out], [gout[nt] for nt in out])

    was_training = model.training
    model.train()
    for _ in range(2):
        once()
    torch.cuda.synchronize()
    times = []
    for i in range(reps):
        flush.fill_(i & 0xFF)
        torch.cuda._sleep(int(0.010 * 1.9e9))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        once()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    for t in params:
        t.grad = None
    model.train(was_training)
    ms = statistics.median(times)
    e_und = spec.e_lab + spec.e_dx + spec.e_med
    bytes_min = 5 * spec.n_patient * d * 4 + 2 * e_und * 4 + 6 * (spec.n_patient + 1) * 4
    return {"what": "one HeteroConv layer (6 relations + SAGE linears) forward + backward, eager behind a spin kernel, median of %d" % reps,
            "ms": ms, "directed_edges_per_s": spec.directed_edges_per_layer / (ms * 1e-3), "bytes_min": bytes_min,
            "achieved_GBps_of_bytes_min": bytes_min / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": bytes_min / (ms * 1e-3) / 1e9 / peak_gbs}


def run_ours(args):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    ops = importlib.import_module(PKG + ".ops")
    M = importlib.import_module(PKG + ".model")
    T = importlib.import_module(PKG + ".trainer")
    L = importlib.import_module(PKG + "._lib")
    lib = L.load()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    # stdout must carry exactly one JSON line (rank 0): park fd 1 on stderr while libraries (NCCL's version banner) may
    # print, restore it for the final line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL's own log (NCCL_DEBUG, if the caller sets it) goes wherever NCCL sends it: fd 1 is parked on stderr until the
        # final JSON line, so stdout still carries exactly one line
        dist.init_process_group("nccl", device_id=dev, timeout=__import__("datetime").timedelta(seconds=90))

    spec = pkg.synth.SPECS[args.workload]
    cfg = _cfg(dropout=0.2, loss="mse")
    # Weak scaling over patients: every rank owns a patient partition of the same shape (its own
    # patients, edges and train pairs; seed differs per rank) of ONE global graph whose lab / diagnosis / medication nodes,
    # dense weights and BatchNorm parameters are replicated.  The exact mode is used: partial type sums, patient
    # BatchNorm statistics and replicated->local gradients are all-reduced so that N ranks compute what one GPU would
    # compute on the union graph (checked by tools/dist_check.py), plus one flat gradient all-reduce per step.
    D = importlib.import_module(PKG + ".dist")
    dctx = D.DistContext(device=dev) if world > 1 else None
    # the large configs are drawn on the GPU (the CPU generator needs minutes at >= 1 M patients); C1 / C2 on the host
    g_host = pkg.synth.make_graph(spec, seed=42 + rank, device=dev if spec.n_patient >= 500_000 else "cpu")
    masker = T.EdgeMasker(g_host, 0.7, 0.15, 0.15, 0.2, 42 + rank)
    torch.manual_seed(0)
    model = M.build_model(cfg, (g_host.node_types, g_host.edge_types), None)
    trainer = T.Trainer(model, g_host, masker, cfg, dev, dist_ctx=dctx)
    model._init_embeddings(trainer.data)          # tables exist before the first timed step (lazy init is not timed)
    if world > 1:
        for name, p in model.named_parameters():
            if not name.startswith("embeddings.patient"):
                dist.broadcast(p.data, 0)

    pi, li = masker.split_rows("train")
    _, ev = masker.split_edges("train")
    n_train = int(pi.numel())
    total_steps = args.warmup + args.steps
    sup_host = [masker.supervision_mask("train", seed=1000 + i).pin_memory() for i in range(total_steps)]
    sup_dev = [s.to(dev) for s in sup_host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    use_graph = not args.no_graph        # for N > 1 the step's NCCL all-reduces are captured into the graph as well
    if use_graph:
        trainer.enable_cuda_graph()

    def step_device(i):
        return trainer.train_step(pi, li, ev, sup_dev[i])

    model.train()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()          # sampled over warm-up + both timed regions (nvidia-smi needs ~1 s to produce its first line)
    # ---- warm-up ----
    for i in range(args.warmup):
        step_device(i)
    torch.cuda.synchronize()

    # ---- timed: device-resident inputs ----
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    lib.b2g_reset_launch_count()
    evs = []
    for i in range(args.warmup, total_steps):
        flush.fill_(i & 0xFF)                       # L2 flush (256 MiB write) outside the per-step events
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        loss = step_device(i)
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = int(lib.b2g_launch_count())            # eager launches in the timed region ...
    if use_graph:
        launches += trainer.graph_kernel_nodes * args.steps   # ... plus the libb2g kernel nodes of every graph replay
    step_ms = [s.elapsed_time(e) for s, e in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    final_loss = float(trainer.global_loss(loss).item())

    # ---- timed: end to end through the public Trainer API with host inputs ----
    # Per-step host input of Trainer.train_epoch is the supervision mask the host-side EdgeMasker draws for that epoch
    # (train.py:150-159); the split's pair indices / targets are dataset state uploaded once with the graph
    # (Trainer.__init__: data.to(device)), exactly like the reference.  Each step: pinned host mask -> device, one
    # train_step, loss read back to the host.
    sup_d = torch.empty(n_train, dtype=torch.bool, device=dev)
    h2d = n_train
    e2e_steps = max(3, args.steps)

    def step_e2e(i):
        sup_d.copy_(sup_host[i % total_steps], non_blocking=True)
        loss = trainer.train_step(pi, li, ev, sup_d)
        return float(loss.item())                  # device -> host read of the step's result

    step_e2e(0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        step_e2e(i)
    torch.cuda.synchronize()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_t.item()) * 1e3 / e2e_steps
    clk = clocks.stop() if rank == 0 else None

    # ---- one instrumented step: per-kernel CUDA-event durations -> dominant kernel + roofline ----
    roof, kernels = None, None
    trainer.enable_cuda_graph(False)             # per-call events cannot live inside a captured graph
    step_device(args.warmup)                     # (every rank runs these two steps: they contain collectives)
    ops.PROFILE = []
    flush.fill_(1)
    # Park the GPU behind a ~40 ms spin while the host enqueues the whole step: every event pair then brackets kernels
    # that run back to back, so the per-call durations are device time, not host launch latency (the step is ~270 short
    # launches; issued into an idle GPU the tiny type-row kernels would be charged ~10 us of launch gap each).
    torch.cuda._sleep(int(0.040 * 1.9e9))
    step_device(args.warmup)
    torch.cuda.synchronize()
    if rank != 0:
        ops.PROFILE = None
    if rank == 0:
        prof, ops.PROFILE = ops.PROFILE, None
        agg = {}
        for name, s, e, nbytes, flops in prof:
            a = agg.setdefault(name, [0.0, 0, 0, 0])
            a[0] += s.elapsed_time(e); a[1] += 1; a[2] += nbytes; a[3] += flops
        tot = sum(a[0] for a in agg.values())
        kernels = {k: {"ms": round(a[0], 4), "calls": a[1], "share": round(a[0] / tot, 4),
                       "GBps": round(a[2] / a[0] / 1e6, 1) if a[0] > 0 else None,
                       "TFLOPs": round(a[3] / a[0] / 1e9, 2) if a[0] > 0 else None}
                   for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])}
        fam = {}
        for k, a in agg.items():
            f = fam.setdefault(KERNEL_OF.get(k, k), [0.0, 0, 0, 0, []])
            f[0] += a[0]; f[1] += a[1]; f[2] += a[2]; f[3] += a[3]; f[4].append(k)
        top = max(fam.items(), key=lambda kv: kv[1][0])
        peaks = _peaks()
        name, (ms, calls, nbytes, flops, members) = top[0], top[1]
        ach = nbytes / calls / (ms / calls * 1e-3) / 1e9          # algorithmic bytes per launch / average launch duration
        tr = _measured_traffic(name)
        roof_kernel = {"bound": "hbm", "kernel": name, "entry_points": sorted(members), "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                "frac": ach / peaks["hbm"], "traffic": (tr or {}).get("dram_bytes_per_launch"), "traffic_source": (tr or {}).get("source"),
                "algorithmic_bytes_per_launch": nbytes / calls, "launches_per_step": calls, "avg_launch_ms": ms / calls,
                "share_of_step": a_share(ms, tot), "tensor_TFLOPs": flops / (ms * 1e-3) / 1e12,
                "timing": "CUDA events around every library call of one eager step enqueued behind a 40 ms spin kernel (device time, no launch gaps), L2 flushed before the step",
                "peak_source": peaks["source"] + " (MEASURED_PEAKS.json hbm_gbs)" if peaks["source"] == "measured" else "fallback"}
        roof = dict(roof_kernel)                       # N > 1: the dominant kernel's numbers; N = 1: replaced by the layer roofline below

    if world > 1 and dctx.peer is not None:
        dctx.peer.check()                          # a timed-out rendezvous would have produced garbage: fail loudly
    if rank == 0:
        parallelism = None
        if world > 1:
            per_step = dctx.n_collectives_per_step or 0
            parallelism = (f"patient-partitioned x{world}, exact mode: {per_step} exchanges per step (incl. the gradient all-reduce), "
                           + (f"{dctx.n_peer_per_step} of them one-shot NVLink peer-memory kernels of libb2g (BatchNorm ones fused into the "
                              f"reduction kernel), rest NCCL" if dctx.peer is not None else "all NCCL" + (f" (peer-memory communicator unavailable: {dctx.peer_unavailable})" if dctx.peer_unavailable else "")) + ", captured in the step's CUDA graph")
        edges_per_step = NUM_LAYERS * spec.directed_edges_per_layer * world
        line = {"metric": METRIC, "value": edges_per_step / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (large-M linears: TF32 operands on tcgen05, fp32 accumulate)" if ops.PRECISION == "tf32" else "f32",
                "data": "synthetic",
                "config": {"workload": f"{args.workload}: {spec.n_patient} patients/{spec.n_lab} labs/{spec.n_dx} dx/{spec.n_med} meds, "
                                       f"{spec.e_lab}/{spec.e_dx}/{spec.e_med} edges per GPU, d=128, L=2, dropout 0.2, mse + lab weights, "
                                       f"{n_train} train pairs, 20% supervised, Adam",
                           "step": "Trainer.train_step: predict_lab_values fwd + weighted loss + bwd" + (" (one CUDA graph replay)" if use_graph else "") + " + Adam",
                           "l2": "flushed with a 256 MiB write before every timed step; per-step working set (>1 GB) also exceeds the 126 MB L2",
                           "parallelism": (parallelism if world > 1 else "single GPU")},
                "e2e": {"value": edges_per_step / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "gpu_launches": launches, "clocks": clk, "roofline": roof, "kernels": kernels, "final_loss": final_loss,
                "step_ms_min_max": [min(step_ms), max(step_ms)]}
        if world == 1:
            # the north_star's yardstick: ONE HeteroConv layer forward + backward against its compulsory traffic
            peaks = _peaks()
            try:
                mb = layer_microbench(model, trainer, spec, dev, flush, peaks["hbm"])
                tr = _measured_traffic("hetero_layer_fwd_bwd:" + args.workload)
                line["roofline"] = {"bound": "hbm", "kernel": "one HeteroConv layer, forward + backward (k_layer_tf32 x2, k_adjT_tf32 x2, "
                                    "k_wgrad_tf32, grouped type-row GEMMs)", "achieved": mb["achieved_GBps_of_bytes_min"],
                                    "peak": peaks["hbm"], "unit": "GB/s", "frac": mb["frac_of_hbm_peak"],
                                    "traffic": (tr or {}).get("dram_bytes_per_layer"), "traffic_source": (tr or {}).get("source"),
                                    "algorithmic_bytes": mb["bytes_min"], "ms": mb["ms"], "directed_edges_per_s": mb["directed_edges_per_s"],
                                    "what": mb["what"] + "; algorithmic bytes = bytes_min = 5 N_p d 4 + 2 E 4 + 6 (N_p + 1) 4 (SURVEY.md 8d)",
                                    "peak_source": peaks["source"] + " (MEASURED_PEAKS.json hbm_gbs)" if peaks["source"] == "measured" else "fallback",
                                    "dominant_kernel": roof}
            except Exception as exc:      # noqa: BLE001   (an extra must never lose the line)
                line["roofline"] = dict(roof or {}, layer_error=f"{type(exc).__name__}: {exc}")
            # the parity-claim mode (exact fp32 kernels: aggregation <= 1e-5) timed on the same workload, eagerly
            try:
                trainer.enable_cuda_graph(False)
                ops.set_precision("fp32")
                step_device(0)
                torch.cuda.synchronize()
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record()
                for i in range(2):
                    step_device(i)
                e_.record()
                torch.cuda.synchronize()
                ms32 = s_.elapsed_time(e_) / 2
                line["exact_fp32_mode"] = {"ms_per_step": ms32, "value": edges_per_step / (ms32 * 1e-3), "unit": UNIT, "steps": 2,
                                           "what": "the same step with every product in exact fp32 (SIMT kernels, CSR gather-reduce): the mode the 1e-5 "
                                                   "aggregation / golden-vector parity claims are made in; eager launches"}
            except Exception as exc:      # noqa: BLE001
                line["exact_fp32_mode"] = {"error": f"{type(exc).__name__}: {exc}"}
            finally:
                ops.set_precision("tf32")
        if world == 1 and not args.no_cpu_baseline:
            # free the GPU arm's memory before the comparators run
            del trainer, model
            torch.cuda.empty_cache()
            te = torch_eager_gpu(args.workload, 2, 1)
            te["unit"] = UNIT
            line["torch_eager_gpu"] = te
            v, ms, sample, cores = cpu_baseline_run(args.workload, 3, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "ms_per_step": ms}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_c5(args):
    """--workload C5: BASELINE config 5, inference-only bulk imputation of all held-out (val + test, 30 %) patient-lab pairs on
    the 100 M-edge graph, 256-d hidden, patients partitioned over the ranks (each rank: a C4s8-shaped share).  One step = one
    full imputation pass: eval-mode encode + 2 HeteroConv layers (the only exchanges: partial type sums) + the gated decoder
    over the rank's held-out pairs (inference.py:140-159 recomputes the forward for every call; so does a step).
    metric: imputed pairs per second, whole job."""
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    ops = importlib.import_module(PKG + ".ops")
    M = importlib.import_module(PKG + ".model")
    T = importlib.import_module(PKG + ".trainer")
    D = importlib.import_module(PKG + ".dist")
    L = importlib.import_module(PKG + "._lib")
    lib = L.load()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev, timeout=__import__("datetime").timedelta(seconds=90))
    share = pkg.synth.SPECS["C4s8"]
    spec = pkg.synth.GraphSpec("C5s8", share.n_patient, share.n_lab, share.n_dx, share.n_med, share.e_lab, share.e_dx, share.e_med,
                               share.low_degree_frac, hidden_dim=256)
    cfg = _cfg(dropout=0.2, loss="mse")
    cfg["model"]["hidden_dim"] = 256
    dctx = D.DistContext(device=dev) if world > 1 else None
    g = pkg.synth.make_graph(spec, seed=42 + rank, device=dev)
    masker = T.EdgeMasker(g, 0.7, 0.15, 0.15, 0.2, 42 + rank)
    torch.manual_seed(0)
    model = M.build_model(cfg, (g.node_types, g.edge_types), None).to(dev)
    if dctx is not None:
        model.set_distributed(dctx)
    model._init_embeddings(g)
    if world > 1:
        for name, p in model.named_parameters():
            if not name.startswith("embeddings.patient"):
                dist.broadcast(p.data, 0)
    model.eval()
    held = (masker.val_mask | masker.test_mask).to(dev)
    ei = g["patient", "has_lab", "lab"].edge_index
    pi, li = ei[0][held].contiguous(), ei[1][held].contiguous()
    n_pairs = int(pi.numel())
    pi_h, li_h = pi.cpu().pin_memory(), li.cpu().pin_memory()
    out_h = torch.empty(n_pairs, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            pred = model.predict_lab_values(g, pi, li)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        lib.b2g_reset_launch_count()
        evs = []
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            pred = model.predict_lab_values(g, pi, li)
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        launches = int(lib.b2g_launch_count())
        tot = torch.tensor([sum(s.elapsed_time(e) for s, e in evs)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        ms = float(tot.item()) / args.steps
        # end to end: pair lists from pinned host memory, predictions back to the host
        pd_, ld_ = torch.empty_like(pi), torch.empty_like(li)

        def e2e_step():
            pd_.copy_(pi_h, non_blocking=True)
            ld_.copy_(li_h, non_blocking=True)
            out_h.copy_(model.predict_lab_values(g, pd_, ld_), non_blocking=True)
            torch.cuda.synchronize()

        e2e_step()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        n_e2e = max(3, min(args.steps, 10))
        for _ in range(n_e2e):
            e2e_step()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e_ms = float(t_e2e.item()) * 1e3 / n_e2e
    clk = clocks.stop() if rank == 0 else None
    n_all = torch.tensor([n_pairs], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(n_all)
        if dctx.peer is not None:
            dctx.peer.check()
    if rank == 0:
        total_pairs = int(n_all.item())
        e_und = spec.e_lab + spec.e_dx + spec.e_med
        d = 256
        bytes_min = NUM_LAYERS * (2 * spec.n_patient * d * 4 + e_und * 4 + 3 * (spec.n_patient + 1) * 4) + n_pairs * (16 + 4)
        peaks = _peaks()
        line = {"metric": "bulk_imputation_pairs_per_sec", "value": total_pairs / (ms * 1e-3), "unit": "imputed (patient, lab) pairs/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (large-M linears with <= 227 KB of weights: TF32 operands on tcgen05)", "data": "synthetic",
                "config": {"workload": f"C5 (BASELINE configs[4]): eval-mode bulk imputation, d=256, L=2; per GPU {spec.n_patient} patients, "
                                       f"{spec.e_lab}/{spec.e_dx}/{spec.e_med} edges, {n_pairs} held-out (val+test) pairs; x{world} GPUs = "
                                       f"{total_pairs} pairs" + (" = the 100 M-edge graph" if world == 8 else ""),
                           "step": "HeteroRGCN.predict_lab_values in eval mode, node embeddings recomputed every step (no cache)",
                           "l2": "flushed with a 256 MiB write before every timed step", "parallelism": (f"patient-partitioned x{world}, the only "
                           f"exchanges are the per-layer partial type sums" if world > 1 else "single GPU")},
                "e2e": {"value": total_pairs / (e2e_ms * 1e-3), "unit": "imputed (patient, lab) pairs/s", "ms_per_step": e2e_ms, "steps": n_e2e,
                        "h2d_bytes_per_step": 16 * n_pairs, "d2h_bytes_per_step": 4 * n_pairs},
                "gpu_launches": launches, "clocks": clk,
                "roofline": {"bound": "hbm", "kernel": "whole imputation pass", "achieved": bytes_min / (ms * 1e-3) / 1e9, "peak": peaks["hbm"],
                             "unit": "GB/s", "frac": bytes_min / (ms * 1e-3) / 1e9 / peaks["hbm"], "traffic": None,
                             "algorithmic_bytes": bytes_min, "what": "compulsory traffic of the pass: per layer read x_p + write out_p + adjacency "
                             "once, plus the pair list and the predictions (SURVEY.md 8d, forward half)"}}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def a_share(ms, tot):
    return round(ms / tot, 4) if tot > 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue the step launch by launch instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch_gpu":
        run_torch_gpu(args)
    elif args.workload == "C5":
        run_c5(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the ICL hot path (BASELINE.json metric: 3D U-Net ICL train voxels/s at 96^3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference|reference-cuda] [--workload cfg2|cfg3|infer]
                    [--precision parity|fast]

Workloads (BASELINE.json configs):
  cfg2  (configs[1], default at N = 1): unet_3D_icl(n_classes=2, in_channels=1), per rank 2 labeled + 2 unlabeled synthetic 1x96^3
        patches; one step = forward (two backbone passes + SSPA/USCL heads) + the five ICL losses + backward + momentum-SGD step with
        poly LR  (reference loop: train_inherent_consistent_unet_3D_BraTS.py:103-119, loss weights 1,1,1,1,10).
  cfg3  (configs[2], default at N > 1): the same step with n_classes=16 and the AMOS loss weights 1,1,1,0.1,10
        (train_inherent_consistent_unet_3D_AMOS22.py:230), data parallel (icl_b200.parallel), weak scaling.
  infer (configs[4]): unet_3D(n_classes=2).eval(), one 240x240x155 volume, 96^3 windows at stride 48 (48 windows), windows sharded
        round-robin over ranks (test_3D_BraTS.py:79-142); one step = one volume; strong scaling.
A default N = 1 run additionally measures cfg3 on the same GPU (key "also"), so that the N > 1 lines (cfg3) have their own
single-GPU denominator measured in the same driver run.

Prints ONE JSON line (rank 0).  `value` = whole-job voxels/s with inputs resident in HBM; `e2e` = the same step driven from pinned
HOST buffers (H2D of that step's volume + labels and D2H of the loss inside the timed region; the H2D of step i+1 is issued on a copy
stream while step i computes).  `--impl reference` times the UNMODIFIED reference (baseline/_ref, or /root/reference in the build
container) on all host cores; if neither tree is present it falls back to the oracle port (kind "port").  `--impl reference-cuda`
runs the same unmodified reference modules under stock PyTorch CUDA (cuDNN / cuBLAS) — the library bar on the same GPU.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VOX = 96 ** 3
LABELED_BS, BATCH = 2, 4
WORKLOADS = {
    # K, loss weights (dice, ce, aux, pse, cons), base_lr, max_iterations, algorithmic conv GFLOP per step
    "cfg2": dict(K=2, weights=(1.0, 1.0, 1.0, 1.0, 10.0), base_lr=0.01, max_iters=30000, conv_gflop=1166.9,
                 desc="config2: unet_3D_icl(K=2,in=1) ICL train step (fwd + 5 losses + bwd + SGD), per-rank batch 4 (2 lab + 2 unlab) x 1x96^3"),
    "cfg3": dict(K=16, weights=(1.0, 1.0, 1.0, 0.1, 10.0), base_lr=0.02, max_iters=60000, conv_gflop=1170.1,
                 desc="config3: unet_3D_icl(K=16,in=1) ICL train step (fwd + 5 losses + bwd + SGD, AMOS loss weights), per-rank batch 4 "
                      "(2 lab + 2 unlab) x 1x96^3"),
}
INFER = dict(K=2, shape=(240, 240, 155), patch=(96, 96, 96), stride=48, conv_gflop_per_window=121.98,
             desc="config5: unet_3D(K=2).eval() sliding-window inference, one 240x240x155 volume, 96^3 windows, stride 48 (48 windows)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="icl_b200", choices=["icl_b200", "reference", "reference-cuda"])
    ap.add_argument("--workload", default="", choices=["", "cfg2", "cfg3", "infer"])
    ap.add_argument("--precision", default="parity", choices=["parity", "fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="N = 1 default run: skip the secondary cfg3 measurement")
    ap.add_argument("--unfused-optimizer", action="store_true", help="materialise the mlp2 weight gradients (reference-style .grad) instead of "
                    "applying them as rank-R updates inside the optimizer")
    ap.add_argument("--no-graph", action="store_true", help="eager step (one Python-enqueued launch per kernel) instead of CUDA-graph replay")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: exchange gradients after backward instead of overlapping with it")
    ap.add_argument("--window-batch", type=int, default=4, help="infer: windows per backbone launch")
    ap.add_argument("--detail", default="", help="write the per-(kernel, shape) CUDA-event breakdown to this JSON file")
    return ap.parse_args()


def default_workload(a, world):
    return a.workload or ("cfg2" if world == 1 else "cfg3")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------ reference arms
def _state_shapes(K):
    from collections import OrderedDict
    keys = json.load(open(os.path.join(ROOT, "tests", "golden", "state_keys.json")))["unet_3D_icl_k2"]
    shapes = OrderedDict()
    for k, s in keys:
        s = list(s)
        if k in ("final.weight", "final.bias"):
            s[0] = K
        elif k.endswith("guided_Q"):
            s[1] = K
        shapes[k] = tuple(s)
    return shapes


def reference_available():
    from oracle import ref_import
    return ref_import.available()


def ref_train_step_factory(wl, device, threads):
    """The UNMODIFIED reference: networks.unet_3D_icl.unet_3D_icl + utils.losses + torch.optim.SGD, driven exactly like
    train_inherent_consistent_unet_3D_BraTS.py:100-119 (AMOS22.py:222-236 for the loss weights of cfg3)."""
    import torch
    from icl_b200.utils import synth
    from oracle import ref_import
    torch.set_num_threads(threads)
    ns = ref_import.load()
    cfg = WORKLOADS[wl]
    K = cfg["K"]
    model = ns.unet_3D_icl(feature_scale=4, n_classes=K, in_channels=1)
    synth.load_synth(model, 1337)
    model.to(device).train()
    opt = torch.optim.SGD(model.parameters(), lr=cfg["base_lr"], momentum=0.9, weight_decay=0.0001)
    ce_loss = torch.nn.CrossEntropyLoss()
    dice_loss, aux_loss, pse_loss = ns.losses.DiceLoss(K), ns.losses.AuxLoss3D(K), ns.losses.PseudoSoftLoss3D(K)
    x = synth.synth_volume((BATCH, 1, 96, 96, 96), 1338).to(device)
    y = synth.synth_labels((BATCH, 96, 96, 96), K, 1339).to(device)
    w = cfg["weights"]
    it = [0]

    def step():
        outputs = model(x[:LABELED_BS], x[LABELED_BS:])
        outputs_soft = torch.softmax(outputs[0], dim=1)
        loss_ce = ce_loss(outputs[0], y[:LABELED_BS])
        loss_dice = dice_loss(outputs_soft, y[:LABELED_BS].unsqueeze(1))
        loss_aux = aux_loss(outputs[2], y[:LABELED_BS])
        loss_pse = pse_loss(outputs[3], outputs[1])
        loss_cons = ns.losses.softmax_mse_loss(outputs[3], outputs[4])
        loss = w[0] * loss_dice + w[1] * loss_ce + w[2] * loss_aux + w[3] * loss_pse + w[4] * loss_cons
        opt.zero_grad()
        loss.backward()
        opt.step()
        lr_ = cfg["base_lr"] * (1.0 - it[0] / cfg["max_iters"]) ** 0.9
        for g in opt.param_groups:
            g["lr"] = lr_
        it[0] += 1
        return float(loss.detach())
    return step


def port_train_step_factory(wl, threads):
    """Fallback when no reference tree is reachable: the oracle restatement of the same step (kind "port")."""
    import torch
    from icl_b200.utils import synth
    from oracle import restate as R
    torch.set_num_threads(threads)
    cfg = WORKLOADS[wl]
    K = cfg["K"]
    P = R.make_params(synth.synth_state_dict(_state_shapes(K), 1337))
    x = synth.synth_volume((BATCH, 1, 96, 96, 96), 1338)
    y = synth.synth_labels((BATCH, 96, 96, 96), K, 1339)
    bufs = {}
    it = [0]

    def step():
        rand = R.TorchRand()
        L, grads, _ = R.train_step_3d(P, x, y, LABELED_BS, K, weights=dict(zip(("dice", "ce", "aux", "pse", "cons"), cfg["weights"])), rand=rand)
        lr = cfg["base_lr"] if it[0] == 0 else R.poly_lr(cfg["base_lr"], it[0] - 1, cfg["max_iters"])
        R.sgd_step(P, grads, bufs, lr)
        it[0] += 1
        return float(L["total"].detach())
    return step


def ref_infer_factory(device, threads):
    """Reference test_single_case (test_3D_BraTS.py:79-142) on the reference unet_3D, one window per net call."""
    import torch
    from icl_b200.utils import synth
    torch.set_num_threads(threads)
    image = synth.synth_volume(INFER["shape"], 78).numpy()
    if reference_available():
        from oracle import ref_import
        ns = ref_import.load()
        net = ns.unet_3D(feature_scale=4, n_classes=INFER["K"], in_channels=1)
        synth.load_synth(net, 77)
        net.to(device).eval()
        if device == "cpu":
            # test_single_case calls .cuda() on every patch (test_3D_BraTS.py:123): on the host the restated window loop of the
            # oracle drives the unmodified reference network instead
            from oracle import restate as R
            return (lambda: R.test_single_case(lambda p: net(p), image, INFER["stride"], INFER["stride"], INFER["patch"], INFER["K"])), "reference"
        tsc = ref_import.load_test_single_case()
        return (lambda: tsc(net, image, INFER["stride"], INFER["stride"], INFER["patch"], num_classes=INFER["K"])), "reference"
    from collections import OrderedDict
    from oracle import restate as R
    keys = json.load(open(os.path.join(ROOT, "tests", "golden", "state_keys.json")))["unet_3D_icl_k2"]
    shapes = OrderedDict((k, tuple(s)) for k, s in keys if not (k.startswith("sspa") or k.startswith("uscl")))
    P = R.make_params(synth.synth_state_dict(shapes, 77), requires_grad=False)
    return (lambda: R.test_single_case(lambda p: R.unet_3d_forward(P, p), image, INFER["stride"], INFER["stride"], INFER["patch"], INFER["K"])), "port"


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import torch
    wl = default_workload(a, max(world, a.gpus))
    cuda = a.impl == "reference-cuda"
    device = "cuda" if cuda else "cpu"
    threads = os.cpu_count() or 1
    if cuda and not (torch.cuda.is_available() and reference_available()):
        print(json.dumps({"impl": a.impl, "unavailable": "needs a CUDA device and the reference tree (baseline/_ref)"}), flush=True)
        return
    with torch.no_grad() if wl == "infer" else torch.enable_grad():
        if wl == "infer":
            step, kind = ref_infer_factory(device, threads)
            units, desc = float(INFER["shape"][0] * INFER["shape"][1] * INFER["shape"][2]), INFER["desc"]
            steps, warmup = (max(1, a.steps), max(1, a.warmup)) if cuda else (1, 0)
        else:
            if reference_available():
                step, kind = ref_train_step_factory(wl, device, threads), "reference"
            else:
                step, kind = port_train_step_factory(wl, threads), "port"
            units, desc = float(BATCH * VOX), WORKLOADS[wl]["desc"]
            steps, warmup = (max(1, a.steps), max(1, a.warmup)) if cuda else (max(1, min(a.steps, 3)), min(a.warmup, 1))
        for _ in range(warmup):
            step()
        ts = []
        for _ in range(steps):
            if cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            step()
            if cuda:
                torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    v = units / t
    sample = "%d warm-up + %d timed full steps of %s, median" % (warmup, steps, wl)
    line = {
        "impl": a.impl, "metric": "3D U-Net ICL train voxels/s (96^3)" if wl != "infer" else "3D U-Net sliding-window inference voxels/s",
        "value": v, "unit": "voxels/s", "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak" if wl != "infer" else "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": desc, "device": "stock PyTorch CUDA (cuDNN/cuBLAS) on 1 GPU" if cuda else "%d host threads" % threads},
        "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if cuda:
        del line["cpu_baseline"]
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


class Timer:
    """K calls of fn bracketed by barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks."""

    def __init__(self, world, dev):
        self.world, self.dev = world, dev
        self.host_ms = 0.0

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def __call__(self, fn, n):
        import torch
        import torch.distributed as dist
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        self.host_ms = (time.perf_counter() - t0) * 1e3 / n   # host time to ENQUEUE one step (no sync inside)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / n


class TrainBench:
    """One workload (cfg2 / cfg3) of the training step on this rank's GPU."""

    def __init__(self, a, wl, rank, world, dev, exchange=True):
        from icl_b200 import _lib, parallel
        from icl_b200.networks.unet_3D_icl import unet_3D_icl
        from icl_b200.optim import SGD
        from icl_b200.utils import losses as L
        from icl_b200.utils import synth
        self.a, self.wl, self.cfg, self.rank, self.world, self.dev = a, wl, WORKLOADS[wl], rank, world, dev
        cfg = self.cfg
        K = cfg["K"]
        self.net = unet_3D_icl(feature_scale=4, n_classes=K, in_channels=1)
        synth.load_synth(self.net, 1337)
        self.net.to(dev).train()
        # fused_factored: the 13 824^2 mlp2 weight gradients are applied as rank-R updates inside the optimizer
        self.opt = SGD(self.net.parameters(), lr=cfg["base_lr"], momentum=0.9, weight_decay=1e-4, fused_factored=not a.unfused_optimizer)
        self.dp = None
        if world > 1 and exchange:
            self.dp = parallel.GradAverager(self.net, world, factored=a.unfused_optimizer, overlap=not a.no_overlap)
        self.aux_loss, self.pse_loss = L.AuxLoss3D(K), L.PseudoSoftLoss3D(K)
        self.L = L
        self.x_host = synth.synth_volume((BATCH, 1, 96, 96, 96), 1338 + rank).pin_memory()
        # only the labeled half of the label batch is ever read (label_batch[:labeled_bs], train_..._BraTS.py:107-109)
        self.y_host = synth.synth_labels((BATCH, 96, 96, 96), K, 1339 + rank)[:LABELED_BS].contiguous().pin_memory()
        self.x_dev, self.y_dev = self.x_host.to(dev), self.y_host.to(dev)
        self.it = 0
        self.graphed = None
        self.graph_launches = 0
        self._lib = _lib

    def lr_now(self):
        return self.cfg["base_lr"] * (1.0 - self.it / self.cfg["max_iters"]) ** 0.9

    def eager_step(self, volume_batch, label_lab):
        L, w = self.L, self.cfg["weights"]
        net, opt = self.net, self.opt
        if self.dp is not None:
            self.dp.begin_step()
        outputs = net(volume_batch[:LABELED_BS], volume_batch[LABELED_BS:])
        loss_ce, loss_dice = L.seg_ce_dice(outputs[0], label_lab)
        loss_aux = self.aux_loss(outputs[2], label_lab)
        loss_pse = self.pse_loss(outputs[3], outputs[1])
        loss_cons = L.softmax_mse_loss(outputs[3], outputs[4])
        loss = w[0] * loss_dice + w[1] * loss_ce + w[2] * loss_aux + w[3] * loss_pse + w[4] * loss_cons
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if self.dp is not None:
            self.dp.average()
        opt.step()
        return loss

    def eager_step_with_lr(self, x, y):
        loss = self.eager_step(x, y)
        lr_ = self.lr_now()
        for g in self.opt.param_groups:
            g["lr"] = lr_
        self.it += 1
        return loss

    def capture(self):
        from icl_b200.graph import GraphedStep
        for _ in range(2):
            self.eager_step_with_lr(self.x_dev, self.y_dev)
        lc0 = self._lib.launch_count()
        # N > 1: the step's NCCL collectives (bucketed gradient all-reduce, factor all-gathers) are captured with it
        self.graphed = GraphedStep(self.eager_step, (self.x_dev, self.y_dev), self.opt, warmup=1)
        # our kernel nodes in the captured graph = C-ABI launches issued by the (1 warm-up + 1 captured) step executions
        self.graph_launches = (self._lib.launch_count() - lc0) // 2

    def step(self, x, y):
        if self.graphed is None:
            return self.eager_step_with_lr(x, y)
        lr_ = self.lr_now()   # the captured step cannot update python state: do it here
        for g in self.opt.param_groups:
            g["lr"] = lr_
        self.it += 1
        return self.graphed(x, y)

    def close(self):
        if self.graphed is not None:
            self.graphed.release()
        self.graphed = None
        if self.dp is not None:
            self.dp.close()


def measure_local_reference(a, wl, rank, world, dev):
    """N > 1: the SAME workload with the exchange switched off (every rank steps on its own batch, nothing crosses NVLink), timed
    like the real run (barrier, CUDA events, max over ranks): the single-GPU denominator of this line's weak-scaling efficiency,
    measured on the same GPUs in the same process."""
    tb = TrainBench(a, wl, rank, world, dev, exchange=False)
    timer = Timer(world, dev)
    if not a.no_graph:
        tb.capture()
    for _ in range(max(a.warmup, 3)):
        tb.step(tb.x_dev, tb.y_dev)
    ms = timer(lambda: tb.step(tb.x_dev, tb.y_dev), a.steps)
    tb.close()
    return ms


def measure_train(a, wl, rank, world, dev, local, full=True):
    """Returns the JSON line (dict) for one training workload; `full` = with e2e, kernel profile and roofline."""
    import torch
    from icl_b200 import ops
    local_ms = None
    if world > 1 and full:
        local_ms = measure_local_reference(a, wl, rank, world, dev)
        torch.cuda.empty_cache()
    tb = TrainBench(a, wl, rank, world, dev)
    timer = Timer(world, dev)
    use_graph = not a.no_graph
    if use_graph:
        tb.capture()
    for _ in range(max(a.warmup, 3)):
        tb.step(tb.x_dev, tb.y_dev)
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = tb._lib.launch_count()
    ms = timer(lambda: tb.step(tb.x_dev, tb.y_dev), a.steps)
    host_enqueue_ms = timer.host_ms
    launches = tb._lib.launch_count() - l0
    if use_graph:
        launches = tb.graph_launches * a.steps   # replays do not pass through the C-ABI launch counter
    clocks = sampler.stop() if sampler else None
    vox = BATCH * VOX * world
    line = {
        "metric": "3D U-Net ICL train voxels/s (96^3)", "value": vox / (ms * 1e-3), "unit": "voxels/s", "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16x3+fp32" if a.precision == "parity" else "bf16+fp32", "data": "synthetic",
        "config": {"workload": tb.cfg["desc"], "global_batch": BATCH * world, "parallelism": "dp%d" % world, "precision_mode": a.precision,
                   "launch": "cuda-graph replay of the whole step" if use_graph else "eager",
                   "streams": "ICL heads on 8 side streams next to the upper decoder levels (same kernels, same values)"
                              if os.environ.get("ICL_HEAD_LANES", "1") != "0" else "one stream",
                   "exchange": None if world == 1 else ("bucketed all-reduce + mlp2 factor all-gathers on a side stream, overlapped with backward"
                                                        if not a.no_overlap else "after backward"),
                   "l2": "no flush needed: per-step working set (785M params + activations, >15 GB) >> 126 MB L2"},
        "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms, "clocks": clocks,
    }
    if local_ms is not None:
        line["weak_scaling_reference"] = {
            "workload": wl, "ms_per_step_without_exchange": local_ms, "value_1gpu": BATCH * VOX / (local_ms * 1e-3),
            "note": "same step, same GPUs, exchange off (each rank on its own batch), max over ranks: the single-GPU denominator for this "
                    "line; a default N = 1 run reports cfg2 as `value` (BASELINE configs[1]) and this workload under `also`"}
    if not full:
        tb.close()
        return line

    # ---- end-to-end arm: pinned host buffers in, loss scalar out, every step; H2D of the next step on a copy stream
    copy_stream = torch.cuda.Stream()
    stage = [(torch.empty_like(tb.x_dev), torch.empty_like(tb.y_dev)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    state = {"i": 0}

    def h2d(slot):
        with torch.cuda.stream(copy_stream):
            stage[slot][0].copy_(tb.x_host, non_blocking=True)
            stage[slot][1].copy_(tb.y_host, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_step():
        i = state["i"]
        slot = i % 2
        torch.cuda.current_stream().wait_event(ready[slot])
        loss = tb.step(stage[slot][0], stage[slot][1])
        h2d(1 - slot)                                           # next step's inputs travel while this step computes
        loss_host.copy_(loss.detach(), non_blocking=False)      # the step's result back on the host (blocks until the step is done)
        state["i"] = i + 1

    h2d(0)
    e2e_step()
    ms_e2e = timer(e2e_step, a.steps)
    line["e2e"] = {"value": vox / (ms_e2e * 1e-3), "unit": "voxels/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": tb.x_host.numel() * 4 + tb.y_host.numel() * 8, "d2h_bytes_per_step": 4,
                   "note": "H2D of step i+1 issued on a copy stream during step i; D2H of the loss blocks the host every step"}

    # ---- where the REPLAYED step spends its time: CUPTI kernel records of `steps` more replays (outside the timed region).  The eager
    # event pass below times every launch with the host in the loop, which inflates kernels of a few microseconds; the kernel that
    # dominates the step is therefore chosen from these records, its roofline from the CUDA-event duration of the same kernel.
    # The timed step runs the ICL heads on side streams next to the backbone (icl_b200/lanes.py); kernels that share the GPU stretch
    # each other, so their durations inside that step say nothing about the kernels themselves.  Every per-kernel number below is
    # therefore taken from the SAME step issued on ONE stream (ICL_HEAD_LANES=0: same kernels, same inputs, serial order), which
    # is also timed (`single_stream`).
    graph_kernels = None
    lanes_prev = os.environ.get("ICL_HEAD_LANES")
    if not a.no_profile:
        os.environ["ICL_HEAD_LANES"] = "0"
        if use_graph:
            tb.graphed.release()
            tb.graphed = None
            tb.capture()
            for _ in range(3):
                tb.step(tb.x_dev, tb.y_dev)
            line["single_stream"] = {"ms_per_step": timer(lambda: tb.step(tb.x_dev, tb.y_dev), a.steps),
                                     "note": "the same step with the ICL-head lanes off (every kernel on one stream); the per-kernel "
                                             "breakdowns (kernels, graph_kernels, conv_tensor_util, hbm_kernels, roofline) are measured on it"}
        if rank == 0:
            graph_kernels = cupti_kernel_times(lambda: tb.step(tb.x_dev, tb.y_dev), min(a.steps, 5))
        else:
            for _ in range(min(a.steps, 5)):   # the other ranks step along (collectives)
                tb.step(tb.x_dev, tb.y_dev)

    # ---- per-kernel breakdown (CUDA events around every launch of our kernels), separate eager pass in this same run
    prof = None
    if not a.no_profile:
        nprof = 2
        if rank == 0:
            ops.profile_start()
        for _ in range(nprof):   # every rank runs the profiled steps (they contain collectives); only rank 0 records events
            tb.eager_step_with_lr(tb.x_dev, tb.y_dev)
        torch.cuda.synchronize()
        if rank == 0:
            prof = ops.profile_stop(nprof)
    if lanes_prev is None:
        os.environ.pop("ICL_HEAD_LANES", None)
    else:
        os.environ["ICL_HEAD_LANES"] = lanes_prev
    if rank == 0 and prof is not None and a.detail:
        os.makedirs(os.path.dirname(os.path.abspath(a.detail)), exist_ok=True)
        json.dump(prof["detail"], open(a.detail, "w"), indent=1)
    tb.close()
    if rank == 0 and prof:
        pk = peaks()
        gf = tb.cfg["conv_gflop"]
        line["kernels"] = prof["kernels"][:14]
        conv_ms = sum(k["ms_per_step"] for k in prof["kernels"] if k["name"].startswith("icl_conv3d"))
        line["conv_tensor_util"] = {
            "algorithmic_gflop_per_step": gf, "conv_kernel_ms_per_step": conv_ms,
            "achieved_tflops": gf / conv_ms if conv_ms else None,
            "frac_of_sustained_peak": (gf / conv_ms) / pk["tf_sus"] if conv_ms else None,
            "mma_issue_frac_of_sustained_peak": ((3.0 if a.precision == "parity" else 1.0) * gf / conv_ms) / pk["tf_sus"] if conv_ms else None}
        if graph_kernels:
            # the same sum from the CUPTI records of the replayed step (no host in the loop): every convolution kernel incl. the
            # weight packing and the fixed-order reductions of the weight-gradient partials
            conv_us = sum(k["us_per_step"] for k in graph_kernels if k["kernel"].startswith(CONV_KERNEL_PREFIXES))
            if conv_us > 0:
                mult = 3.0 if a.precision == "parity" else 1.0
                line["conv_tensor_util"].update({
                    "conv_kernel_ms_per_step_in_graph": conv_us * 1e-3, "achieved_tflops_in_graph": gf / (conv_us * 1e-3),
                    "frac_of_sustained_peak_in_graph": gf / (conv_us * 1e-3) / pk["tf_sus"],
                    "mma_issue_frac_of_sustained_peak_in_graph": mult * gf / (conv_us * 1e-3) / pk["tf_sus"]})
        # HBM-bound kernels the north star names (attention / loss / mlp2 / optimizer): achieved GB/s of algorithmic bytes vs measured peak
        line["hbm_kernels"] = [
            {"name": k["name"], "gbs": k["gbs"], "frac_of_hbm_peak": k["gbs"] / pk["hbm"], "mbytes_per_launch": k["mbytes_per_launch"],
             "ms_per_step": k["ms_per_step"]}
            for k in prof["kernels"] if k["gbs"] and not k["name"].startswith("icl_conv3d")][:16]
        top = prof["kernels"][0]
        if graph_kernels:
            line["graph_kernels"] = graph_kernels[:12]
            by_name = {k["name"]: k for k in prof["kernels"]}
            for gk in graph_kernels:          # largest in-graph share whose C-ABI entry point the event pass timed
                ent = by_name.get(KERNEL_ENTRY.get(gk["kernel"], ""))
                if ent:
                    top = dict(ent, share=gk["share"], graph_kernel=gk["kernel"], graph_us_per_launch=gk["us_per_launch"])
                    break
        line["roofline"] = roofline_of(top, pk, 3.0 if a.precision == "parity" else 1.0)
    return line


CONV_KERNEL_PREFIXES = ("conv3d_", "wgrad_ts_reduce_k", "wgrad_reduce_k", "pack_w_umma_k", "pack_w_walk_k", "repack_w")

# CUDA kernel (CUPTI name) -> the C-ABI entry point that launches it (the names the CUDA-event pass records)
KERNEL_ENTRY = {
    "sgd_factored_umma_k": "icl_sgd_factored_apply", "conv3d_umma_walk_k": "icl_conv3d_umma_walk_fwd", "conv3d_umma_k": "icl_conv3d_umma_fwd",
    "conv3d_wgrad_ts_k": "icl_conv3d_wgrad_ts", "conv3d_wgrad_umma_k": "icl_conv3d_wgrad_umma", "sgemm_k": "icl_sgemm",
    "instnorm_relu_fwd_k": "icl_instnorm_relu_fwd", "instnorm_relu_bwd_apply_cl_k": "icl_instnorm_relu_bwd",
    "instnorm_relu_bwd_reduce_k": "icl_instnorm_relu_bwd", "bigw_gemm_k<0>": "icl_bigw_linear_fwd", "bigw_gemm_k<1>": "icl_bigw_linear_dgrad",
    "upsample2x_bwd_v4_k": "icl_upsample2x_bwd", "upsample2x_fwd_cl_k": "icl_upsample2x_fwd", "layernorm_fwd_k": "icl_layernorm_fwd",
    "sgd_multi_k": "icl_sgd_multi", "dropout_k": "icl_dropout",
}


def cupti_kernel_times(step, n):
    """Device time per kernel over n calls of step() from CUPTI activity records (torch.profiler): [{kernel, launches_per_step,
    us_per_step, us_per_launch, share}] sorted by time, or None if the profiler is unavailable."""
    import collections
    import torch
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(n):
                step()
            torch.cuda.synchronize()
        agg = collections.defaultdict(lambda: [0.0, 0])
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                name = ev.name.replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "")
                agg[name][0] += ev.device_time
                agg[name][1] += 1
    except Exception as e:   # noqa: BLE001 — measurement aid only
        sys.stderr.write("bench: CUPTI kernel records unavailable (%r)\n" % (e,))
        return None
    tot = sum(v[0] for v in agg.values()) or 1.0
    return [{"kernel": k, "launches_per_step": c / n, "us_per_step": us / n, "us_per_launch": us / c, "share": us / tot}
            for k, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])]


def roofline_of(top, pk, mma_mult=3.0):
    # dominant kernel = largest share of the step's device time (CUDA events around every launch of our kernels in an eager pass of
    # the same step, on the launching stream); achieved = algorithmic FLOPs or bytes / that time
    # a kernel with both an algorithmic byte and FLOP count (the fused rank-R mlp2 update: HBM-bound for R <= 256, tensor-bound for
    # the R >= 512 of data-parallel runs) is measured against the resource it uses the larger fraction of (MMAs ISSUED: x mma_mult)
    tf = top["gflop_per_launch"] / top["ms_per_launch"] if top.get("gflop_per_launch") else 0.0
    gb = top["mbytes_per_launch"] / top["ms_per_launch"] if top.get("mbytes_per_launch") else 0.0
    bound = "tensor" if top["name"].startswith("icl_conv3d") or (mma_mult * tf / pk["tf_sus"] > gb / pk["hbm"]) else "hbm"
    if bound == "tensor":
        ach, peak, unit = tf, pk["tf_sus"], "TFLOP/s"
    else:
        ach, peak, unit = gb, pk["hbm"], "GB/s"
    traffic = None
    # (the DRAM capture of the fused update is the R = 32 one: not attached when the kernel runs tensor-bound at another R)
    for tname in (("traffic_r02.json", "traffic_r01.json") if bound == "hbm" or top["name"].startswith("icl_conv3d") else ()):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            ent = json.load(open(tpath)).get(top["name"])
            if ent:
                traffic = {"bytes_per_launch": ent["traffic_bytes"], "algorithmic_bytes": ent["algorithmic_bytes"], "shape": ent["shape"],
                           "source": ent["source"]}
                break
    out = {"kernel": top["name"], "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak if peak else None,
           "traffic": traffic, "peak_source": pk["src"] + " (sustained)", "share_of_step": top["share"],
           "launches_per_step": top["launches_per_step"], "ms_per_launch_cuda_events": top["ms_per_launch"]}
    if bound == "tensor":
        out["mma_issue_frac"] = mma_mult * ach / peak if peak else None
        out["note"] = "achieved = algorithmic TFLOP/s; the parity mode issues %g MMAs per algorithmic MMA (mma_issue_frac)" % mma_mult
    if "graph_kernel" in top:
        out["selected_by"] = "largest share of the replayed step's kernel time (CUPTI records): %s, %.1f us per launch in the graph" % (
            top["graph_kernel"], top["graph_us_per_launch"])
    return out


def measure_infer(a, rank, world, dev, local):
    """config 5: sliding-window inference of one BraTS-sized volume, windows sharded round-robin over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from icl_b200 import _lib, inference, ops
    from icl_b200.networks.unet_3D import unet_3D
    from icl_b200.utils import synth
    K = INFER["K"]
    net = unet_3D(feature_scale=4, n_classes=K, in_channels=1)
    synth.load_synth(net, 77)
    net.to(dev).eval()
    image = synth.synth_volume(INFER["shape"], 78)
    img_dev = image.to(dev)
    image_np = image.numpy()
    st, patch, wb = INFER["stride"], INFER["patch"], a.window_batch
    timer = Timer(world, dev)

    def resident():
        score, cnt, _ = inference.sliding_window_scores(net, img_dev, st, st, patch, K, rank=rank, world_size=world, window_batch=wb)
        if world > 1:
            dist.all_reduce(score)
            dist.all_reduce(cnt)
        return inference._finalize(score, cnt)

    def e2e():
        return inference.test_single_case_sharded(net, image_np, st, st, patch, num_classes=K, window_batch=wb)

    for _ in range(max(a.warmup, 3)):
        resident()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = _lib.launch_count()
    ms = timer(resident, a.steps)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    e2e()
    ms_e2e = timer(e2e, a.steps)
    vox = float(np.prod(INFER["shape"]))
    sx, sy, sz = (len(inference.window_starts(max(s, p), p, st)) for s, p in zip(INFER["shape"], patch))
    nwin = sx * sy * sz
    line = {
        "metric": "3D U-Net sliding-window inference voxels/s", "value": vox / (ms * 1e-3), "unit": "voxels/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16x3+fp32" if a.precision == "parity" else "bf16+fp32", "data": "synthetic",
        "config": {"workload": INFER["desc"], "windows": nwin, "window_batch": wb, "parallelism": "windows round-robin over %d rank(s), one "
                   "all-reduce of score map + counts" % world, "precision_mode": a.precision,
                   "l2": "volume + score map + per-window activations (>1 GB per window batch) >> 126 MB L2"},
        "gpu_launches": int(launches), "clocks": clocks,
        "e2e": {"value": vox / (ms_e2e * 1e-3), "unit": "voxels/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(vox) * 4,
                "d2h_bytes_per_step": int(vox), "note": "numpy volume in, numpy int64 label map out (test_single_case contract); the label map "
                                                         "crosses PCIe as one byte per voxel and is widened to int64 on the host"},
    }
    if not a.no_profile:
        if rank == 0:
            ops.profile_start()
        resident()
        torch.cuda.synchronize()
        if rank == 0:
            prof = ops.profile_stop(1)
            pk = peaks()
            gf = INFER["conv_gflop_per_window"] * len(range(rank, nwin, world))
            conv_ms = sum(k["ms_per_step"] for k in prof["kernels"] if k["name"].startswith("icl_conv3d"))
            line["kernels"] = prof["kernels"][:10]
            line["conv_tensor_util"] = {"algorithmic_gflop_per_step_rank0": gf, "conv_kernel_ms_per_step": conv_ms,
                                        "achieved_tflops": gf / conv_ms if conv_ms else None,
                                        "frac_of_sustained_peak": (gf / conv_ms) / pk["tf_sus"] if conv_ms else None,
                                        "mma_issue_frac_of_sustained_peak": ((3.0 if a.precision == "parity" else 1.0) * gf / conv_ms) / pk["tf_sus"]
                                        if conv_ms else None}
            line["roofline"] = roofline_of(prof["kernels"][0], pk)
    return line


def cpu_baseline(wl):
    """Rank 0, N = 1: the reference's CPU implementation of the same workload on the box's host cores, bounded sample."""
    import torch
    threads = os.cpu_count() or 1
    if wl == "infer":
        with torch.no_grad():
            step, kind = ref_infer_factory("cpu", threads)
            t0 = time.perf_counter()
            step()
            t = time.perf_counter() - t0
        units = float(INFER["shape"][0] * INFER["shape"][1] * INFER["shape"][2])
        sample = "one full volume (48 windows, one window per net call as test_3D_BraTS.py:113-133), %.1f s" % t
    else:
        if reference_available():
            step, kind = ref_train_step_factory(wl, "cpu", threads), "reference"
        else:
            step, kind = port_train_step_factory(wl, threads), "port"
        step()
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            step()
            ts.append(time.perf_counter() - t0)
        t = statistics.median(ts)
        units = float(BATCH * VOX)
        sample = "1 warm-up + 2 timed full %s steps (4 x 96^3 voxels each), median; %.2f s/step" % (wl, t)
    return {"value": units / t, "unit": "voxels/s", "cores": threads, "kind": kind, "sample": sample}


def main_gpu(a):
    import torch
    import torch.distributed as dist
    import icl_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    icl_b200.set_precision(a.precision)
    wl = default_workload(a, world)
    if wl == "infer":
        line = measure_infer(a, rank, world, dev, local)
    else:
        line = measure_train(a, wl, rank, world, dev, local)
        if world == 1 and not a.workload and not a.no_also:
            # the N > 1 lines run cfg3: give them a single-GPU denominator measured in this same run
            torch.cuda.empty_cache()
            sub = measure_train(a, "cfg3", rank, world, dev, local, full=False)
            line["also"] = {"cfg3": {k: sub[k] for k in ("value", "unit", "ms_per_step", "gpu_launches", "config")}}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        torch.cuda.empty_cache()
        line["cpu_baseline"] = cpu_baseline(wl)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl != "icl_b200":
        main_reference(args)
    else:
        main_gpu(args)

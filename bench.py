#!/usr/bin/env python
"""Benchmark of the ICL hot path (BASELINE.json metric: 3D U-Net ICL train voxels/s at 96^3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision parity|fast]

Workload (config 2 of BASELINE.json): unet_3D_icl(n_classes=2, in_channels=1), per rank 2 labeled + 2 unlabeled
synthetic 1x96^3 patches, one step = forward (two backbone passes + SSPA/USCL heads) + the five ICL losses +
backward + momentum-SGD step with poly LR  (reference loop: train_inherent_consistent_unet_3D_BraTS.py:103-119).
N > 1: weak scaling, one process per GPU (torchrun), rank-local batches, gradients averaged (icl_b200.parallel).

Prints ONE JSON line (rank 0).  `value` = whole-job voxels/s with inputs resident in HBM; `e2e` = the same step
driven from pinned HOST buffers (H2D of volume+labels and D2H of the loss inside the timed region).
`--impl reference` times the reference's CPU implementation of the same step (the oracle port of it — the
reference tree itself does not travel to the GPU box) on all host cores.
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
K_CLASSES = 2
BASE_LR, MAX_ITERS = 0.01, 30000
# algorithmic conv FLOPs per step: fwd 4 samples + pruned backward
CONV_GFLOP_PER_STEP = 1166.9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="icl_b200")
    ap.add_argument("--precision", default="parity", choices=["parity", "fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--unfused-optimizer", action="store_true", help="materialise the mlp2 weight gradients (reference-style .grad) instead of "
                    "applying them as rank-R updates inside the optimizer")
    ap.add_argument("--no-graph", action="store_true", help="eager step (one Python-enqueued launch per kernel) instead of CUDA-graph replay")
    ap.add_argument("--detail", default="", help="write the per-(kernel, shape) CUDA-event breakdown to this JSON file")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_step_factory(threads):
    """The reference's CPU path for this step, as restated in oracle/restate.py (kind = "port")."""
    import torch
    from collections import OrderedDict
    from oracle import restate as R
    from oracle import synth
    import json as _json
    torch.set_num_threads(threads)
    keys = _json.load(open(os.path.join(ROOT, "tests", "golden", "state_keys.json")))["unet_3D_icl_k2"]
    shapes = OrderedDict((k, tuple(s)) for k, s in keys)
    P = R.make_params(synth.synth_state_dict(shapes, 1337))
    x = synth.synth_volume((BATCH, 1, 96, 96, 96), 1338)
    y = synth.synth_labels((BATCH, 96, 96, 96), K_CLASSES, 1339)
    bufs = {}
    it = [0]

    def step():
        rand = R.TorchRand()
        L, grads, _ = R.train_step_3d(P, x, y, LABELED_BS, K_CLASSES, rand=rand)
        lr = BASE_LR if it[0] == 0 else R.poly_lr(BASE_LR, it[0] - 1, MAX_ITERS)
        R.sgd_step(P, grads, bufs, lr)
        it[0] += 1
        return float(L["total"])
    return step


def run_cpu(steps, warmup):
    import torch
    threads = os.cpu_count() or 1
    step = cpu_step_factory(threads)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts), threads, torch.get_num_threads()


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(a.steps, 3)), min(a.warmup, 1)
    t, threads, _ = run_cpu(steps, warmup)
    v = BATCH * VOX / t
    sample = "%d warm-up + %d timed full config-2 steps (4 x 96^3 voxels each), median" % (warmup, steps)
    line = {
        "impl": "reference", "metric": "3D U-Net ICL train voxels/s (96^3)", "value": v, "unit": "voxels/s", "n_gpus": a.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "config2: unet_3D_icl(K=2,in=1) ICL train step, batch 4 (2 lab + 2 unlab) x 1x96^3, CPU host cores"},
        "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
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


def main_gpu(a):
    import torch
    import torch.distributed as dist
    import icl_b200
    from icl_b200 import _lib, ops, parallel
    from icl_b200.networks.unet_3D_icl import unet_3D_icl
    from icl_b200.optim import SGD
    from icl_b200.utils import losses as L
    from oracle import synth  # deterministic synthetic parameters / inputs only (numpy RNG); no oracle compute here

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

    net = unet_3D_icl(feature_scale=4, n_classes=K_CLASSES, in_channels=1)
    synth.load_synth(net, 1337)
    net.to(dev).train()
    # fused_factored: the 13 824^2 mlp2 weight gradients are applied as rank-32 updates inside the optimizer
    opt = SGD(net.parameters(), lr=BASE_LR, momentum=0.9, weight_decay=1e-4, fused_factored=not a.unfused_optimizer)
    dp = parallel.GradAverager(net, world, factored=a.unfused_optimizer) if world > 1 else None
    ce_loss, dice_loss = L.CrossEntropyLoss(), L.DiceLoss(K_CLASSES)
    aux_loss, pse_loss = L.AuxLoss3D(K_CLASSES), L.PseudoSoftLoss3D(K_CLASSES)

    x_host = synth.synth_volume((BATCH, 1, 96, 96, 96), 1338 + rank).pin_memory()
    y_host = synth.synth_labels((BATCH, 96, 96, 96), K_CLASSES, 1339 + rank).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    it = [0]

    phases = os.environ.get("ICL_BENCH_PHASES") == "1"   # debug: CUDA-event time of forward / losses / backward / optimizer
    ph_ev = []

    def mark():
        if phases:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ph_ev.append(e)

    def step(volume_batch, label_batch):
        mark()
        outputs = net(volume_batch[:LABELED_BS], volume_batch[LABELED_BS:])
        mark()
        loss_ce, loss_dice = L.seg_ce_dice(outputs[0], label_batch[:LABELED_BS])
        loss_aux = aux_loss(outputs[2], label_batch[:LABELED_BS])
        loss_pse = pse_loss(outputs[3], outputs[1])
        loss_cons = L.softmax_mse_loss(outputs[3], outputs[4])
        loss = loss_dice + loss_ce + loss_aux + loss_pse + 10 * loss_cons
        opt.zero_grad(set_to_none=True)
        mark()
        loss.backward()
        mark()
        if dp is not None:
            dp.average()
        opt.step()
        lr_ = BASE_LR * (1.0 - it[0] / MAX_ITERS) ** 0.9
        for g in opt.param_groups:
            g["lr"] = lr_
        it[0] += 1
        mark()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        host_ms[0] = (time.perf_counter() - t0) * 1e3 / n   # host time to ENQUEUE one step (no sync inside)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / n

    # ---- CUDA-graph replay of the whole step (single GPU; the eager step is host-bound)
    use_graph = not a.no_graph
    eager_step = step
    if use_graph:
        from icl_b200.graph import GraphedStep
        for _ in range(2):
            eager_step(x_dev, y_dev)
        lc0 = _lib.launch_count()
        # N > 1: the step's NCCL collectives (gradient all-reduce, factor all-gather) are captured with it
        graphed = GraphedStep(eager_step, (x_dev, y_dev), opt, warmup=1)
        # our kernel nodes in the captured graph = C-ABI launches issued by the (1 warm-up + 1 captured) step executions
        graph_launches_per_step = (_lib.launch_count() - lc0) // 2

        def step(volume_batch, label_batch):  # noqa: F811
            lr_ = BASE_LR * (1.0 - it[0] / MAX_ITERS) ** 0.9   # the captured step cannot update python state: do it here
            for g in opt.param_groups:
                g["lr"] = lr_
            it[0] += 1
            return graphed(volume_batch, label_batch)

    # ---- device-resident arm
    for _ in range(max(a.warmup, 3)):
        step(x_dev, y_dev)
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = _lib.launch_count()
    del ph_ev[:]
    ms = timed(lambda: step(x_dev, y_dev), a.steps)
    host_enqueue_ms = host_ms[0]
    launches = (_lib.launch_count() - l0)
    if use_graph:
        launches = graph_launches_per_step * a.steps   # replays do not pass through the C-ABI launch counter
    if phases and rank == 0:
        names = ["forward", "losses", "backward", "optimizer"]
        tot = [0.0] * 4
        for i in range(0, len(ph_ev) - 4, 5):
            for j in range(4):
                tot[j] += ph_ev[i + j].elapsed_time(ph_ev[i + j + 1])
        n = max(1, len(ph_ev) // 5)
        sys.stderr.write("[phases] " + "  ".join("%s %.2f ms" % (nm, t / n) for nm, t in zip(names, tot)) + "\n")
    clocks = sampler.stop() if sampler else None

    # ---- end-to-end arm: pinned host buffers in, loss scalar out, every step
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def e2e_step():
        xb = x_host.to(dev, non_blocking=True)
        yb = y_host.to(dev, non_blocking=True)
        loss = step(xb, yb)
        loss_host.copy_(loss.detach(), non_blocking=False)

    e2e_step()
    ms_e2e = timed(e2e_step, a.steps)

    # ---- per-kernel breakdown (CUDA events around every launch of our kernels), separate pass in this same run
    prof = None
    if not a.no_profile:
        # every rank runs the profiled steps (they contain collectives); only rank 0 records events
        nprof = 2
        if rank == 0:
            ops.profile_start()
        for _ in range(nprof):
            eager_step(x_dev, y_dev)
        torch.cuda.synchronize()
        if rank == 0:
            prof = ops.profile_stop(nprof)
    if rank == 0 and prof is not None and a.detail:
        if True:
            os.makedirs(os.path.dirname(os.path.abspath(a.detail)), exist_ok=True)
            json.dump(prof["detail"], open(a.detail, "w"), indent=1)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        del net, opt
        torch.cuda.empty_cache()
        t, threads, _ = run_cpu(2, 1)
        cpu = {"value": BATCH * VOX / t, "unit": "voxels/s", "cores": threads, "kind": "port",
               "sample": "1 warm-up + 2 timed full config-2 steps (4 x 96^3 voxels each) of the oracle port, median; %.2f s/step" % t}

    if rank == 0:
        pk = peaks()
        vox = BATCH * VOX * world
        line = {
            "metric": "3D U-Net ICL train voxels/s (96^3)", "value": vox / (ms * 1e-3), "unit": "voxels/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16x3+fp32" if a.precision == "parity" else "bf16+fp32", "data": "synthetic",
            "config": {"workload": "config2: unet_3D_icl(K=2,in=1) ICL train step (fwd + 5 losses + bwd + SGD), per-rank batch 4 "
                                   "(2 lab + 2 unlab) x 1x96^3", "global_batch": BATCH * world, "parallelism": "dp%d" % world,
                       "precision_mode": a.precision, "launch": "cuda-graph replay of the whole step" if use_graph else "eager",
                       "l2": "no flush needed: per-step working set (785M params + activations, >15 GB) >> 126 MB L2"},
            "e2e": {"value": vox / (ms_e2e * 1e-3), "unit": "voxels/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms,
            "clocks": clocks,
            "conv_tensor_util": {"algorithmic_gflop_per_step": CONV_GFLOP_PER_STEP},
        }
        if prof:
            # dominant kernel = largest share of the step's device time (CUDA events around every launch of our kernels in an
            # eager pass of the same step, on the launching stream); achieved = algorithmic FLOPs or bytes / that time
            top = prof["kernels"][0]
            line["kernels"] = prof["kernels"][:12]
            conv_ms = sum(k["ms_per_step"] for k in prof["kernels"] if k["name"].startswith("icl_conv3d"))
            line["conv_tensor_util"].update({
                "conv_kernel_ms_per_step": conv_ms,
                "achieved_tflops": CONV_GFLOP_PER_STEP / conv_ms if conv_ms else None,
                "frac_of_sustained_peak": (CONV_GFLOP_PER_STEP / conv_ms) / pk["tf_sus"] if conv_ms else None,
                "mma_issue_frac_of_sustained_peak": ((3.0 if a.precision == "parity" else 1.0) * CONV_GFLOP_PER_STEP / conv_ms) / pk["tf_sus"]
                if conv_ms else None})
            bound = "tensor" if top["name"].startswith("icl_conv3d") else "hbm"
            if bound == "tensor":
                ach, peak, unit = top["gflop_per_launch"] / top["ms_per_launch"], pk["tf_sus"], "TFLOP/s"
            else:
                ach, peak, unit = top["mbytes_per_launch"] / top["ms_per_launch"], pk["hbm"], "GB/s"
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
            if os.path.exists(tpath):
                ent = json.load(open(tpath)).get(top["name"])
                if ent:
                    traffic = {"bytes_per_launch": ent["traffic_bytes"], "algorithmic_bytes": ent["algorithmic_bytes"], "shape": ent["shape"],
                               "source": ent["source"]}
            line["roofline"] = {"kernel": top["name"], "bound": bound, "achieved": ach, "peak": peak, "unit": unit,
                                "frac": ach / peak if peak else None, "traffic": traffic, "peak_source": pk["src"] + " (sustained)",
                                "share_of_step": top["share"], "launches_per_step": top["launches_per_step"]}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        # destroy_process_group() hangs while a captured CUDA graph still holds NCCL work; the JSON line is out, so leave
        # without the collective teardown
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_gpu(args)

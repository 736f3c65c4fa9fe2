#!/usr/bin/env python
"""PCAA hot-path benchmark (BASELINE.json metric: PCAA train samples/s on B200; inference seq/s; % of roofline).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N ...            # the reference's own train_variant4 loop (baseline/_ref) on the host CPU cores
    python bench.py --impl reference --device cuda ...       # the same unmodified loop on cuda:0, stock torch eager (gpu_eager_baseline)

One "step" = one full variant-4 AAE iteration (encoder fwd, WGAN-GP critic step, decoder + Chamfer + adversarial +
CE generator step, both Adam updates; reference PCAA_ablation.py:882-1021) on one batch of synthetic
mmGait10-shaped crops.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NMAX, NCLS = 150, 4
METRIC, UNIT = "pcaa_train_samples_per_sec", "samples/s"
INFER_METRIC, INFER_UNIT = "pcaa_openset_inference_seq_per_sec", "seq/s"
FLOP_PER_POINT_FWD = 2 * (512 * 512 + 512 * 1024 + 1024 * 1024)          # PointNet layers 2-4, forward only
# algorithmic tensor-core work of one train step, per sample (SURVEY.md 8d / DESIGN.md): PointNet layers 2-4,
# forward + data gradient + weight gradient = 3 x 2 x (512*512 + 512*1024 + 1024*1024) MAC-FLOPs per point,
# minus the layer-2 data gradient... (layer 2 HAS a data gradient towards layer 1's BatchNorm) -> 3 GEMMs per layer.
FLOP_PER_POINT_TC = 3 * 2 * (512 * 512 + 512 * 1024 + 1024 * 1024)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md: 1.4 PFLOP/s sustained under the power cap -- burst 1.59 --, 6.65 TB/s)"


def measured_traffic(kind: str, batch: int, nmax: int):
    """DRAM bytes per launch of the dominant kernel family from the committed `ncu --set full` capture
    (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum averaged over the PointNet GEMM launches of one
    step at the profiled batch); None when no capture exists for this configuration."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        d = json.load(f)
    e = d.get(f"{kind}_B{batch}_N{nmax}")
    return None if e is None else e.get("bytes_per_launch")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])), mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        os.unlink(self.f.name)
        return out


# ---------------------------------------------------------------------------------------------- reference / CPU arm
def cpu_step_rate(batch: int, steps: int, warmup: int, threads: int, NMAX: int = NMAX):
    """The reference's algorithm on the host cores (oracle port of PCAA_ablation.py:882-1021), fp32, `threads` threads."""
    from oracle import pcaa_oracle as O
    torch.set_num_threads(threads)
    p = O.det_params(NCLS, NMAX, seed=0)
    means = O.sample_distant_points(32, NCLS, 10, 10).float()
    cfg = dict(LR=1e-4, B1=0.9, B2=0.99, GP_WEIGHT=15, ADV_WEIGHT=1, NMAX=NMAX)
    pcs, gt = O.synth_batch(batch, NMAX, NCLS, seed=1234)
    rng = np.random.default_rng(0)
    ost = {}
    times = []
    for i in range(warmup + steps):
        z0 = torch.from_numpy(rng.normal(0, 1, (batch, 32))).float()
        al = torch.from_numpy(rng.uniform(0, 1, (batch, 1)).astype(np.float32))
        t0 = time.perf_counter()
        O.train_step_variant4(p, ost, pcs, gt, z0, al, means, cfg, clone_leaves=False)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return batch / dt, dt


def reference_train_rate(device: str, batch: int, nmax: int, steps: int, warmup: int, budget_s: float):
    """The reference's train loop on `device`.  Preferred: the UNMODIFIED PCAA_ablation.train_variant4 from baseline/_ref on an
    in-memory synthetic dataset (baseline/ref_loop.py; kind "reference").  Without baseline/_ref: the oracle port on the CPU
    (kind "port").  Returns dict(rate, dt, steps, warmup, kind, note)."""
    from baseline import refenv
    if refenv.available():
        from baseline import ref_loop
        r = ref_loop.time_train_variant4(device, batch, nmax, NCLS, warmup, steps, budget_s=budget_s)
        dt = sum(r["times"]) / len(r["times"])
        return {"rate": batch / dt, "dt": dt, "steps": len(r["times"]), "warmup": r["warmup"], "kind": "reference",
                "note": "the reference's own PCAA_ablation.train_variant4 loop, unmodified (baseline/_ref), stock torch eager fp32 on "
                        + device + "; in-memory synthetic crops"}
    if device != "cpu":
        raise RuntimeError("baseline/_ref is not installed: no reference modules to run on " + device)
    steps = max(1, min(steps, max(1, int(budget_s // 5))))
    rate, dt = cpu_step_rate(batch, steps, warmup, os.cpu_count() or 1, nmax)
    return {"rate": rate, "dt": dt, "steps": steps, "warmup": warmup, "kind": "port",
            "note": "oracle port of PCAA_ablation.py:882-1021 on the host CPU cores (baseline/_ref absent)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    NMAX = args.nmax
    device = args.device or "cpu"
    if args.workload == "infer":
        cb = cpu_infer_rate(NMAX, args.k)
        print(json.dumps({
            "impl": "reference", "metric": INFER_METRIC, "value": cb["value"], "unit": INFER_UNIT, "n_gpus": args.gpus,
            "steps": 1, "warmup": 1, "ms_per_step": cb["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"pcaa_openset_inference_N{NMAX}_C{NCLS}_k{args.k}", "note": cb["sample"]},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": INFER_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    b = args.batch if args.batch_given else (32 if device == "cpu" else 256)
    r = reference_train_rate(device, b, NMAX, args.steps, args.warmup, budget_s=float(os.environ.get("PCAA_REF_BUDGET_S", "200")))
    sample = (f"variant-4 train step, batch {b}, N={NMAX}, C={NCLS}, fp32, {r['warmup']} warm-up + {r['steps']} timed iterations "
              f"({r['dt']:.3f} s/iteration); {r['note']}")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["rate"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
        "warmup": r["warmup"], "ms_per_step": r["dt"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"pcaa_variant4_train_step_N{NMAX}_C{NCLS}", "batch_per_step": b, "device": device,
                   "steps_requested": args.steps, "warmup_requested": args.warmup, "note": r["note"]},
        "cpu_baseline": {"value": r["rate"], "unit": UNIT, "cores": threads if device == "cpu" else 0, "kind": r["kind"],
                         "sample": sample},
        "e2e": {"value": r["rate"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def sub_bench(extra, timeout=900):
    """Run `bench.py <extra>` in a fresh process (the reference's modules register themselves under top-level names such as
    `models` / `utils`; a separate process keeps them away from this one) and return its JSON line, or {"error": ...}."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + extra, capture_output=True, text=True, timeout=timeout,
                           env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
    except subprocess.TimeoutExpired:
        return {"error": f"timeout after {timeout} s"}
    js = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not js:
        return {"error": (r.stderr or r.stdout)[-400:]}
    return json.loads(js[-1])


# ---------------------------------------------------------------------------------------------- sm_100a arm
def measure_infer(args, world, rank, local, dev, B, steps, warmup):
    """Config 5 of BASELINE.json: batch-sharded open-set inference, no collective.  One step = one batch of B crops on every
    rank: eval-mode encoder (BatchNorm folded into the GEMM epilogues), float64 log-likelihood, k-window vote.  Returns the
    result dict on rank 0 (None elsewhere); collective when world > 1 (barriers, one max-reduce of the timings)."""
    from opensetgaitrecognition_pcaa_b200 import _lib, inference, models, ops, utils, synth
    from opensetgaitrecognition_pcaa_b200.loader import DevicePrefetcher
    nmax, k = args.nmax, args.k
    B = max(k, (B // k) * k)
    torch.manual_seed(0)
    enc = models.CGEncoder(n_out_labels=NCLS, use_projection_head=True, nmax_points=nmax).to(dev).float().eval()
    means = utils.sample_distant_points(32, NCLS, 10, 10).float().to(dev)
    nb = 3
    host = [synth.synth_batch(B, nmax, NCLS, seed=4321 + 17 * rank + i)[0].pin_memory() for i in range(nb)]
    devb = [h.to(dev) for h in host]
    lthr = -60.0

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    tc_events = []
    orig_tc, orig_pooled = ops.gemm_tc, ops.gemm_tc_pooled
    record = {"on": False}

    def timed(orig, flops):
        def f(*a, **kw):
            if not record["on"] or (orig is orig_tc and a[2] != _lib.TC_T_AFFINE_ELU):
                return orig(*a, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig(*a, **kw)
            e1.record()
            tc_events.append((e0, e1, flops(a)))
            return r
        return f

    ops.gemm_tc = timed(orig_tc, lambda a: 2.0 * a[3] * a[4] * a[5])                # (a, b, mode, M, N, K)
    ops.gemm_tc_pooled = timed(orig_pooled, lambda a: 2.0 * a[2] * a[3] * a[4])     # (w, aT, M, P, K, n): last layer + mean pool
    try:
        for i in range(warmup):
            inference.sharded_stream_inference(enc, means, devb[i % nb], k, lthr, NCLS, encode_batch=B)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        calls0 = _lib.CALLS
        record["on"] = True
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(steps):
            ll, votes, pred = inference.sharded_stream_inference(enc, means, devb[i % nb], k, lthr, NCLS, encode_batch=B)
        t1.record()
        barrier()
        record["on"] = False
        launches = _lib.CALLS - calls0
        ms = t0.elapsed_time(t1)
        tc_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in tc_events)
        tc_flops = sum(f for _, _, f in tc_events)
        # end to end: pinned host crops -> H2D (copy stream, one batch ahead: loader.DevicePrefetcher) -> encode + score +
        # vote -> D2H of the window labels; the host reads every batch's labels, one batch behind the device
        votes_host = [torch.empty(B // k, dtype=torch.int32).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        # ONE loader for the whole measurement, as in a real stream: its first two batches are untimed warm-up (copy stream,
        # device slots), then K timed steps during which K host->device copies are issued (one batch ahead), one spare batch
        # at the end keeps the last timed step's prefetch identical to the others
        feed = iter(DevicePrefetcher(((host[i % nb],) for i in range(2 + steps + 1)), dev, depth=2))
        for _ in range(2):
            (x,) = next(feed)
            inference.sharded_stream_inference(enc, means, x, k, lthr, NCLS, encode_batch=B)
        barrier()
        host_seen = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            (x,) = next(feed)
            ll, votes, pred = inference.sharded_stream_inference(enc, means, x, k, lthr, NCLS, encode_batch=B)
            j = i & 1
            votes_host[j].copy_(votes, non_blocking=True)
            done[j].record()
            if i > 0:
                done[j ^ 1].synchronize()
                host_seen += int(votes_host[j ^ 1][0])
        done[(steps - 1) & 1].synchronize()
        e1.record()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms_e2e = e0.elapsed_time(e1)
    finally:
        ops.gemm_tc, ops.gemm_tc_pooled = orig_tc, orig_pooled
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return None
    peak_tf, peak_hbm, peak_src = peaks()
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    return {
        "metric": INFER_METRIC, "value": world * B * steps / (ms * 1e-3), "unit": INFER_UNIT, "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"pcaa_openset_inference_N{nmax}_C{NCLS}_k{k}", "batch_per_gpu": B, "global_batch": B * world,
                   "crops_streamed": world * B * steps,
                   "parallelism": f"dp{world} (batch-sharded stream, no collective)",
                   "l2": "3 rotating input batches; per-step activations exceed the 126 MB L2"},
        "e2e": {"value": world * B * steps / (ms_e2e * 1e-3), "unit": INFER_UNIT, "h2d_bytes_per_step": host[0].numel() * 4,
                "d2h_bytes_per_step": votes_host[0].numel() * 4, "ms_per_step": ms_e2e / steps},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf if peak_tf else None, "traffic": measured_traffic("infer", B, nmax),
                     "kernel": "gemm_tc_kernel (tcgen05 PointNet forward GEMMs, BatchNorm + ELU epilogue; last layer also mean-pools over points)",
                     "launches_timed": len(tc_events), "share_of_step": tc_ms / ms if ms else None, "peak_source": peak_src,
                     "whole_step_tensor_frac": (world * B * steps * 30 * nmax * FLOP_PER_POINT_FWD) / (ms * 1e-3) / 1e12 / (peak_tf * world)},
    }


def run_infer(args, world, rank, local, dev):
    line = measure_infer(args, world, rank, local, dev, args.batch if args.batch_given else 1020, args.steps, args.warmup)
    if world > 1:
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_infer_rate(args.nmax, args.k, n=480)
    print(json.dumps(line))


def cpu_infer_rate(nmax: int, k: int, n: int = 48):
    """The reference's inference arithmetic on the host cores: eval-mode encoder forward in windows of k crops plus the float64
    mixture likelihood and the vote, as inference_PCAA.py:239-271 does per window.  The encoder is the reference's own
    CGEncoder when baseline/_ref is installed (kind "reference"; the likelihood / vote are nested functions of the reference's
    procedure and are restated by the oracle), else the oracle port."""
    from oracle import pcaa_oracle as O
    from baseline import refenv
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    means = O.sample_distant_points(32, NCLS, 10, 10).float().numpy()
    n = (n // k) * k
    if refenv.available():
        from baseline import ref_loop
        fv, logits, dt = ref_loop.time_eval_encoder("cpu", nmax, NCLS, k, n // k)
        t0 = time.perf_counter()
        for w in range(n // k):
            lik = O.joint_likelihood(fv[w * k:(w + 1) * k].numpy(), means)
            O.openset_vote(lik, logits[w * k:(w + 1) * k].argmax(1).numpy(), 1e-30, k, NCLS)
        dt += time.perf_counter() - t0
        kind, what = "reference", "the reference's own CGEncoder (baseline/_ref) in eval mode on windows of k crops + oracle float64 likelihood + vote"
    else:
        p = O.det_params(NCLS, nmax, seed=0)
        pcs, _ = O.synth_batch(n, nmax, NCLS, seed=99)

        def run():
            with torch.no_grad():
                for w in range(n // k):
                    logits, fv = O.encoder_forward(p, pcs[w * k:(w + 1) * k], False, True)
                    lik = O.joint_likelihood(fv.numpy(), means)
                    O.openset_vote(lik, logits.argmax(1).numpy(), 1e-30, k, NCLS)
        run()
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        kind, what = "port", "oracle port: eval encoder forward in windows of k + float64 likelihood + vote"
    return {"value": n / dt, "unit": INFER_UNIT, "cores": threads, "kind": kind, "seconds": dt,
            "sample": f"{what}, k={k}, {n} crops, N={nmax} ({dt:.2f} s)"}


def run_b200(args):
    from opensetgaitrecognition_pcaa_b200 import _lib, dp, ops, synth
    from opensetgaitrecognition_pcaa_b200.train import build_variant4

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    B = args.batch
    NMAX = args.nmax
    if args.workload == "infer":
        return run_infer(args, world, rank, local, dev)
    trainer = build_variant4(NCLS, NMAX, seed=0, device=dev)
    if world > 1:
        # identical replicas: broadcast rank 0's initial weights (flat buffers)
        torch.distributed.broadcast(trainer.G.p, 0)
        torch.distributed.broadcast(trainer.D.p, 0)
        for b in list(trainer.enc.buffers()):
            torch.distributed.broadcast(b, 0)
        ops.convert_into(trainer.G.p, trainer.G.shadow)
    # distinct synthetic batches (host, pinned), rotated so consecutive steps never see the same input
    nb = 3
    host = []
    rng = np.random.default_rng(100 + rank)
    for i in range(nb):
        pcs, gt = synth.synth_batch(B, NMAX, NCLS, seed=1234 + 17 * rank + i)
        z0 = torch.from_numpy(rng.normal(0, 1, (B, 32))).float()
        al = torch.from_numpy(rng.uniform(0, 1, (B, 1)).astype(np.float32))
        host.append(tuple(t.pin_memory() for t in (pcs, gt, z0, al)))
    devb = [tuple(t.to(dev) for t in h) for h in host]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput, with CUDA events around every tensor-core GEMM launch
    tc_events = []
    orig_tc = ops.gemm_tc
    record = {"on": False}
    POINTNET_MODES = (_lib.TC_T_BIAS_STATS, _lib.TC_T_DGRAD_ELUBN, _lib.TC_WGRAD_ACC)

    def timed_gemm_tc(a, b, mode, M, N, K, **k):
        # CUDA events (torch's current stream = the launch stream) around every PointNet tcgen05 GEMM launch (the weight
        # gradients of the PointNet layers are the TC_WGRAD_ACC launches whose k extent is the point count)
        if not record["on"] or mode not in POINTNET_MODES or (mode == _lib.TC_WGRAD_ACC and K != B * 30 * NMAX):
            return orig_tc(a, b, mode, M, N, K, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_tc(a, b, mode, M, N, K, **k)
        e1.record()
        tc_events.append((e0, e1, 2.0 * M * N * K))
        return r

    ops.gemm_tc = timed_gemm_tc
    # CUDA-graph replay of the step (PCAATrainer.step_graphed): first call eager, second captures, then replays
    use_graph = args.graph != "off"
    stepfn = trainer.step_graphed if use_graph else trainer.step

    for i in range(args.warmup):
        stepfn(*devb[i % nb])
    barrier()
    # ---- data-parallel parity THROUGH the path that is about to be timed (split graphs + the selected exchange): one
    # iteration vs a single-device emulation of the same global iteration (dp.graphed_step_parity); the run fails if it does
    dp_parity = None
    if world > 1 and use_graph and not args.no_dp_parity:
        dp_parity = dp.graphed_step_parity(trainer, devb[0], lambda: build_variant4(NCLS, NMAX, seed=0, device=dev, process_group=dp.SINGLE))
        if not dp_parity["ok"]:
            if rank == 0:
                print(json.dumps({"error": "data-parallel parity failed", "dp_parity": dp_parity}))
            torch.distributed.destroy_process_group()
            sys.exit(3)
        stepfn(*devb[1 % nb])
        barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    calls0 = _lib.CALLS
    record["on"] = not use_graph
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        out = stepfn(*devb[i % nb])
    t1.record()
    barrier()
    record["on"] = False
    launches = trainer.graph_launches(devb[0][0].shape) * args.steps if use_graph else _lib.CALLS - calls0
    clocks = sampler.stop() if rank == 0 else None
    ms = t0.elapsed_time(t1)
    ms_eager = ms / args.steps
    r_steps = max(1, min(args.steps, 5))
    if use_graph:
        # per-launch CUDA events cannot be read back from inside a replayed graph: the tensor-core GEMM launches are
        # timed in an eager pass of the same step (same kernels, same inputs) right after the timed region
        trainer.step(*devb[0])         # untimed: graph capture emptied the allocator cache, this refills it
        barrier()
        record["on"] = True
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for i in range(r_steps):
            trainer.step(*devb[i % nb])
        r1.record()
        barrier()
        record["on"] = False
        ms_eager = r0.elapsed_time(r1) / r_steps
    tc_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in tc_events)
    tc_flops = sum(f for _, _, f in tc_events)
    tc_steps = r_steps if use_graph else args.steps
    ops.gemm_tc = orig_tc

    # ---- end-to-end through the public API with HOST buffers: pinned host batch -> H2D (copy stream, one batch ahead,
    # loader.DevicePrefetcher) -> train step -> D2H of the losses and predictions, host waits for them every step
    from opensetgaitrecognition_pcaa_b200.loader import DevicePrefetcher
    # The host reads every step's losses / predictions (the reference prints them each iteration, PCAA_ablation.py:1023-1030),
    # one step behind the device: the D2H copies of step i are stream-ordered right behind it (before the next replay
    # overwrites the graph-owned outputs) into one of two pinned buffers, and the host waits for them while step i+1 runs.
    res_host = [torch.empty(5, dtype=torch.float32).pin_memory() for _ in range(2)]
    pred_host = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_steps = args.steps
    # ONE loader for the whole measurement, as in a training epoch: its first two batches are untimed warm-up (copy
    # stream, device slots), then K timed steps during which K host->device copies are issued (one batch ahead), one spare
    # batch at the end keeps the last timed step's prefetch identical to the others
    feed = iter(DevicePrefetcher((host[i % nb] for i in range(2 + e2e_steps + 1)), dev, depth=2))
    for _ in range(2):
        stepfn(*next(feed))
    barrier()
    host_seen = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(e2e_steps):
        out = stepfn(*next(feed))
        j = i & 1
        res_host[j].copy_(torch.stack([out["rec_loss"], out["d_loss"], out["gp"], out["loss_g"], out["sup_loss"]]), non_blocking=True)
        pred_host[j].copy_(out["pred"], non_blocking=True)
        done[j].record()
        if i > 0:
            done[j ^ 1].synchronize()
            host_seen += float(res_host[j ^ 1][0]) + int(pred_host[j ^ 1][0])
    done[(e2e_steps - 1) & 1].synchronize()
    host_seen += float(res_host[(e2e_steps - 1) & 1][0])
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    d2h = res_host[0].numel() * 4 + pred_host[0].numel() * 4
    losses_last = {k: float(out[k]) for k in ("rec_loss", "d_loss", "sup_loss", "loss_g")}
    exchange = {"mode": "none"}
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        peer = trainer.G.peer
        exchange = {"mode": ("peer (copy engines over NVLink symmetric memory: chunked reduce-scatter / all-gather, one-shot for spans < 4 MB"
                             + ("; decoder span: reduce-scatter of gradients, Adam on the owned 1/world, all-gather of the updated fp32 weights and bf16 copies)"
                                if trainer.shard_adam else ")")) if peer is not None else "nccl",
                    "bytes_reduced_per_step": trainer.xG.bytes_reduced / max(1, trainer.G.step),
                    "bytes_pulled_per_step": (peer.bytes_pulled / max(1, trainer.G.step)) if peer is not None else 0,
                    "spans": ["critic", "decoder + projection head"] + (["encoder"] if trainer.enc_buckets == 1 else
                                                                        ["encoder: heads, TCN, PointNet 4-3", "encoder: PointNet 2-1"])}

    # ---- optional: where the step spends its time, phase by phase (CUDA events on the main stream, max over ranks)
    phases_ms = None
    if args.phases:
        trainer.phase_timing = True
        if not trainer.split_graphs:
            pstep = trainer.step                              # one rank: the phases of the eager step (the graph is one node)
        else:
            pstep = stepfn
        for i in range(3):
            pstep(*devb[i % nb])
        trainer.phase_ms()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        pe0.record()
        for i in range(10):
            pstep(*devb[i % nb])
        pe1.record()
        pm = trainer.phase_ms()
        pm["step_total"] = pe0.elapsed_time(pe1) / 10
        trainer.phase_timing = False
        keys = list(pm)
        tt = torch.tensor([pm[k] for k in keys], device=dev, dtype=torch.float64)
        if world > 1:
            tmax = tt.clone()
            torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
            tmin = tt.clone()
            torch.distributed.all_reduce(tmin, op=torch.distributed.ReduceOp.MIN)
        else:
            tmax = tmin = tt
        phases_ms = {"path": "eager step" if pstep == trainer.step else "split-graph replay + eager exchanges",
                     "max_over_ranks": {k: round(float(v), 4) for k, v in zip(keys, tmax)},
                     "min_over_ranks": {k: round(float(v), 4) for k, v in zip(keys, tmin)}}

    # ---- config 5 in the same record: open-set inference over a synthetic stream (>= 1 s timed, >= 1 M crops at 8 GPUs)
    inference_rec = None
    if not args.no_infer:
        del trainer, devb
        torch.cuda.empty_cache()
        ib = 1020
        isteps = 123            # 8 ranks x 123 batches x 1020 crops > 1 M crops streamed at --gpus 8; > 2 s timed per rank
        inference_rec = measure_infer(args, world, rank, local, dev, ib, isteps, 3)
    if world > 1:
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    value = world * B * args.steps / (ms * 1e-3)
    e2e = world * B * e2e_steps / (ms_e2e * 1e-3)
    peak_tf, peak_hbm, peak_src = peaks()
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"pcaa_variant4_train_step_N{NMAX}_C{NCLS}", "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"dp{world}", "l2": "per-step activations (>10 GB) and 3 rotating input batches exceed the 126 MB L2",
                   "launch": "cuda_graph_replay" if use_graph else "eager", "losses_last_step": losses_last,
                   "exchange": exchange, "dp_parity": dp_parity},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / e2e_steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf if peak_tf else None, "traffic": measured_traffic("train", B, NMAX),
                     "traffic_source": "profiles/traffic.json (regenerated from an ncu --set full capture of this command by scripts/roofline_table.py --write-traffic)",
                     "kernel": "gemm_tc_kernel (tcgen05 PointNet fwd/dgrad/wgrad GEMMs)",
                     "launches_timed": len(tc_events), "share_of_step": tc_ms / (ms_eager * tc_steps) if tc_events else None,
                     "timed_in": ("eager pass after the timed region (CUDA events around each launch; %.3f ms/step eager)" % ms_eager)
                     if use_graph else "the timed region", "peak_source": peak_src,
                     "whole_step_tensor_frac": (world * B * args.steps * 30 * NMAX * FLOP_PER_POINT_TC) / (ms * 1e-3) / 1e12 / (peak_tf * world)},
    }
    if phases_ms is not None:
        line["phases_ms"] = phases_ms
    if inference_rec is not None:
        line["inference"] = inference_rec
    if world == 1 and not args.no_cpu:
        # the reference's own loop on the host cores (bounded sample: batch 32, 1 warm-up + 5 timed iterations) and on cuda:0
        # (stock torch eager, same batch as this arm), each in a fresh process: `bench.py --impl reference`
        cpu = sub_bench(["--impl", "reference", "--steps", "5", "--warmup", "1", "--nmax", str(NMAX)])     # ~12 s of CPU work
        line["cpu_baseline"] = cpu.get("cpu_baseline", cpu)
        if inference_rec is not None:
            inference_rec["cpu_baseline"] = cpu_infer_rate(NMAX, args.k, n=480)                              # ~10 s
        torch.cuda.empty_cache()
        eager = sub_bench(["--impl", "reference", "--device", "cuda", "--batch", str(B), "--steps", "5", "--warmup", "2", "--nmax", str(NMAX)])
        if "error" in eager:
            line["gpu_eager_baseline"] = eager
        else:
            line["gpu_eager_baseline"] = {"value": eager["value"], "unit": UNIT, "ms_per_step": eager["ms_per_step"], "batch": B,
                                          "steps": eager["steps"], "warmup": eager["warmup"], "kind": eager["cpu_baseline"]["kind"],
                                          "what": eager["config"]["note"], "speedup_of_this_arm_e2e": e2e / eager["value"]}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (weak scaling); default 256 (train), 1020 (infer), "
                    "32 for the CPU reference arm")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--device", default=None, choices=[None, "cpu", "cuda"], help="--impl reference only: where the reference's "
                    "own loop runs (default cpu = the reference arm; cuda = stock torch eager on cuda:0)")
    ap.add_argument("--workload", default="train", choices=["train", "infer"],
                    help="train: one variant-4 AAE iteration per step (default, BASELINE.json metric) + an `inference` sub-record; "
                         "infer: eval-mode encoder + fused open-set scoring + k-window vote over one batch of crops per step")
    ap.add_argument("--nmax", type=int, default=NMAX, help="points per frame (train_pointsubsampling sweep: 50..150)")
    ap.add_argument("--k", type=int, default=6, help="voting window of the inference workload")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / gpu_eager_baseline legs")
    ap.add_argument("--no-infer", action="store_true", help="skip the inference sub-record of the train workload")
    ap.add_argument("--no-dp-parity", action="store_true", help="skip the data-parallel parity check (N > 1)")
    ap.add_argument("--phases", action="store_true", help="add per-phase CUDA-event timings of the step (phases_ms)")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the train step from a CUDA graph (auto: on)")
    args = ap.parse_args()
    args.batch_given = args.batch is not None
    if args.batch is None:
        args.batch = 256
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

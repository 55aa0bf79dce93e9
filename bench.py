#!/usr/bin/env python
"""bench.py -- the north-star measurement: QuartzNet asr13x1 + CTC training throughput in audio-seconds / second.

    python bench.py --gpus N --steps K --warmup W                 (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path on the host cores)

A "step" is one pass of the hot path over one per-GPU batch of synthetic input (SURVEY.md section 8d, config 2:
batch 32 x 16 s, bf16): H2D of the batch is excluded for `value` (inputs resident in HBM) and included for `e2e`.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LABELS28 = [" ", "'"] + [chr(ord("a") + i) for i in range(26)]  # conf/conf.yaml:12-13

WORKLOADS = {
    # name: (model_name, per-GPU batch, seconds, vocabulary, precision)
    "asr13x1_b32_16s_bf16": ("asr13x1", 32, 16.0, "labels28", "bf16"),
    "asr13x1_b4_10s_fp32": ("asr13x1", 4, 10.0, "labels28", "fp32"),
    "contextse_b64_20s_bf16": ("asr13x1contextse", 64, 20.0, "labels28", "bf16"),
    "context_aishell_b32_16s_bf16": ("asr13x1context", 32, 16.0, "aishell", "bf16"),
    # BASELINE config 5: validation / inference path (frontend + encoder in eval mode + greedy CTC decode)
    "infer_asr13x1_b256_30s_bf16": ("asr13x1", 256, 30.0, "labels28", "bf16"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def load_traffic(family):
    """Mean DRAM bytes (read + write) per launch of a kernel family, from the committed `ncu --set full` capture of
    the same workload (profiles/traffic.json, written by tools/ncu_traffic.py).  None if the family was not captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    try:
        d = json.load(open(p)).get(family)
        return None if d is None else d["dram_bytes_per_launch"]
    except (ValueError, KeyError):
        return None


def labels_for(vocab):
    if vocab == "labels28":
        return LABELS28
    # AISHELL-1 char vocabulary (data/aishell1-vocab.txt has 4333 entries): only its SIZE matters for synthetic data
    return [chr(0x4E00 + i) for i in range(4333)]


# ---------------------------------------------------------------------------------------------------------------
# algorithmic bytes / flops per C-ABI call (DESIGN.md "kernels"): args are the call's int/float arguments in order
# ---------------------------------------------------------------------------------------------------------------
def algorithmic(name, a, has):
    """-> (family, bytes, flops) for one entry-point call; es = element size of the activation dtype."""
    def es(code):
        return 4 if code == 0 else 2
    if name == "lasr_pwconv_fwd":  # T, M, Cin, Cout, ldx, ldw, ldy, dtype
        T, M, Cin, Cout, ldx, ldw, ldy, dt = a
        return "pwconv_gemm", es(dt) * M * (Cin + Cout), 2.0 * M * Cin * Cout
    if name == "lasr_pwconv_dgrad":  # M, Cin, Cout, lddy, ldw, lddx, dtype
        M, Cin, Cout, _, _, _, dt = a
        return "pwconv_gemm", es(dt) * M * (Cin + Cout), 2.0 * M * Cin * Cout
    if name == "lasr_pwconv_wgrad":  # M, Cin, Cout, lddy, ldx, lddw, dtype
        M, Cin, Cout, _, _, _, dt = a
        return "pwconv_gemm", es(dt) * M * (Cin + Cout) + 4 * Cin * Cout, 2.0 * M * Cin * Cout
    if name == "lasr_pwconv_fwd2":  # T, M, Cin, Cout, dtype  (two problems)
        T, M, Cin, Cout, dt = a
        return "pwconv_gemm", 2 * es(dt) * M * (Cin + Cout), 4.0 * M * Cin * Cout
    if name == "lasr_pwconv_dgrad2":  # M, Cin, Cout, dtype
        M, Cin, Cout, dt = a
        return "pwconv_gemm", 2 * es(dt) * M * (Cin + Cout), 4.0 * M * Cin * Cout
    if name == "lasr_pwconv_wgrad2":  # M, Cin, Cout, lddy, ldx, lddw, dtype
        M, Cin, Cout, _, _, _, dt = a
        return "pwconv_gemm", 2 * (es(dt) * M * (Cin + Cout) + 4 * Cin * Cout), 4.0 * M * Cin * Cout
    if name == "lasr_dwconv1d_bwd":  # N, T, C, K, dtype; ptrs x, dy, w, addend, dx, dw: dgrad (+addend) and wgrad
        N, T, C, K, dt = a
        return "dwconv", es(dt) * N * C * T * (4 + (1 if has[3] else 0)) + 4 * C * K, 4.0 * N * T * C * K
    if name == "lasr_dwconv1d_fwd":  # N, T_in, T_out, C, K, stride, flip, dtype
        N, Ti, To, C, K, s, flip, dt = a
        return "dwconv", es(dt) * N * C * (Ti + To + (To if has[3] else 0)) + 4 * C * K, 2.0 * N * To * C * K
    if name == "lasr_dwconv1d_wgrad":  # N, T_in, T_out, C, K, stride, dtype
        N, Ti, To, C, K, s, dt = a
        return "dwconv", es(dt) * N * C * (Ti + To), 2.0 * N * To * C * K
    if name == "lasr_bn_apply_act_fwd":  # M, C, T, count, eps, momentum, act, side_effects, dtype; ptrs y,bn1,r,bn2,gate,out
        M, C, dt = a[0], a[1], a[8]
        return "bn_pass", es(dt) * M * C * (3 if has[2] else 2), 4.0 * M * C
    if name == "lasr_bn_act_bwd_reduce":  # N, T, C, act, dtype; ptrs dout,out,y,r,totals,per_n
        N, T, C, act, dt = a
        return "bn_pass", es(dt) * N * T * C * ((3 if act else 2) + (1 if has[3] else 0)), 6.0 * N * T * C
    if name == "lasr_bn_act_bwd_apply":  # count, T, M, C, act, dtype; ptrs dout,out,y,r,...
        _, T, M, C, act, dt = a
        return "bn_pass", es(dt) * M * C * ((4 if act else 3) + (2 if has[3] else 0)), 6.0 * M * C
    if name == "lasr_novograd_step":  # reads p, g, m; writes p, m (+ bf16 shadow): 22 B per parameter element
        return "novograd", 0, 0.0
    return name.replace("lasr_", ""), 0, 0.0


def kernel_breakdown(engine, steps=3):
    """Eager (non-graph) steps with CUDA events around every C-ABI call on the launching stream."""
    import torch
    from lightning_asr_b200 import _lib

    fam = {}
    calls = 0
    for i in range(steps):
        _lib.PROFILE = []
        _lib.CALLS["n"] = 0
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        engine._step_eager()
        t1.record()
        torch.cuda.synchronize()
        prof, _lib.PROFILE = _lib.PROFILE, None
        calls = _lib.CALLS["n"]
        if i == 0:
            continue  # warm-up of the instrumented path
        for name, args, has, e0, e1 in prof:
            f, b, fl = algorithmic(name, args, has)
            d = fam.setdefault(f, {"ms": 0.0, "bytes": 0, "flops": 0.0, "calls": 0})
            d["ms"] += e0.elapsed_time(e1)
            d["bytes"] += b
            d["flops"] += fl
            d["calls"] += 1
    n = max(steps - 1, 1)
    for d in fam.values():
        for k in d:
            d[k] = d[k] / n
    return fam, calls


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: an NVML polling thread (5 ms period; the timed
    region is only ~100 ms long, nvidia-smi's 100 ms loop would see one or two samples).  Falls back to an
    `nvidia-smi -lms` child process if NVML cannot be opened."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.thread = None
        self.samples = []
        self.smax = 0.0
        self.reason_bits = 0
        self._stop = False

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _poll(self, nv, h):
        while not self._stop:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import threading

            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.nv = nv
            self.thread = threading.Thread(target=self._poll, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            nv = self.nv
            names = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                     ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                     ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)]
            if not self.samples:
                return None
            return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.smax,
                    "reasons": sorted(n for n, bit in names if self.reason_bits & bit), "samples": len(self.samples),
                    "source": "nvml, 5 ms period, during the timed regions"}
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = max(smax, float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        busy = [v for v in sm if v > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi -lms 100"}


# ---------------------------------------------------------------------------------------------------------------
# CPU arms: the reference's own algorithm (oracle port: the reference is pure Python on torch, it cannot travel to the
# GPU box, so the restated oracle is what runs there) timed on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_training_throughput(model_name, seconds, labels, sample_n, steps, warmup, precision_note="fp32"):
    import torch
    from lightning_asr_b200.quartznet import build_model
    from lightning_asr_b200.trainer import synthetic_batch
    from oracle import optim_oracle, train_oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = {k: v.detach().clone() for k, v in build_model(model_name, labels, mask=True).state_dict().items()}
    params = [v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k]
    batch = synthetic_batch(sample_n, seconds, len(labels), seed=1234, ragged=False)
    times = []
    opt_state = [{} for _ in params]
    for i in range(warmup + steps):
        for p in params:
            p.grad = None
        t0 = time.perf_counter()
        loss, _, _ = train_oracle.training_step(sd, batch, labels, mask=True, training=True, update_buffers=True)
        loss.backward()
        with torch.no_grad():  # the reference's optimizer step (scheduler/novograd.py, train.py:46), like the GPU arm
            optim_oracle.novograd_step([p.data for p in params], [p.grad for p in params], opt_state, 1e-4,
                                       betas=(0.8, 0.5), weight_decay=1e-4)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return {"value": sample_n * seconds / med, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_n} x {seconds:g} s utterances of the workload, fp32 fwd+bwd+CTC+Novograd, median of {steps} "
                      f"steps ({med * 1e3:.0f} ms/step)", "ms_per_step": med * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model_name, n, seconds, vocab, precision = WORKLOADS[args.workload]
    labels = labels_for(vocab)
    sample_n = min(n, 4)
    steps = max(1, min(args.steps, 5))
    warmup = max(1, min(args.warmup, 2))
    cb = cpu_training_throughput(model_name, seconds, labels, sample_n, steps, warmup)
    line = {
        "impl": "reference", "metric": "train audio-seconds/sec (QuartzNet+CTC)", "value": cb["value"],
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "model_name": model_name, "per_gpu_batch": n, "seconds": seconds,
                   "frames": 1 + (int(seconds * 16000) + 64) // 160,
                   "encoder_steps": (1 + (int(seconds * 16000) + 64) // 160 - 1) // 2 + 1, "vocab": len(labels) + 1,
                   "mask": True, "step": "forward + CTC + backward + Novograd update",
                   "note": "reference algorithm (oracle port of models/QuartNet.py + torch CTCLoss + "
                           "scheduler/novograd.py) on the host cores; each step is a bounded sample of the workload "
                           f"({sample_n} of the {n} utterances)"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_infer(args):
    """Inference workload: waveforms (pinned host) -> H2D -> log-mel -> encoder (eval) -> greedy decode -> tokens D2H."""
    import torch
    import torch.distributed as dist

    from lightning_asr_b200 import _lib
    from lightning_asr_b200.trainer import InferEngine, LightingModule, synthetic_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _lib.require_device()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model_name, n, seconds, vocab, precision = WORKLOADS[args.workload]
    labels = labels_for(vocab)
    peaks = load_peaks()
    torch.manual_seed(0)
    module = LightingModule(labels=labels, mask=True, drop_rate=0.0, model_name=model_name,
                            precision=precision).cuda().eval()
    (waves, lens), _, _, _, _ = synthetic_batch(n, seconds, len(labels), seed=1234 + rank, ragged=False, features=False)
    engine = InferEngine(module, waves, lens, graph=not args.no_graph)
    for _ in range(max(args.warmup, 3)):
        engine.step_device()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        engine.step_device()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    f0.record()
    engine.prefetch()
    for i in range(args.steps):
        toks, cnts = engine.step_host(prefetch_next=(i + 1 < args.steps))
    f1.record()
    barrier()
    e2e_ms = max_over_ranks(max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3))
    clocks = sampler.stop() if rank == 0 else None
    fam, calls = kernel_breakdown(engine, steps=3)
    if world > 1:
        barrier()
    if rank != 0:
        sys.stdout.flush()
        os._exit(0)
    ms_per_step = ms_total / args.steps
    audio_s = n * seconds * world
    T = 1 + (int(seconds * 16000) + 64) // 160
    Tp = (T - 1) // 2 + 1
    # schedule-L forward bytes: asr13x1 V'=29 fwd = 50 775 elements per encoder step, + the waveform
    sched_bytes = 50775 * 2 * n * Tp + 4 * n * int(seconds * 16000)
    total_ms = sum(d["ms"] for d in fam.values())
    tname, t = max(fam.items(), key=lambda kv: kv[1]["ms"])
    ach = (t["flops"] / (t["ms"] * 1e-3) / 1e12) if tname == "pwconv_gemm" else (t["bytes"] / (t["ms"] * 1e-3) / 1e9)
    peak = peaks["bf16_tflops_sustained"] if tname == "pwconv_gemm" else peaks["hbm_gbs"]
    roof = {"bound": "tensor" if tname == "pwconv_gemm" else "hbm", "achieved": ach, "peak": peak,
            "unit": "TFLOP/s" if tname == "pwconv_gemm" else "GB/s", "frac": ach / peak, "traffic": None,
            "kernel": tname, "peak_source": peaks["source"] + " (sustained)",
            "share_of_step": t["ms"] / total_ms if total_ms else None,
            "families": {k: {"ms": round(v["ms"], 4), "calls": v["calls"]} for k, v in fam.items()},
            "step_hbm_frac_scheduleL": sched_bytes / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    line = {
        "metric": "inference audio-seconds/sec (log-mel + QuartzNet + greedy CTC decode)",
        "value": audio_s / (ms_per_step * 1e-3), "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": precision, "data": "synthetic",
        "config": {"workload": args.workload, "model_name": model_name, "per_gpu_batch": n, "seconds": seconds,
                   "frames": T, "encoder_steps": Tp, "vocab": len(labels) + 1, "mask": True,
                   "parallelism": f"dp{world} (independent replicas, no collective)",
                   "step": "H2D waveforms -> log-mel -> encoder (eval) -> decoder -> greedy CTC decode -> tokens D2H",
                   "cuda_graph": not args.no_graph,
                   "l2": "no flush needed: each pass streams tens of GB of activations >> 126 MB L2"},
        "e2e": {"value": audio_s / (e2e_ms / args.steps * 1e-3), "unit": "audio-s/s", "h2d_bytes_per_step": engine.h2d_bytes,
                "d2h_bytes_per_step": int(toks.numel() * 4 + cnts.numel() * 4), "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": calls * args.steps, "roofline": roof, "cpu_baseline": None, "clocks": clocks,
        "tokens_decoded": int(cnts.sum()),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        os._exit(0)


def run_b200(args):
    import torch
    import torch.distributed as dist

    from lightning_asr_b200 import _lib, ddp
    from lightning_asr_b200.trainer import LightingModule, TrainEngine, synthetic_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _lib.require_device()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model_name, n, seconds, vocab, precision = WORKLOADS[args.workload]
    labels = labels_for(vocab)
    peaks = load_peaks()

    torch.manual_seed(0)
    module = LightingModule(labels=labels, mask=True, drop_rate=0.0, model_name=model_name,
                            precision=precision).cuda().train()
    if world > 1:
        ddp.broadcast_parameters(module)
    batch = synthetic_batch(n, seconds, len(labels), seed=1234 + rank, ragged=False)
    use_graph = not args.no_graph
    engine = TrainEngine(module, batch, graph=use_graph, fused=True, world_sync=(None, float(os.environ.get("LASR_BUCKET_MB", "8"))) if world > 1 else None,
                         optimizer=None if args.no_optimizer else "novograd")
    graph_note = use_graph
    try:
        for _ in range(max(args.warmup, 3)):
            engine.step_device()
        torch.cuda.synchronize()
    except Exception as e:  # graph capture refused (e.g. a collective that cannot be captured): eager launches
        if not use_graph:
            raise
        sys.stderr.write(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); falling back to eager\n")
        torch.cuda.synchronize()
        engine.use_graph, engine.graph, graph_note = False, None, False
        for _ in range(max(args.warmup, 3)):
            engine.step_device()
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident timing (value) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        engine.step_device()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    # ---- end-to-end timing (pinned host batch -> H2D -> step -> loss on the host, every step) ----
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    f0.record()
    loss_val = 0.0
    engine.prefetch()  # H2D of step 0's batch; every later step's H2D is issued inside the loop (one per step)
    for i in range(args.steps):
        # one H2D of the batch and one D2H of the loss per step; the host reads step i-1's loss while step i runs
        engine.step_host(prefetch_next=(i + 1 < args.steps), defer_loss=True)
    loss_val = engine.flush_loss()
    f1.record()
    barrier()
    e2e_ms = max_over_ranks(max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3))
    clocks = sampler.stop() if rank == 0 else None

    ms_per_step = ms_total / args.steps
    audio_s = n * seconds * world
    value = audio_s / (ms_per_step * 1e-3)
    e2e_value = audio_s / (e2e_ms / args.steps * 1e-3)

    # ---- per-kernel breakdown with CUDA events (eager, after the timed region) ----
    fam, calls = kernel_breakdown(engine, steps=3)
    if world > 1:
        # every rank is done with the device; leave together.  The process group is NOT torn down explicitly: with
        # NCCL all-reduces captured in a live CUDA graph destroy_process_group() can block forever, so the ranks
        # simply exit (os._exit below) once rank 0 has printed its line.
        barrier()
    if rank != 0:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    total_ms = sum(d["ms"] for d in fam.values())
    ncalls = sum(d["calls"] for d in fam.values())
    over = max(0.0, (total_ms - ms_per_step) / ncalls) if ncalls else 0.0

    def in_step(d):
        return max(d["ms"] - d["calls"] * over, 0.25 * d["ms"])
    tname, t = max(fam.items(), key=lambda kv: in_step(kv[1]))  # the dominant family of the timed step
    t_ms = in_step(t)
    if tname == "pwconv_gemm":
        peak = peaks["bf16_tflops_sustained"]
        ach = t["flops"] / (t_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "achieved_eager_events": t["flops"] / (t["ms"] * 1e-3) / 1e12}
    else:
        peak = peaks["hbm_gbs"]
        ach = t["bytes"] / (t_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "achieved_eager_events": t["bytes"] / (t["ms"] * 1e-3) / 1e9}
    roof.update({"traffic": load_traffic(tname), "kernel": tname, "peak_source": peaks["source"] + " (sustained)",
                 "share_of_step": t_ms / ms_per_step, "ms_in_step": t_ms, "launches_per_step": t["calls"],
                 "eager_launch_overhead_us": over * 1e3,
                 "timing": "CUDA events around every C-ABI call (eager replay, launching stream) minus the calibrated "
                           "per-call launch latency, so that the families sum to the measured graph step",
                 "families": {k: {"ms": round(v["ms"], 4), "ms_in_step": round(in_step(v), 4),
                                  "GB/s": round(v["bytes"] / max(in_step(v), 1e-9) / 1e6, 1),
                                  "TFLOP/s": round(v["flops"] / max(in_step(v), 1e-9) / 1e9, 1), "calls": v["calls"],
                                  "frac_of_bound": round(
                                      (v["flops"] / max(in_step(v), 1e-9) / 1e9 / peaks["bf16_tflops_sustained"])
                                      if k == "pwconv_gemm" else
                                      (v["bytes"] / max(in_step(v), 1e-9) / 1e6 / peaks["hbm_gbs"]), 3)}
                              for k, v in fam.items()}})
    # schedule-L bytes of the whole step: asr13x1 V'=29: 159 726 elements per encoder step
    T = 1 + (int(seconds * 16000) + 64) // 160
    Tp = (T - 1) // 2 + 1
    es = 2 if precision == "bf16" else 4
    if model_name == "asr13x1":
        sched_bytes = 159726 * es * n * Tp
        roof["step_hbm_frac_scheduleL"] = sched_bytes / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"]

    cb = None
    if not args.no_cpu_baseline:
        cb = cpu_training_throughput(model_name, seconds, labels, min(n, 4), 3, 1)
        cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {
        "metric": "train audio-seconds/sec (QuartzNet+CTC)", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": precision, "data": "synthetic",
        "config": {"workload": args.workload, "model_name": model_name, "per_gpu_batch": n, "seconds": seconds,
                   "frames": T, "encoder_steps": Tp, "vocab": len(labels) + 1, "mask": True, "parallelism": f"dp{world}",
                   "step": "forward + CTC + backward" + (" + NCCL grad all-reduce" if world > 1 else "")
                           + ("" if args.no_optimizer else " + fused Novograd/LR-schedule update"),
                   "cuda_graph": bool(graph_note),
                   "l2": "no flush needed: each step streams ~8 GB of activations >> 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": engine.h2d_bytes,
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": calls * args.steps, "roofline": roof, "cpu_baseline": cb, "clocks": clocks,
        "loss": loss_val,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="asr13x1_b32_16s_bf16", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-optimizer", action="store_true", help="time fwd+bwd only (diagnostics; not the headline)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload.startswith("infer_"):
        run_infer(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

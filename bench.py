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
    """(mean DRAM bytes read + written per launch of a kernel family, where that number comes from).  DRAM counters
    cannot be read without a profiler, so this is NOT measured in the run: it is the `ncu --set full` capture of this
    same bench command committed under profiles/ (tools/ncu_traffic.py writes profiles/traffic.json with the capture it
    was derived from).  (None, None) when the family was not captured at the current kernel revision."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    try:
        d = json.load(open(p))
        e = d.get(family)
        if e is None:
            return None, None
        return e["dram_bytes_per_launch"], e.get("source", d.get("source", "profiles/traffic.json"))
    except (ValueError, KeyError):
        return None, None


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
        # SURVEY.md 8d "dwconv bwd (fused)": every logical tensor once -- read x, read dy, write dx (+ read addend)
        N, T, C, K, dt = a
        return "dwconv", es(dt) * N * C * T * (3 + (1 if has[3] else 0)) + 4 * C * K, 4.0 * N * T * C * K
    if name == "lasr_dwconv1d_fwd":  # N, T_in, T_out, C, K, stride, flip, dtype
        N, Ti, To, C, K, s, flip, dt = a
        return "dwconv", es(dt) * N * C * (Ti + To + (To if has[3] else 0)) + 4 * C * K, 2.0 * N * To * C * K
    if name == "lasr_dwconv1d_wgrad":  # N, T_in, T_out, C, K, stride, dtype
        N, Ti, To, C, K, s, dt = a
        return "dwconv", es(dt) * N * C * (Ti + To), 2.0 * N * To * C * K
    if name == "lasr_dwconv1d_fwd_cm":  # N, T, C, K, S, flip; ptrs xT, w, y, addend, addendT (series operands, bf16)
        N, T, C, K = a[0], a[1], a[2], a[3]
        return "dwconv", 2 * N * C * T * (2 + (1 if (has[3] or has[4]) else 0)) + 4 * C * K, 2.0 * N * T * C * K
    if name == "lasr_dwconv1d_bwd_cm":  # N, T, C, K, S; ptrs xT, dyT, w, addend, addendT, dx, dw
        N, T, C, K = a[0], a[1], a[2], a[3]
        return "dwconv", 2 * N * C * T * (3 + (1 if (has[3] or has[4]) else 0)) + 4 * C * K, 4.0 * N * T * C * K
    if name == "lasr_pwconv_dgrad_cm":  # N, T, Cin, Cout, S, off; ptrs dy1, w1, dxT1, dy2, w2, dxT2 (one or two problems)
        N, T, Cin, Cout = a[0], a[1], a[2], a[3]
        k = 2 if has[3] else 1
        return "pwconv_gemm", k * 2 * N * T * (Cin + Cout), k * 2.0 * N * T * Cin * Cout
    if name == "lasr_bn_apply_act_fwd_cm":  # N, T, C, S, off, ...; ptrs y, bn1, r, bn2, gate, out, outT: + the series write
        N, T, C = a[0], a[1], a[2]
        return "bn_pass", 2 * N * T * C * ((3 if has[2] else 2) + 1), 4.0 * N * T * C
    if name == "lasr_bn_apply_act_fwd":  # M, C, T, count, eps, momentum, act, side_effects, dtype; ptrs y,bn1,r,bn2,gate,out
        M, C, dt = a[0], a[1], a[8]
        return "bn_pass", es(dt) * M * C * (3 if has[2] else 2), 4.0 * M * C
    if name == "lasr_bn_act_bwd_reduce":  # N, T, C, act, dtype; ptrs dout,out,y,r,totals,per_n
        N, T, C, act, dt = a
        # the ReLU gate is read from `out` (has[1]) or from its sign bits (1 byte per 8 channels)
        gate_b = (es(dt) if has[1] else 0.125) if act else 0
        return "bn_pass", N * T * C * (2 * es(dt) + gate_b + (es(dt) if has[3] else 0)), 6.0 * N * T * C
    if name == "lasr_bn_act_bwd_apply":  # count, T, M, C, act, dtype; ptrs dout,out,y,r,...
        _, T, M, C, act, dt = a
        gate_b = (es(dt) if has[1] else 0.125) if act else 0
        return "bn_pass", M * C * (3 * es(dt) + gate_b + (2 * es(dt) if has[3] else 0)), 6.0 * M * C
    if name == "lasr_novograd_step":  # reads p, g, m; writes p, m (+ bf16 shadow): 22 B per parameter element
        return "novograd", 0, 0.0
    return algorithmic_family(name), 0, 0.0


def algorithmic_family(name):
    if name.startswith("lasr_pwconv"):
        return "pwconv_gemm"
    if name.startswith("lasr_dwconv"):
        return "dwconv"
    if name in ("lasr_bn_apply_act_fwd", "lasr_bn_apply_act_fwd_cm", "lasr_bn_act_bwd_reduce", "lasr_bn_act_bwd_apply"):
        return "bn_pass"
    if name == "lasr_novograd_step":
        return "novograd"
    return name.replace("lasr_", "")


def kernel_breakdown(engine, steps=3):
    """Per-family device time of one step.  With a CUDA-graph engine: MEASURED inside a graph (trainer.
    profile_step_graph: timing events as external event-record nodes around every C-ABI call, median over replays);
    eager engines: CUDA events around every call on the launching stream.  -> (families, calls per step, info)."""
    import torch
    from lightning_asr_b200 import _lib
    from lightning_asr_b200.trainer import profile_step_graph

    fam = {}
    info = {}
    if getattr(engine, "use_graph", False):
        try:
            # pass 1: every call instrumented -> the call list, algorithmic bytes / flops, and a first time per family
            calls, step_all = profile_step_graph(engine._step_eager, replays=max(steps, 3))
            for name, args, has, ms in calls:
                f, b, fl = algorithmic(name, args, has)
                d = fam.setdefault(f, {"ms": 0.0, "bytes": 0, "flops": 0.0, "calls": 0})
                d["ms"] += ms
                d["bytes"] += b
                d["flops"] += fl
                d["calls"] += 1
            for d in fam.values():
                d["ms_all_instrumented"] = d["ms"]
            # pass 2: ONE family instrumented per captured graph (the rest of the step undisturbed) for the heavy ones
            steps_one = {}
            for f in sorted(fam, key=lambda k: -fam[k]["ms"])[:4]:
                sel, st = profile_step_graph(engine._step_eager, replays=max(steps, 3),
                                             only=lambda n, f=f: algorithmic_family(n) == f)
                fam[f]["ms"] = sum(ms for _, _, _, ms in sel if ms is not None)
                steps_one[f] = st
            info = {"timing": "CUDA events recorded as external event-record nodes INSIDE a CUDA graph of the step around "
                              "the C-ABI calls of ONE kernel family per captured graph (the other families run "
                              "undisturbed), median of replays; small families from the all-instrumented graph. No fitted "
                              "launch overhead.",
                    "instrumented_graph_step_ms": {"all_families": step_all, **steps_one}}
            return fam, len(calls), info
        except Exception as e:  # pragma: no cover - e.g. external event nodes refused by the driver
            sys.stderr.write(f"[bench] graph-node timing failed ({type(e).__name__}: {e}); eager events instead\n")
            _lib.PROFILE, _lib.PROFILE_EXTERNAL, _lib.PROFILE_ONLY = None, False, None
            torch.cuda.synchronize()
            fam = {}
    calls = 0
    for i in range(steps):
        _lib.PROFILE = []
        _lib.CALLS["n"] = 0
        engine._step_eager()
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
    info = {"timing": "CUDA events around every C-ABI call, eager launches on the launching stream (each pair also "
                      "contains the call's launch latency)"}
    return fam, calls, info


def roofline_of(fam, info, ms_per_step, peaks):
    """The `roofline` object for the dominant kernel family of the measured step."""
    tname, t = max(fam.items(), key=lambda kv: kv[1]["ms"])
    t_ms = t["ms"]
    if tname == "pwconv_gemm":
        peak = peaks["bf16_tflops_sustained"]
        ach = t["flops"] / (t_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak}
    else:
        peak = peaks["hbm_gbs"]
        ach = t["bytes"] / (t_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak}
    total = sum(d["ms"] for d in fam.values())
    traffic, traffic_src = load_traffic(tname)
    roof.update({
        "traffic": traffic, "traffic_source": traffic_src, "kernel": tname,
        "peak_source": peaks["source"] + " (sustained: the kernel is timed inside a long step)",
        "share_of_step": t_ms / ms_per_step, "ms_in_step": t_ms, "launches_per_step": t["calls"],
        "kernel_ms_sum": total, "graph_gaps_ms": ms_per_step - total,
        "families": {k: {"ms": round(v["ms"], 4), "ms_all_instrumented": round(v.get("ms_all_instrumented", v["ms"]), 4),
                         "GB/s": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1),
                         "TFLOP/s": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1), "calls": v["calls"],
                         "frac_of_bound": round((v["flops"] / max(v["ms"], 1e-9) / 1e9 / peaks["bf16_tflops_sustained"])
                                                if k == "pwconv_gemm" else
                                                (v["bytes"] / max(v["ms"], 1e-9) / 1e6 / peaks["hbm_gbs"]), 3)}
                     for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}})
    roof.update(info)
    return roof


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
# Baseline arms: the reference's own algorithm (oracle port: the reference is pure Python on torch and cannot travel to
# the GPU box, so the restatement pinned to it by tests/test_oracle_golden.py is what runs there)
# ---------------------------------------------------------------------------------------------------------------
def config_of(workload, world, extra=None):
    model_name, n, seconds, vocab, precision = WORKLOADS[workload]
    T = 1 + (int(seconds * 16000) + 64) // 160
    cfg = {"workload": workload, "model_name": model_name, "per_gpu_batch": n, "seconds": seconds, "frames": T,
           "encoder_steps": (T - 1) // 2 + 1, "vocab": len(labels_for(vocab)) + 1, "mask": True,
           "parallelism": f"dp{world}"}
    cfg.update(extra or {})
    return cfg


def oracle_training_throughput(model_name, seconds, labels, n, steps, warmup, device="cpu", autocast=None):
    """The reference's training step (oracle port of models/*.py + torch CTCLoss + scheduler/novograd.py, train.py:64-86
    and :36-62) on `device`: fp32 on the host cores, or eager PyTorch on the GPU (fp32 with TF32 off, or bf16 autocast)
    -- SURVEY.md 2.1's "eager PyTorch on the same box".  The whole per-GPU batch of the workload is processed."""
    import torch
    from lightning_asr_b200.quartznet import build_model
    from lightning_asr_b200.trainer import synthetic_batch
    from oracle import optim_oracle, train_oracle

    cores = os.cpu_count() or 1
    on_gpu = device != "cpu"
    if not on_gpu:
        torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = {k: v.detach().clone().to(device) for k, v in build_model(model_name, labels, mask=True).state_dict().items()}
    params = [v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k]
    from oracle import frontend_oracle
    (waves, lens), targets, _, tgt_sizes, _ = synthetic_batch(n, seconds, len(labels), seed=1234, ragged=False,
                                                              features=False)

    def host_frontend():
        """data_module.py:150-174 per utterance (dither / augmentation off) + _collate_fn :222-248."""
        feats = [frontend_oracle.logmel(waves[j, : int(lens[j])]) for j in range(n)]
        inputs, _, percents, _, _ = frontend_oracle.collate([(f, [0], "") for f in feats])
        return (inputs, targets, percents, tgt_sizes)

    batch = host_frontend()
    batch = tuple(t.to(device) if torch.is_tensor(t) else t for t in batch)
    if on_gpu:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    times = []
    opt_state = [{} for _ in params]
    for i in range(warmup + steps):
        for p in params:
            p.grad = None
        if on_gpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        if not on_gpu:  # the reference computes the features on the host (DataLoader workers): part of its path
            batch = host_frontend()
        if autocast is not None:
            with torch.autocast("cuda", dtype=autocast):
                loss, _, _ = train_oracle.training_step(sd, batch, labels, mask=True, training=True, update_buffers=True)
        else:
            loss, _, _ = train_oracle.training_step(sd, batch, labels, mask=True, training=True, update_buffers=True)
        loss.backward()
        with torch.no_grad():  # the reference's optimizer step (scheduler/novograd.py, train.py:46), like the B200 arm
            optim_oracle.novograd_step([p.data for p in params], [p.grad for p in params], opt_state, 1e-4,
                                       betas=(0.8, 0.5), weight_decay=1e-4)
        if on_gpu:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return {"value": n * seconds / med, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"the whole per-GPU batch ({n} x {seconds:g} s), "
                      + ("log-mel frontend + " if not on_gpu else "features resident on the device, ")
                      + f"fwd+bwd+CTC+Novograd, median of {steps} steps after {warmup} warm-up ({med * 1e3:.0f} ms/step)",
            "ms_per_step": med * 1e3}


def oracle_inference_throughput(model_name, seconds, labels, n, steps, warmup):
    """The reference's validation path on the host cores: frontend (data_module.py:155-172 restated) -> model in eval mode
    -> argmax -> greedy collapse (utils/asr_metrics.py:153-171), fp32."""
    import torch
    from lightning_asr_b200.quartznet import build_model
    from lightning_asr_b200.trainer import synthetic_batch
    from oracle import ctc_oracle, frontend_oracle, quartznet_oracle

    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    sd = {k: v.detach().clone() for k, v in build_model(model_name, labels, mask=True).state_dict().items()}
    (waves, lens), _, _, _, _ = synthetic_batch(n, seconds, len(labels), seed=1234, ragged=False, features=False)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        with torch.no_grad():
            feats = [frontend_oracle.logmel(waves[j, : int(lens[j])]) for j in range(n)]
            inputs, _, percents, _, _ = frontend_oracle.collate([(f, [0], "") for f in feats])
            out = quartznet_oracle.model(inputs, percents, sd, mask=True, training=False)
            t_len = torch.mul(out.size(1), percents).int()
            ctc_oracle.ctc_decoder_predictions(out.argmax(-1).tolist(), labels, t_len.tolist())
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return {"value": n * seconds / med, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} x {seconds:g} s utterances of the workload, fp32 frontend + eval forward + greedy decode, "
                      f"median of {steps} passes ({med * 1e3:.0f} ms/pass)", "ms_per_step": med * 1e3}


def gpu_eager_baseline(model_name, seconds, labels, n):
    """SURVEY.md 2.1 / BASELINE.md 4: the reference's eager PyTorch path ON THE B200 (cuDNN / cuBLAS / ATen kernels), the
    kernel-level bar.  Uses the oracle's vectorised MaskCNN (the reference's own builds the mask on the host with N
    .item() syncs per layer, models/QuartNet.py:309-321, which would only make this baseline slower)."""
    import torch
    out = {}
    for key, ac in (("bf16_autocast", torch.bfloat16), ("fp32_tf32_off", None)):
        try:
            r = oracle_training_throughput(model_name, seconds, labels, n, 5, 2, device="cuda", autocast=ac)
            out[key] = {"value": r["value"], "unit": r["unit"], "ms_per_step": r["ms_per_step"]}
        except Exception as e:  # e.g. out of memory on a large workload
            out[key] = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.cuda.empty_cache()
    out["what"] = ("oracle port of the reference modules + torch.nn.functional.ctc_loss + the reference's Novograd, eager "
                   "PyTorch on this GPU, whole per-GPU batch, FEATURES already resident on the device (the reference "
                   "computes them in CPU DataLoader workers; our timed step includes the frontend), median of 5 steps "
                   "after 2 warm-up, wall clock with synchronize on both sides")
    return out


def run_reference(args):
    """`--impl reference`: the reference's CPU path (oracle port) on the host cores, same workload, same per-GPU batch,
    honouring --steps / --warmup.  Under torchrun rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model_name, n, seconds, vocab, precision = WORKLOADS[args.workload]
    labels = labels_for(vocab)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    infer = args.workload.startswith("infer_")
    if infer:
        n_run = min(n, 32)  # 256 x 30 s of fp32 activations do not fit a sensible host-memory budget: bounded sample
        cb = oracle_inference_throughput(model_name, seconds, labels, n_run, steps, warmup)
        metric = "inference audio-seconds/sec (log-mel + QuartzNet + greedy CTC decode)"
        step = "log-mel -> encoder (eval) -> decoder -> greedy CTC decode"
    else:
        n_run = n
        cb = oracle_training_throughput(model_name, seconds, labels, n, steps, warmup)
        metric = "train audio-seconds/sec (QuartzNet+CTC)"
        step = "waveforms -> log-mel frontend -> forward + CTC + backward + Novograd update"
    line = {
        "impl": "reference", "metric": metric, "value": cb["value"], "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(args.workload, args.gpus, {
            "step": step, "cuda_graph": False,
            "note": "reference algorithm (oracle port of models/QuartNet*.py + torch CTCLoss + scheduler/novograd.py) on "
                    f"the host cores, {n_run} utterances per step"}),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def _dist_helpers(world):
    import torch
    import torch.distributed as dist

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

    return barrier, max_over_ranks


def measure_infer(workload, steps, warmup, graph, world, rank, local, peaks, sampler=None):
    """Inference workload: waveforms (pinned host) -> H2D -> log-mel -> encoder (eval) -> greedy decode -> tokens D2H."""
    import torch
    from lightning_asr_b200.trainer import InferEngine, LightingModule, synthetic_batch

    model_name, n, seconds, vocab, precision = WORKLOADS[workload]
    labels = labels_for(vocab)
    barrier, max_over_ranks = _dist_helpers(world)
    torch.manual_seed(0)
    module = LightingModule(labels=labels, mask=True, drop_rate=0.0, model_name=model_name,
                            precision=precision).cuda().eval()
    (waves, lens), _, _, _, _ = synthetic_batch(n, seconds, len(labels), seed=1234 + rank, ragged=False, features=False)
    # 16-bit PCM on the wire, as the samples sit in the audio files (predict.py:46 torchaudio.load -> x / 32768)
    engine = InferEngine(module, waves, lens, graph=graph, wave_dtype=torch.int16)
    for _ in range(max(warmup, 3)):
        engine.step_device()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        engine.step_device()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    f0.record()
    engine.prefetch()
    for i in range(steps):
        toks, cnts = engine.step_host(prefetch_next=(i + 1 < steps))
    f1.record()
    barrier()
    e2e_ms = max_over_ranks(max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3))
    clocks = sampler.stop() if sampler is not None else None
    # what the PCIe copy of one batch costs on its own (the e2e loop hides it under the previous pass as long as it is
    # shorter than the pass): explains any gap between `value` and `e2e`
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    h0.record()
    engine.staging.copy_(engine.host, non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_ms_alone = h0.elapsed_time(h1)
    fam, calls, info = kernel_breakdown(engine, steps=3)
    ms_per_step = ms_total / steps
    audio_s = n * seconds * world
    T = 1 + (int(seconds * 16000) + 64) // 160
    Tp = (T - 1) // 2 + 1
    roof = roofline_of(fam, info, ms_per_step, peaks)
    # schedule-L forward bytes: asr13x1 V'=29 fwd = 50 775 elements per encoder step, + the waveform
    sched_bytes = 50775 * 2 * n * Tp + engine.static.element_size() * n * int(seconds * 16000)
    roof["step_hbm_frac_scheduleL"] = sched_bytes / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"]
    res = {
        "metric": "inference audio-seconds/sec (log-mel + QuartzNet + greedy CTC decode)",
        "value": audio_s / (ms_per_step * 1e-3), "unit": "audio-s/s", "n_gpus": world, "steps": steps,
        "warmup": max(warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": precision, "data": "synthetic",
        "config": config_of(workload, world, {
            "parallelism": f"dp{world} (independent replicas, no collective)",
            "step": "H2D 16-bit PCM waveforms -> log-mel -> encoder (eval) -> decoder -> greedy CTC decode -> tokens D2H",
            "cuda_graph": bool(graph), "l2": "no flush needed: each pass streams tens of GB of activations >> 126 MB L2"}),
        "e2e": {"value": audio_s / (e2e_ms / steps * 1e-3), "unit": "audio-s/s", "h2d_bytes_per_step": engine.h2d_bytes,
                "d2h_bytes_per_step": int(toks.numel() * 4 + cnts.numel() * 4), "ms_per_step": e2e_ms / steps,
                "h2d_ms_alone": h2d_ms_alone, "h2d_gbs": engine.h2d_bytes / (h2d_ms_alone * 1e-3) / 1e9},
        "gpu_launches": calls * steps, "roofline": roof, "cpu_baseline": None, "clocks": clocks,
        "tokens_decoded": int(cnts.sum()),
    }
    engine.close()
    del engine, module
    torch.cuda.empty_cache()
    return res


def measure_train(workload, steps, warmup, graph, world, rank, local, peaks, optimizer=True, sampler=None,
                  exposed=False):
    """Training workload -> result dict (rank-local timings are reduced with MAX over ranks)."""
    import torch
    from lightning_asr_b200 import ddp
    from lightning_asr_b200.trainer import LightingModule, TrainEngine, synthetic_batch

    model_name, n, seconds, vocab, precision = WORKLOADS[workload]
    labels = labels_for(vocab)
    barrier, max_over_ranks = _dist_helpers(world)
    torch.manual_seed(0)
    module = LightingModule(labels=labels, mask=True, drop_rate=0.0, model_name=model_name,
                            precision=precision).cuda().train()
    if world > 1:
        ddp.broadcast_parameters(module)
    # the step starts from WAVEFORMS (16-bit PCM on the wire): log-mel frontend -> encoder -> CTC -> backward -> optimizer
    batch = synthetic_batch(n, seconds, len(labels), seed=1234 + rank, ragged=False, features=False)
    engine = TrainEngine(module, batch, graph=graph, fused=True,
                         world_sync=(None, float(os.environ.get("LASR_BUCKET_MB", "8")),
                                     tuple(float(v) for v in os.environ.get("LASR_TAIL_MB", "0.75,2.5,4").split(",") if v))
                         if world > 1 else None,
                         optimizer="novograd" if optimizer else None)
    graph_note = graph
    try:
        for _ in range(max(warmup, 3)):
            engine.step_device()
        torch.cuda.synchronize()
    except Exception as e:  # graph capture refused (e.g. a collective that cannot be captured): eager launches
        if not graph:
            raise
        sys.stderr.write(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); falling back to eager\n")
        torch.cuda.synchronize()
        engine.use_graph, engine.graph, graph_note = False, None, False
        for _ in range(max(warmup, 3)):
            engine.step_device()
        torch.cuda.synchronize()

    def timed_device_steps():
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            engine.step_device()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- device-resident timing (value) ----
    if sampler is not None:
        sampler.start()
    ms_total = timed_device_steps()
    # ---- end-to-end timing (pinned host batch -> H2D -> step -> loss on the host, every step) ----
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    f0.record()
    engine.prefetch()  # H2D of step 0's batch; every later step's H2D is issued inside the loop (one per step)
    for i in range(steps):
        # one H2D of the batch and one D2H of the loss per step; the host reads step i-1's loss while step i runs
        engine.step_host(prefetch_next=(i + 1 < steps), defer_loss=True)
    loss_val = engine.flush_loss()
    f1.record()
    barrier()
    e2e_ms = max_over_ranks(max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3))
    clocks = sampler.stop() if sampler is not None else None

    ms_per_step = ms_total / steps
    audio_s = n * seconds * world
    # ---- how much of the gradient all-reduce is NOT hidden behind backward: same step, exchange switched off ----
    exposed_us = None
    if world > 1 and exposed and engine.grad_sync is not None:
        engine.grad_sync.enabled = False
        engine.graph = None  # re-capture without the collective
        for _ in range(3):
            engine.step_device()
        ms_nosync = timed_device_steps() / steps
        exposed_us = (ms_per_step - ms_nosync) * 1e3
        engine.grad_sync.enabled = True
        engine.graph = None
        for _ in range(3):
            engine.step_device()
        torch.cuda.synchronize()
    # ---- per-kernel breakdown (after the timed regions) ----
    fam, calls, info = kernel_breakdown(engine, steps=3)
    roof = roofline_of(fam, info, ms_per_step, peaks)
    T = 1 + (int(seconds * 16000) + 64) // 160
    Tp = (T - 1) // 2 + 1
    es = 2 if precision == "bf16" else 4
    # schedule-L bytes of the whole step, elements per encoder step
    sched = {("asr13x1", 29): 159726, ("asr13x1contextse", 29): 175534, ("asr13x1context", 29): 175534,
             ("asr13x1context", 4334): 201364}.get((model_name, len(labels) + 1))
    if sched is not None:
        roof["step_hbm_frac_scheduleL"] = sched * es * n * Tp / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"]
    res = {
        "metric": "train audio-seconds/sec (QuartzNet+CTC)", "value": audio_s / (ms_per_step * 1e-3),
        "unit": "audio-s/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": precision, "data": "synthetic",
        "config": config_of(workload, world, {
            "step": "int16 waveforms -> log-mel frontend -> forward + CTC + backward"
                    + (" + NCCL grad all-reduce" if world > 1 else "")
                    + (" + fused Novograd/LR-schedule update" if optimizer else ""),
            "cuda_graph": bool(graph_note),
            "l2": "no flush needed: each step streams GBs of activations >> 126 MB L2"}),
        "e2e": {"value": audio_s / (e2e_ms / steps * 1e-3), "unit": "audio-s/s", "h2d_bytes_per_step": engine.h2d_bytes,
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / steps},
        "gpu_launches": calls * steps, "roofline": roof, "cpu_baseline": None, "clocks": clocks, "loss": loss_val,
    }
    if exposed_us is not None:
        res["allreduce_exposed_us"] = exposed_us
    engine.close()
    del engine, module
    torch.cuda.empty_cache()
    return res


def run_b200(args):
    import torch
    import torch.distributed as dist

    from lightning_asr_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _lib.require_device()
    if world > 1:
        # the gradient exchange is 20-40 MB per step in <= 8 MB buckets: latency-, not bandwidth-bound.  NCCL's CTAs get
        # their own SMs: TrainEngine sizes every persistent kernel for 148 - NCCL_MAX_CTAS SMs (lasr_set_sm_budget), so the
        # compute kernels no longer run a second wave on the SMs the all-reduce occupies.  Measured at N = 2 (exposed
        # all-reduce per step, r3q/r3r sweeps): 2 CTAs 350 us, 4 CTAs 136 us, 8 CTAs 28 us, 12 CTAs 55 us, 16 CTAs 54 us;
        # without the SM budget (round-1 behaviour) 4 CTAs were the optimum at 187 us
        os.environ.setdefault("NCCL_MAX_CTAS", "8")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = load_peaks()
    infer = args.workload.startswith("infer_")
    sampler = ClockSampler(local) if rank == 0 else None
    graph = not args.no_graph
    if infer:
        line = measure_infer(args.workload, args.steps, args.warmup, graph, world, rank, local, peaks, sampler)
    else:
        line = measure_train(args.workload, args.steps, args.warmup, graph, world, rank, local, peaks,
                             optimizer=not args.no_optimizer, sampler=sampler, exposed=True)
    if world > 1:
        # every rank is done with the device; leave together.  The process group is NOT torn down explicitly: with
        # NCCL all-reduces captured in a live CUDA graph destroy_process_group() can block forever, so the ranks
        # simply exit (os._exit below) once rank 0 has printed its line.
        dist.barrier()
        torch.cuda.synchronize()
    if rank != 0:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    model_name, n, seconds, vocab, precision = WORKLOADS[args.workload]
    labels = labels_for(vocab)
    if world == 1 and not args.no_cpu_baseline:
        # reported baselines (rank 0, N = 1 only): the reference's algorithm on the host cores and, the real kernel-level
        # bar, its eager PyTorch path on this same GPU
        if infer:
            cb = oracle_inference_throughput(model_name, seconds, labels, min(n, 16), 2, 1)
        else:
            cb = oracle_training_throughput(model_name, seconds, labels, n, 3, 1)
            line["gpu_eager_baseline"] = gpu_eager_baseline(model_name, seconds, labels, n)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if world == 1 and not args.no_workloads and args.workload == "asr13x1_b32_16s_bf16":
        # the other BASELINE configs as short runs, so the driver's record carries them too (parity cases, not the
        # headline): value / ms_per_step / e2e / dominant kernel family each
        others = {}
        for wl in ("asr13x1_b4_10s_fp32", "contextse_b64_20s_bf16", "context_aishell_b32_16s_bf16",
                   "infer_asr13x1_b256_30s_bf16"):
            try:
                if wl.startswith("infer_"):
                    r = measure_infer(wl, 5, 3, graph, 1, 0, local, peaks)
                else:
                    r = measure_train(wl, 8, 3, graph, 1, 0, local, peaks, optimizer=True)
                others[wl] = {"value": r["value"], "unit": r["unit"], "ms_per_step": r["ms_per_step"], "steps": r["steps"],
                              "dtype": r["dtype"], "e2e": r["e2e"]["value"], "gpu_launches": r["gpu_launches"],
                              "config": {k: r["config"][k] for k in ("model_name", "per_gpu_batch", "seconds", "vocab")},
                              "step_hbm_frac_scheduleL": r["roofline"].get("step_hbm_frac_scheduleL"),
                              "dominant": {k: r["roofline"][k] for k in ("kernel", "bound", "achieved", "unit", "frac",
                                                                         "share_of_step")}}
            except Exception as e:
                others[wl] = {"error": f"{type(e).__name__}: {e}"[:300]}
                torch.cuda.empty_cache()
        line["workloads"] = others
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
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the reported CPU / GPU-eager baselines")
    ap.add_argument("--no-workloads", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--no-optimizer", action="store_true", help="time fwd+bwd only (diagnostics; not the headline)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

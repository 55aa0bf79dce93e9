"""Minimal driver for ncu: a few eager training steps of a bench workload (no timing, no CPU baseline)."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from lightning_asr_b200.trainer import LightingModule, TrainEngine, synthetic_batch
wl = sys.argv[1] if len(sys.argv) > 1 else "asr13x1_b32_16s_bf16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model_name, n, seconds, vocab, precision = bench.WORKLOADS[wl]
labels = bench.labels_for(vocab)
torch.manual_seed(0)
mod = LightingModule(labels=labels, mask=True, model_name=model_name, precision=precision).cuda().train()
eng = TrainEngine(mod, synthetic_batch(n, seconds, len(labels), seed=1234), graph=False, optimizer="novograd")
for _ in range(steps):
    eng.step_device()
torch.cuda.synchronize()
print("loss", float(eng.loss_dev))

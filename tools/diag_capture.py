"""Find which C-ABI call (if any) invalidates a CUDA-graph capture of the training step."""
import ctypes, sys, traceback, glob, os
import torch
sys.path.insert(0, ".")
from lightning_asr_b200 import _lib
from lightning_asr_b200.trainer import LightingModule, TrainEngine, synthetic_batch
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = ctypes.CDLL(cands[0])
def status():
    st = ctypes.c_int(-1)
    rc = rt.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    return rc, st.value
orig = _lib.call
bad = []
def call(name, *a):
    r = orig(name, *a)
    rc, st = status()
    if (rc != 0 or st == 2) and not bad:
        bad.append(name); print("capture invalidated after", name, "rc", rc, "status", st, flush=True)
    return r
_lib.call = call
import lightning_asr_b200.ops as ops, lightning_asr_b200.runtime as runtime
ops.call = call
labels = [" ", "'"] + [chr(ord("a") + i) for i in range(26)]
n, sec = (int(sys.argv[1]), float(sys.argv[2])) if len(sys.argv) > 2 else (4, 2.0)
mod = LightingModule(labels=labels, mask=True, precision="bf16").cuda().train()
batch = synthetic_batch(n, sec, 28)
eng = TrainEngine(mod, batch, graph=True)
try:
    eng.step_device(); torch.cuda.synchronize(); print("capture ok; loss", float(eng.loss_dev))
    eng.step_device(); torch.cuda.synchronize(); print("replay ok; loss", float(eng.loss_dev))
except Exception:
    traceback.print_exc()

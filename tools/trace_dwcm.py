"""Timeline of the TMA-fed depthwise kernels (debug hook in dwconv_cm.cu): %globaltimer stamps per role and item."""
import ctypes
import sys
import torch
sys.path.insert(0, ".")
from lightning_asr_b200 import _lib, ops
_lib.require_device()
lib = _lib.load()
lib.lasr_debug_set_dw_trace.argtypes = [ctypes.c_void_p]
lib.lasr_debug_set_dw_trace.restype = None
N, T = 32, 801
names = ["tma:refill", "-", "mma:acc_free", "mma:series_full", "mma:issued", "epi:tmem_full", "epi:done",
         "cta:entry/prologue_done/roles_done"]
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
for c, k in [(512, 63), (256, 33)]:
    x = torch.randn(N, T, c, device="cuda").bfloat16()
    w = torch.randn(c, 1, k, device="cuda") / k ** 0.5
    xs = ops.series_from_ntc(x, k)
    add = ops.series_from_ntc(torch.randn_like(x), k) if which in ("dgrad", "bwd") else None
    dwb = torch.zeros(c, 1, k, device="cuda")
    if which == "bwd":
        run = lambda: ops.dwconv_bwd_cm(xs, add, w, addend=add, out_dw=dwb)
    elif which == "wgrad":
        add = ops.series_from_ntc(torch.randn_like(x), k)
        run = lambda: ops.dwconv_wgrad_cm(xs, add, k, out=dwb)
    else:
        run = lambda: ops.dwconv_fwd_cm(xs, w, flip=which == "dgrad", addend=add)
    run()
    trace = torch.zeros(148 * 8 * 16, device="cuda", dtype=torch.int64)
    torch.cuda.synchronize()
    lib.lasr_debug_set_dw_trace(trace.data_ptr())
    run()
    torch.cuda.synchronize()
    lib.lasr_debug_set_dw_trace(None)
    tr = trace.cpu().view(148, 8, 16)
    t0 = int(tr[:, 7, 0][tr[:, 7, 0] > 0].min())
    print(f"=== {which} C={c} k={k}: last CTA done {int(tr[:, 7, 2].max()) - t0} ns; CTA entry spread "
          f"{int(tr[:, 7, 0][tr[:, 7, 0] > 0].max()) - t0} ns")
    for cta in (0, 77):
        print(f"-- CTA {cta}")
        for slot, nm in enumerate(names):
            vals = [int(v) - t0 for v in tr[cta, slot] if int(v) > 0]
            print(f"   {nm:18s} {vals}")

"""Two CTC lattice launches at the config-2 shape for ncu (tools/prof_ctc.py [0|1]: LASR_CTC_WARP).
ncu --set full --import-source on -k regex:ctc_lattice -c 2 python tools/prof_ctc.py 1"""
import os
import sys

import torch

sys.path.insert(0, ".")
os.environ["LASR_CTC_WARP"] = sys.argv[1] if len(sys.argv) > 1 else "1"
from lightning_asr_b200 import _lib, ops  # noqa: E402

_lib.require_device()
N, T, V, ld = 32, 801, 29, 32
S = T // 4
torch.manual_seed(0)
logits = torch.randn(N, T, ld, device="cuda").bfloat16()
targets = torch.randint(0, 28, (N, S), device="cuda")
il = torch.full((N,), T, device="cuda", dtype=torch.int32)
tl = torch.full((N,), S, device="cuda", dtype=torch.int32)
lse, _ = ops.log_softmax_fwd(logits, V, want_lp=False)
for _ in range(2):
    ops.ctc_fwd(logits, lse, targets, il, tl, V, 28, want_beta=True)
torch.cuda.synchronize()

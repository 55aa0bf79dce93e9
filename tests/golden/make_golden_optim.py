"""Golden fixture for the optimizer path: runs the REFERENCE's own scheduler/novograd.py::Novograd and
scheduler/cosine_annearing_with_warmup.py::CosineAnnealingWarmupRestarts (imported from /root/reference, build
container only) on seeded parameters / gradients and stores the per-step results in tests/golden/optim.pt.

    python tests/golden/make_golden_optim.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

from golden_common import optim_case  # noqa: E402
from scheduler.cosine_annearing_with_warmup import CosineAnnealingWarmupRestarts  # noqa: E402
from scheduler.novograd import Novograd  # noqa: E402


def main():
    fx = {}
    for name, (hyper, sched, steps) in optim_case.CASES.items():
        params = [torch.nn.Parameter(p.clone()) for p in optim_case.params()]
        opt = Novograd(params, **hyper)
        sch = CosineAnnealingWarmupRestarts(opt, **sched) if sched is not None else None
        lrs, snaps = [], []
        for k in range(steps):
            for p, g in zip(params, optim_case.grads(k)):
                p.grad = g.clone()
            lrs.append(float(opt.param_groups[0]["lr"]))
            opt.step()
            if sch is not None:
                sch.step()
            if k in optim_case.SNAP_STEPS or k == steps - 1:
                snaps.append((k, [p.detach().clone() for p in params]))
        fx[name] = {
            "lrs": lrs,
            "snaps": snaps,
            "exp_avg_sq": [float(opt.state[p]["exp_avg_sq"]) for p in params],
            "exp_avg": [opt.state[p]["exp_avg"].clone() for p in params],
        }
        print(name, "lrs[:6]", lrs[:6], "final |p0|", float(params[0].norm()))
    torch.save(fx, os.path.join(HERE, "optim.pt"))


if __name__ == "__main__":
    main()

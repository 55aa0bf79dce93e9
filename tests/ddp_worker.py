"""Worker of tests/test_ddp_gpu.py (one process per GPU, launched by torch.distributed.run; NCCL).

Checks the gradient exchange the scaling bench times (ddp.GradSync over the ParamBank's flat gradient buffer, buckets
fired from runtime.grad_ready() on the comm stream, all inside the step's CUDA graph) -- conf/conf.yaml:30
`accelerator: ddp`:
  1. after one TrainEngine step with world_sync, every rank's gradient buffer equals the AVERAGE of the ranks'
     independent single-GPU gradients (each rank has its own batch);
  2. with the fused Novograd in the step, the parameters are bit-identical on all ranks after two steps;
  3. BatchNorm running statistics stay per-rank (the reference has no SyncBN).
Prints 'DDP_WORKER_OK' on rank 0."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    from lightning_asr_b200 import ddp, runtime
    from lightning_asr_b200.trainer import LightingModule, TrainEngine, synthetic_batch

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    labels = [" ", "'"] + [chr(ord("a") + i) for i in range(26)]
    precision = os.environ.get("LASR_TEST_PRECISION", "bf16")
    graph = os.environ.get("LASR_TEST_GRAPH", "1") == "1"
    torch.manual_seed(100 + rank)  # deliberately different initial weights: the broadcast must fix that
    mod = LightingModule(labels=labels, mask=True, precision=precision).cuda().train()
    ddp.broadcast_parameters(mod)
    sd0 = copy.deepcopy(mod.state_dict())
    batch = synthetic_batch(3, 2.0, len(labels), seed=50 + rank, ragged=True)

    # (a) this rank's own gradients, no exchange
    try:
        eng = TrainEngine(mod, batch, graph=False, optimizer=None)
        eng.step_host()
        local_g = eng.bank.grads.clone()
        rm_local = mod.encoder.encoder.first_cnn.bn.running_mean.clone()
    finally:
        runtime.uninstall()
    gathered = [torch.empty_like(local_g) for _ in range(world)]
    dist.all_gather(gathered, local_g)
    expect = torch.stack(gathered).double().mean(0)
    assert (gathered[0] - gathered[1]).abs().max() > 0  # the ranks really saw different batches

    # (b) the exchanged gradients
    mod2 = LightingModule(labels=labels, mask=True, precision=precision).cuda().train()
    mod2.load_state_dict(sd0)
    try:
        eng2 = TrainEngine(mod2, batch, graph=graph, optimizer=None, world_sync=(None, 2.0))
        assert len(eng2.grad_sync.buckets) >= 3
        eng2.step_host()
        eng2.step_host()  # replay: buckets re-armed, same result
        got = eng2.bank.grads.clone().double()
    finally:
        runtime.uninstall()
    err = float((got - expect).norm() / expect.norm())
    tol = 1e-5 if precision == "fp32" else 2e-3  # split-K / RED accumulation order differs run to run in bf16
    assert err < tol, (rank, err)
    allg = [torch.empty_like(got) for _ in range(world)]
    dist.all_gather(allg, got)
    assert torch.equal(allg[0], allg[1])  # NCCL leaves the same bits on every rank
    rm2 = mod2.encoder.encoder.first_cnn.bn.running_mean.clone()

    # (c) optimizer in the step: parameters stay in lock-step
    mod3 = LightingModule(labels=labels, mask=True, precision=precision).cuda().train()
    mod3.load_state_dict(sd0)
    try:
        eng3 = TrainEngine(mod3, batch, graph=graph, optimizer="novograd", world_sync=(None, 2.0))
        eng3.step_host()
        eng3.step_host()
        params = eng3.bank.master.clone()
    finally:
        runtime.uninstall()
    allp = [torch.empty_like(params) for _ in range(world)]
    dist.all_gather(allp, params)
    assert torch.equal(allp[0], allp[1])
    assert not torch.equal(params, eng2.bank.master)  # and they really moved
    # BatchNorm statistics are per rank (different batches -> different running means), no SyncBN
    rms = [torch.empty_like(rm_local) for _ in range(world)]
    dist.all_gather(rms, rm2)
    assert (rms[0] - rms[1]).abs().max() > 0
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"DDP_WORKER_OK grad_err={err:.3e}", flush=True)
    os._exit(0)  # NCCL captured in a live CUDA graph: do not tear the process group down (see bench.py)


if __name__ == "__main__":
    main()

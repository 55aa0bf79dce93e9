"""CPU-side checks: the C-ABI library loads and exports every symbol include/lasr.h declares (no compute calls),
the product never routes through the oracle or a CPU fallback, the drop-in modules keep the reference's
constructor / state_dict schema, and the data-parallel gradient exchange works over gloo with world_size 2."""
import ctypes
import os
import re
import socket
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)

from golden_common import LABELS28, golden_weights  # noqa: E402


def test_abi_library_exports_header_symbols():
    from lightning_asr_b200 import _lib

    header = open(os.path.join(ROOT, "include", "lasr.h")).read()
    declared = set(re.findall(r"\b(lasr_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", header, flags=re.S)))
    declared -= {"lasr_stream_t"}
    assert len(declared) >= 25
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()  # raises if the .so has not been built: there is no fallback
    for name in declared:
        assert isinstance(getattr(lib, name), ctypes._CFuncPtr), name
    assert lib.lasr_abi_version() == _lib.ABI_VERSION
    assert lib.lasr_strerror(0) == b"ok" and b"shape" in lib.lasr_strerror(-1)
    # pure host helpers (no device needed)
    assert lib.lasr_bn_bwd_chunks(32, 801) >= 1


def test_no_cpu_fallback_and_no_oracle_in_product():
    pkg = os.path.join(ROOT, "lightning-asr_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
            assert "/root/reference" not in src, fn
    from lightning_asr_b200 import _lib, ops

    x = torch.zeros(1, 4, 8)
    with pytest.raises(_lib.LasrError):
        ops.nct_to_ntc(x, torch.float32)  # CPU tensors are refused: the product path has no CPU route
    if not torch.cuda.is_available():
        with pytest.raises(_lib.LasrError):
            _lib.require_device()


@pytest.mark.parametrize("variant,cls", [("base", "MyModel2"), ("context", "MyModel2Context"),
                                         ("contextse", "MyModel2ContextSE")])
def test_state_dict_schema_matches_reference(variant, cls):
    """The golden fixture stores the reference module's state_dict keys / shapes: checkpoints must interchange."""
    import lightning_asr_b200.quartznet as q

    fx = torch.load(os.path.join(GOLDEN, f"model_{variant}.pt"), weights_only=False)
    model = getattr(q, cls)(LABELS28, drop_rate=0.0, mask=True)
    ours = [(k, tuple(v.shape), str(v.dtype)) for k, v in model.state_dict().items()]
    assert ours == [(k, tuple(s), d) for k, s, d in fx["schema"]]
    model.load_state_dict(golden_weights(fx["schema"]), strict=True)
    assert q.build_model({"base": "asr13x1", "context": "asr13x1_context", "contextse": "QuartNetContextSE"}[variant],
                         LABELS28).__class__ is getattr(q, cls)
    with pytest.raises(KeyError):
        q.build_model("no_such_model", LABELS28)


def test_default_init_matches_reference_order():
    """Same torch seed -> same default initialisation as the reference (fixture keeps a checksum per variant)."""
    import lightning_asr_b200.quartznet as q

    torch.manual_seed(0)
    m = q.MyModel2(LABELS28, mask=True)
    n_params = sum(p.numel() for p in m.parameters())
    assert n_params == 5045597  # asr13x1 parameter count


def test_synthetic_batch_contract():
    from lightning_asr_b200.trainer import num_frames, synthetic_batch

    assert [num_frames(16000 * s) for s in (10, 16, 20, 30)] == [1001, 1601, 2001, 3001]
    x, targets, percents, sizes, paths = synthetic_batch(4, 1.0, 28, ragged=True)
    assert x.shape == (4, 1, 64, 101) and targets.dtype == torch.int64 and sizes.dtype == torch.int32
    assert float(percents.max()) == 1.0 and len(paths) == 4
    Tp = 51
    t_len = torch.mul(Tp, percents).int()
    assert bool(((2 * sizes + 1) <= t_len).all())  # feasible CTC alignments


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ddp_worker(rank, world, port, outdir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lightning_asr_b200 import ddp

    torch.manual_seed(rank)  # different initial parameters per rank -> broadcast must equalise them
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    ddp.broadcast_parameters(net)
    out = {}
    for overlap in (False, True):
        sync = ddp.GradSync(net, bucket_mb=1e-4, overlap=overlap)  # tiny buckets -> several of them
        assert len(sync.buckets) > 1
        sync.zero_and_attach()
        g = torch.Generator().manual_seed(100 + rank)
        x = torch.randn(6, 7, generator=g)
        net(x).square().mean().backward()
        local = [p.grad.clone() for p in net.parameters()]
        sync(net)
        out[overlap] = ([p.grad.clone() for p in net.parameters()], local)
    torch.save(([p.detach().clone() for p in net.parameters()], out[False], out[True]),
               os.path.join(outdir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_sync_gloo_world2(tmp_path):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = {r: torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"), weights_only=False) for r in range(2)}
    # identical parameters on every rank after the broadcast
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b)
    for mode in (1, 2):
        (avg0, loc0), (avg1, loc1) = res[0][mode], res[1][mode]
        for a0, a1, l0, l1 in zip(avg0, avg1, loc0, loc1):
            assert torch.allclose(a0, a1)  # every rank ends the step with the same gradient ...
            assert torch.allclose(a0, (l0 + l1) / 2, atol=1e-7)  # ... the average of the per-rank gradients


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the reference's CPU algorithm on the host cores) prints ONE JSON line with the
    contract's keys; runs on any machine (no GPU)."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload",
                          "asr13x1_b4_10s_fp32", "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("train audio-seconds/sec") and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"] == "asr13x1_b4_10s_fp32" and d["config"]["encoder_steps"] == 501
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_bench_reference_arm_under_torchrun_world2():
    """Under torchrun (N > 1) rank 0 alone runs and prints the reference arm; the other ranks exit 0 without work."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(root, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--workload", "asr13x1_b4_10s_fp32", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0


def test_series_layout_helpers_hold_their_invariants():
    """lasr_cm_offset / lasr_cm_pitch / lasr_cm_ks are host functions (no GPU): the invariants the TMA-fed depthwise
    kernels and the pre-swizzled global layout rely on (include/lasr.h, "channel-major series")"""
    from lightning_asr_b200 import _lib
    lib = _lib.load()
    for K in range(3, 91, 2):
        off = lib.lasr_cm_offset(K)
        ks = lib.lasr_cm_ks(K)
        assert off % 8 == 0 and K // 2 <= off < K // 2 + 8          # left padding rounded up to one 16-byte group
        assert ks % 16 == 0 and K + 15 + (off - K // 2) <= ks <= 112  # Toeplitz factor covers taps + 15 + the remainder
        for T in (1, 7, 40, 157, 801, 895, 896, 897, 1501, 1792, 1793, 3000, 12000):
            S = lib.lasr_cm_pitch(T, K)
            assert S % 128 == 0                       # 256-byte rows: the swizzle phase of a block is its own
            assert S >= off + T + 48                  # zero tail >= the weight gradient's backward reach
            chunks = -(-T // 896)
            if chunks > 1:                            # every 1024-position slot of every 896-output chunk lies in the row
                assert S >= 896 * (chunks - 1) + 1024
    # one group index permutation for the whole layout: an involution that only swaps neighbours in the upper half of a
    # 128-position block
    for g in range(64):
        gs = g ^ ((g >> 3) & 1)
        assert (gs ^ ((gs >> 3) & 1)) == g and gs // 2 == g // 2


def test_flat_batch_views_alias_one_buffer():
    """TrainEngine stages a batch as ONE flat buffer (one H2D copy per step): every tensor of the batch is a view at a
    256-byte boundary with its own dtype and shape, writes through a view land in the flat buffer, non-tensor entries
    pass through, and the wire formats keep their widths (16-bit PCM stays 2 bytes per sample)."""
    from lightning_asr_b200.trainer import _flat_views

    waves = torch.randint(-32768, 32767, (3, 1001), dtype=torch.int16)
    targets = torch.randint(0, 28, (3, 7))
    lens = torch.tensor([1001, 900, 77], dtype=torch.int32)
    tsz = torch.tensor([7, 5, 1], dtype=torch.int32)
    flat, views = _flat_views([waves, targets, None, lens, tsz], "cpu")
    assert flat.dtype == torch.uint8 and views[2] is None
    payload = waves.numel() * 2 + targets.numel() * 8 + 3 * 4 + 3 * 4
    assert payload <= flat.numel() < payload + 4 * 256
    base = flat.data_ptr()
    last_end = 0
    for v, t in zip(views, [waves, targets, None, lens, tsz]):
        if t is None:
            continue
        assert v.dtype == t.dtype and v.shape == t.shape and v.is_contiguous()
        off = v.data_ptr() - base
        assert off % 256 == 0 and off >= last_end
        last_end = off + v.numel() * v.element_size()
        v.copy_(t)
    assert last_end <= flat.numel()
    # the views alias the flat buffer: a byte-level copy of it carries the whole batch
    clone = flat.clone()
    off_t = views[1].data_ptr() - base
    again = clone[off_t:off_t + targets.numel() * 8].view(torch.int64).view(targets.shape)
    assert torch.equal(again, targets)
    off_w = views[0].data_ptr() - base
    assert torch.equal(clone[off_w:off_w + waves.numel() * 2].view(torch.int16).view(waves.shape), waves)

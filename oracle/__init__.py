"""CPU oracle for the lightning-asr training hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (plain PyTorch fp32/fp64 + numpy) of the reference
algorithm for the path BASELINE.json's north_star names: log-mel frontend -> QuartzNet-style
encoder (3 variants) -> log-softmax -> CTC loss / greedy CTC decode.

Rules (enforced by tests/test_host_cpu.py::test_no_cpu_fallback_and_no_oracle_in_product):
  * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
    legs may import anything from here;
  * the product package (lightning_asr_b200) never imports it and has no CPU fallback.

Pinning: the reference ships NO tests, golden vectors or fixtures for this path
(SURVEY.md section 4 / 8c), so the oracle is pinned against *outputs of the reference itself*:
tests/golden/make_golden.py imports the reference's own modules from /root/reference in the
build container, runs them on seeded inputs and commits small input/output fixtures under
tests/golden/; tests/test_oracle_golden.py checks every function here against those.
The parts of the reference that cannot be imported (data_module.py, utils/asr_metrics.py,
train.py need pytorch_lightning/hydra/torchmetrics/editdistance) are restated from the
cited lines and pinned against torchaudio / torch.nn.CTCLoss run in the build container.
"""

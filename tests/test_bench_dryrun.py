"""bench.py's own arm, dry on CPU: ``run_ours`` with the kernel wrappers emulated (tests/emu_ops.py), CUDA
events / pinned memory / device selection replaced by host stand-ins, a reduced model and tiny workloads.
Nothing is measured; this proves that the default line, the batched workload (cfg4) and the rollout mode
run end to end through the public API and that the JSON line carries every key of the contract."""
import json

import pytest
import torch

class FakeEvent:
    def __init__(self, enable_timing=True): self.t = 0.0
    def record(self):
        import time; self.t = time.perf_counter()
    def elapsed_time(self, other): return (other.t - self.t) * 1e3

def _run(monkeypatch, capsys, argv, workload):
    import bench
    class ProxyTorch:
        def __getattr__(self, k): return getattr(torch, k)
        @staticmethod
        def device(*a, **k): return torch.device("cpu")
    monkeypatch.setattr(bench, "torch", ProxyTorch())
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(bench, "WORKLOADS", {**bench.WORKLOADS, **workload})
    monkeypatch.setattr(bench, "BATCH_VIDEOS", {"cfg4": 4})
    def build_model(dims, dev):
        from transformers import Phi3Config
        from videogpt_b200 import LVM
        m = LVM(Phi3Config(**dims.phi3_kwargs()), device="cpu", materialize_pos_embed=False)
        with torch.no_grad():
            for lin in (m.final_layer.linear, getattr(m.final_layer.adaLN_modulation, "1")):
                lin.weight.normal_(std=0.02)
        return m.float().eval()
    monkeypatch.setattr(bench, "build_model", build_model)
    monkeypatch.setattr(bench, "gemm_roofline", lambda model, rows, reps=3: (1e-4, 1e9, 12))
    monkeypatch.setattr(bench, "cpu_reference_sample", lambda *a, **k: (100.0, "stub", 1.0))
    monkeypatch.setattr(bench, "gpu_eager_oracle", lambda *a, **k: 0.05)
    # pipeline casts the model to bf16 by default: keep fp32 for the emulation
    from videogpt_b200 import pipeline
    for name in ("next_clip_latents", "next_clip_latents_batch", "rollout_latents"):
        fn = getattr(pipeline.LVMPipeline, name)
        monkeypatch.setattr(pipeline.LVMPipeline, name, (lambda f: lambda self, *a, **k: f(self, *a, **{**k, "dtype": torch.float32}))(fn))
    args = bench.argparse.Namespace(**argv)
    line = bench.run_ours(args, 0, 1, 0)
    if argv.get("strong_single"):
        line["strong_scaling"] = bench.strong_scaling_single(argv["strong_single"])
    return json.loads(json.dumps(line))        # must be JSON-serialisable

BASE = dict(gpus=1, steps=1, warmup=1, impl="ours", parallelism="dp", sp=0, batch=2, videos=0, rollout=0, recompute=False,
            no_baselines=False, strong="none", strong_timeout=10)

def test_default(emu, monkeypatch, capsys):
    line = _run(monkeypatch, capsys, dict(BASE, config="cfg2"), {"cfg2": ("reduced", 2, 2, 64, 64, 2)})
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["config"]["workload"] == "cfg2" and line["scaling"] == "weak" and line["e2e"]["h2d_bytes_per_step"] > 0

def test_strong_scaling_single_gpu_block(emu, monkeypatch, capsys):
    monkeypatch.setattr("bench._dims", lambda kind: __import__("videogpt_b200").synth.REDUCED)
    line = _run(monkeypatch, capsys, dict(BASE, config="cfg2", strong_single=["cfg5", "cfg3"]),
                {"cfg2": ("reduced", 2, 2, 64, 64, 2), "cfg5": ("reduced", 1, 1, 64, 96, 2), "cfg3": ("reduced", 3, 1, 64, 64, 2)})
    for c in ("cfg5", "cfg3"):
        assert "error" not in line["strong_scaling"][c], line["strong_scaling"][c]
        assert line["strong_scaling"][c]["s_per_clip"] > 0

def test_cfg4(emu, monkeypatch, capsys):
    line = _run(monkeypatch, capsys, dict(BASE, config="cfg4"), {"cfg4": ("reduced", 2, 2, 64, 64, 2)})
    assert line["scaling"] == "strong" and "videos" in line["config"]["parallelism"]
    assert line["gpu_launches"] > 0

@pytest.mark.parametrize("recompute", [False, True])
def test_rollout(emu, monkeypatch, capsys, recompute):
    line = _run(monkeypatch, capsys, dict(BASE, config="cfg3", rollout=2, recompute=recompute), {"cfg3": ("reduced", 3, 2, 64, 64, 2)})
    assert line["config"]["rollout"]["rounds"] == 2
    assert ("recomputed" in line["config"]["rollout"]["context"]) == recompute

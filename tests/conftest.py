import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def emu(monkeypatch):
    """The product's host side (model / scheduler / engine / pipeline) on CPU in fp32, with every
    kernel wrapper replaced by its torch emulation (tests/emu_ops.py) -- for numerical checks of
    plan construction and caching against the oracle on machines without a GPU."""
    import torch
    import emu_ops
    from videogpt_b200 import engine, model, scheduler
    for mod in (engine, model, scheduler):
        monkeypatch.setattr(mod, "ops", emu_ops)
    monkeypatch.setattr(engine, "ACT_DTYPE", torch.float32)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)

    def engine_cpu(self):
        dev = torch.device("cpu")
        key = tuple(p._version for p in self.parameters())
        if self._engine is None or self._engine_key != key:
            d = self.dims()
            w = engine.EngineWeights(self.state_dict(), d.num_hidden_layers, dev)
            self._engine = engine.NextClipEngine(w, d.hidden_size, d.intermediate_size, d.num_hidden_layers,
                                                 d.num_attention_heads, d.rms_norm_eps, d.rope_theta, dev,
                                                 self.pos_embed_max_size, self.patch_size, use_cuda_graph=False)
            self._engine_key, self._plan_key, self._layout_key = key, None, None
        return self._engine
    monkeypatch.setattr(model.LVM, "engine", engine_cpu)
    emu_ops.calls.clear()
    emu_ops._peer_bufs.clear()
    return emu_ops

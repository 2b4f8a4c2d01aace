"""Shared test helpers: fake tokenizer, prompts, oracle/product input builders."""
import hashlib
import json
import os
import re

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class FakeTokenizer:
    """One id per tag after BOS; the released tokenizer is not available."""
    TAGS = {"<img>": 32001, "</img>": 32002, "<|diffusion|>": 32003}
    eos_token_id = 2

    class _Out:
        def __init__(self, ids):
            self.input_ids = ids

    def __call__(self, text):
        ids = [1]
        for t in re.findall(r"<img>|</img>|<\|diffusion\|>", text):
            ids.append(self.TAGS[t])
        return self._Out(ids)


def h16(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()[:16]


def prompts(n_ctx, n_gen):
    p = "".join(f"<img><|image_{i + 1}|></img>" for i in range(n_ctx))
    p += "".join(f"<|diffusion|><|image_{n_ctx + i + 1}|>" for i in range(n_gen))
    p_ = "".join(f"<|diffusion|><|image_{i + 1}|>" for i in range(n_gen))
    return p, p_


def processor_golden():
    with open(os.path.join(GOLDEN, "processor_golden.json")) as f:
        return json.load(f)


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosine(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


class FakeVAE:
    """Stands in for diffusers' AutoencoderKL (outside the hot path): 8x8 average pooling of the
    three colour channels into 4 latent channels and its nearest-neighbour inverse; records what
    it is asked to decode."""
    class config:
        shift_factor = None
        scaling_factor = 0.5

    def __init__(self):
        self.decoded = []

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def encode(self, x):
        lat = torch.nn.functional.avg_pool2d(x, 8)
        lat = torch.cat([lat, lat.mean(1, keepdim=True)], 1)

        class _D:
            class latent_dist:
                @staticmethod
                def sample():
                    return lat.clone()
        return _D

    def decode(self, lat):
        self.decoded.append(lat.clone())
        img = torch.nn.functional.interpolate(lat[:, :3], scale_factor=8, mode="nearest")

        class _S:
            sample = img
        return _S


def random_pil(seed, size=64):
    import numpy as np
    from PIL import Image
    rng = np.random.default_rng(seed)
    return Image.fromarray(rng.integers(0, 256, (size, size, 3), dtype=np.uint8))

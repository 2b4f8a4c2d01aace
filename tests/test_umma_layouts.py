"""Pins the shared-memory layouts / tcgen05 descriptor encodings the production kernels use,
with the probe entry point (vgpt_debug_umma_probe): raw smem images built on the host + the
descriptor fields -> the accumulator tile must equal A @ B^T exactly (small integers)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def smem_desc(lbo_bytes, sbo_bytes, layout):
    return ((lbo_bytes >> 4) << 16) | ((sbo_bytes >> 4) << 32) | (1 << 46) | (layout << 61)


def idesc(m, n, a_mn=0, b_mn=0):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def kmajor_image(mat: torch.Tensor, swizzle_bytes: int) -> torch.Tensor:
    """[rows, k] bf16 -> bytes of the K-major swizzled tile TMA would write.  Rows are
    `swizzle_bytes` wide and dense; the hardware swizzle XORs byte-address bits [4,7) with bits
    [7,10), masked to the swizzle width: addr ^= ((addr >> 7) & (swizzle_bytes/16 - 1)) << 4."""
    rows, k = mat.shape
    assert k * 2 == swizzle_bytes
    chunks = swizzle_bytes // 16
    raw = mat.contiguous().view(torch.uint8).view(rows * chunks, 16)
    addr = torch.arange(rows * chunks) * 16
    dst = addr ^ (((addr >> 7) & (chunks - 1)) << 4)
    out = torch.empty_like(raw)
    out[dst // 16] = raw
    return out.reshape(-1)


def ints(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-4, 5, shape, generator=g).to(torch.bfloat16)


@pytest.mark.parametrize("n", [128, 256])
def test_kmajor_sw128_gemm_layout(n):
    """The GEMM's operand layout: 64-element (128 B) rows, SBO = 1024 B, +32 B per K=16 step."""
    from videogpt_b200 import ops
    a, b = ints((128, 64), 1), ints((n, 64), 2)
    d = ops.umma_probe(kmajor_image(a, 128).to(DEV), kmajor_image(b, 128).to(DEV),
                       smem_desc(16, 1024, 2), smem_desc(16, 1024, 2), idesc(128, n), 4, 32, 32, n)
    assert torch.equal(d.cpu(), a.float() @ b.float().t())


def test_kmajor_sw64_layout():
    """32-element (64 B) rows, SBO = 512 B: the head_dim-96 = 3 x 32 split of attention Q/K."""
    from videogpt_b200 import ops
    a, b = ints((128, 32), 3), ints((128, 32), 4)
    d = ops.umma_probe(kmajor_image(a, 64).to(DEV), kmajor_image(b, 64).to(DEV),
                       smem_desc(16, 512, 4), smem_desc(16, 512, 4), idesc(128, 128), 2, 32, 32, 128)
    assert torch.equal(d.cpu(), a.float() @ b.float().t())


@pytest.mark.parametrize("d,row_bytes,layout", [(96, 64, 4), (64, 128, 2), (128, 128, 2)])
def test_mn_major_b_operand_is_v_as_stored(d, row_bytes, layout):
    """O = P V with V kept as stored in the KV cache, [keys][d] (d contiguous) = an MN-major B
    operand: chunks of row_bytes/2 d-columns (LBO = chunk stride), 8-key groups (SBO = 8 rows),
    16 keys per MMA K-step (= 16 rows of start-address advance)."""
    from videogpt_b200 import ops
    keys = 64
    p, v = ints((128, keys), 5), ints((keys, d), 6)
    cw = row_bytes // 2
    v_img = torch.cat([kmajor_image(v[:, c * cw:(c + 1) * cw].contiguous(), row_bytes) for c in range(d // cw)])
    out = ops.umma_probe(kmajor_image(p, 128).to(DEV), v_img.to(DEV), smem_desc(16, 1024, 2),
                         smem_desc(keys * row_bytes, 8 * row_bytes, layout), idesc(128, d, b_mn=1),
                         keys // 16, 32, 16 * row_bytes, d)
    assert torch.equal(out.cpu(), p.float() @ v.float())


@pytest.mark.parametrize("d,row_bytes,layout", [(96, 64, 4), (64, 128, 2)])
def test_a_operand_in_tensor_memory(d, row_bytes, layout):
    """O = P V with P read from TMEM (lane = row, 32-bit column c = elements 2c | 2c+1 << 16) and V
    MN-major from shared memory: the layout the attention kernel writes P in."""
    from videogpt_b200 import ops
    keys = 64
    p, v = ints((128, keys), 7), ints((keys, d), 8)
    words = p.contiguous().view(torch.int16).to(torch.int32) & 0xFFFF
    a_words = (words[:, 0::2] | (words[:, 1::2] << 16)).to(torch.int32).contiguous()
    cw = row_bytes // 2
    v_img = torch.cat([kmajor_image(v[:, c * cw:(c + 1) * cw].contiguous(), row_bytes) for c in range(d // cw)])
    out = ops.umma_probe_ts(a_words.to(DEV), v_img.to(DEV), smem_desc(keys * row_bytes, 8 * row_bytes, layout),
                            idesc(128, d, b_mn=1), keys // 16, 16 * row_bytes, d)
    assert torch.equal(out.cpu(), p.float() @ v.float())

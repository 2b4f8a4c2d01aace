"""Index / mask construction: product closed form and oracle restatement vs the golden
fixtures generated from the reference (tests/golden/make_golden.py).  Bit-exact."""
import pytest
import torch

from oracle import processor_oracle as po
from videogpt_b200.processor import (FrameGeometry, LVMCollator, LVMProcessor, frame_block_mask,
                                     frame_block_positions)

from helpers import FakeTokenizer, h16, processor_golden, prompts

GOLD = processor_golden()
FB = [g for g in GOLD if g["path"] == "frame_block"]
SF = [g for g in GOLD if g["path"] == "single_frame"]


def _norm(d):
    return {str(k): v for k, v in d.items()}


def _check_frame_block(d, g):
    assert d["attention_mask"].dtype == torch.bool and g["mask_dtype"] == "torch.bool"
    assert d["attention_mask"].shape[-1] == g["L"]
    assert int(d["attention_mask"].sum()) == g["mask_ones"]
    assert h16(d["attention_mask"]) == g["mask_sha"]
    assert h16(d["position_ids"]) == g["pos_sha"] and int(d["position_ids"].sum()) == g["pos_sum"]
    assert h16(d["input_ids"]) == g["ids_sha"]
    for k in ("input_image_sizes", "denoise_image_sizes", "time_emb_inx"):
        assert _norm(d[k]) == g[k], k


@pytest.mark.parametrize("g", FB, ids=lambda g: "x".join(map(str, g["case"])))
def test_oracle_frame_block_matches_reference_golden(g):
    n_ctx, n_gen, H, W, sp = g["case"]
    _check_frame_block(po.frame_block_inputs(n_ctx, n_gen, H, W, True, sp), g)


@pytest.mark.parametrize("g", FB, ids=lambda g: "x".join(map(str, g["case"])))
def test_product_frame_block_matches_reference_golden(g):
    n_ctx, n_gen, H, W, sp = g["case"]
    proc = LVMProcessor(FakeTokenizer(), sequence_parallel_size=sp)
    imgs = [torch.zeros(3, H, W) for _ in range(n_ctx)]
    p, p_ = prompts(n_ctx, n_gen)
    d = proc.prompt_condition_frame_block_inference(
        [p, p_], [imgs, []], height=H, width=W, use_img_cfg=True,
        use_input_image_size_as_output=True, frame_blocks=[n_ctx, n_gen])
    _check_frame_block(d, g)
    # first three rows of each block of the cond row (SURVEY.md 8(c) known answers)
    m = d["attention_mask"]
    starts = [s for s, _ in d["input_image_sizes"][0]] + [r[0] - 1 for r in d["denoise_image_sizes"][0]]
    assert [[int(x) for x in m[0, s - 1:s + 2].sum(-1)] for s in starts] == g["row_sums_first3_cond"]
    geoms = d["frame_geometry"]
    assert geoms[0].n_ctx == n_ctx and geoms[1].n_ctx == 0 and geoms[1].pad == geoms[1].seq_len - geoms[1].t_gen


@pytest.mark.parametrize("g", SF, ids=lambda g: "x".join(map(str, g["case"])))
def test_single_frame_path_matches_reference_golden(g):
    n_ctx, _, H, W, sp = g["case"]
    proc = LVMProcessor(FakeTokenizer(), sequence_parallel_size=sp)
    imgs = [torch.zeros(3, H, W) for _ in range(n_ctx)]
    p = "".join(f"<img><|image_{i + 1}|></img>" for i in range(n_ctx))
    for d in (proc([p], [imgs], height=H, width=W, use_img_cfg=True, use_input_image_size_as_output=True),
              po.single_frame_inputs(n_ctx, H, W, True, sp)):
        assert d["attention_mask"].dtype == torch.uint8 and g["mask_dtype"] == "torch.uint8"
        assert h16(d["attention_mask"]) == g["mask_sha"] and int(d["attention_mask"].sum()) == g["mask_ones"]
        assert h16(d["position_ids"]) == g["pos_sha"]
        assert h16(d["input_ids"]) == g["ids_sha"]
        assert _norm(d["input_image_sizes"]) == g["input_image_sizes"]


def test_known_answers_cfg2():
    """Hashes recorded at survey time from the reference."""
    d = po.frame_block_inputs(4, 4, 256, 256, True, 1)
    assert h16(d["attention_mask"]) == "240f778342fac7ba" and int(d["attention_mask"].sum()) == 5972292
    assert h16(d["position_ids"]) == "9e1653f48781217c" and int(d["position_ids"].sum()) == 2661012


def test_closed_form_equals_oracle_loops_random_geometries():
    g = torch.Generator().manual_seed(0)
    for _ in range(25):
        n_ctx = int(torch.randint(1, 5, (1,), generator=g))   # the reference needs >= 1 context frame in row 0 (processor.py:508-509)
        n_gen = int(torch.randint(1, 5, (1,), generator=g))
        n_tok = int(torch.randint(1, 9, (1,), generator=g))
        sp = int(torch.randint(1, 9, (1,), generator=g))
        d = po.frame_block_inputs(n_ctx, n_gen, 16, 16 * n_tok, True, sp)
        L = d["input_ids"].shape[1]
        for b, (nc, ng) in enumerate([(n_ctx, n_gen), (0, n_gen)]):
            geom = FrameGeometry(L, L - (nc + ng) * (n_tok + 2), nc, ng, n_tok + 2)
            assert torch.equal(frame_block_mask(geom), d["attention_mask"][b])
            assert torch.equal(frame_block_positions(geom), d["position_ids"][b])


def test_pad_rows_and_columns():
    geom = FrameGeometry(seq_len=40, pad=4, n_ctx=1, n_gen=1, block=18)
    m = frame_block_mask(geom)
    assert bool(m[:4].all())               # pad query rows see everything (processor.py:726-727)
    assert not bool(m[4:, :4].any())       # real rows never see pad columns


def test_sequence_parallel_padding_multiple():
    c = LVMCollator(sequence_parallel_size=8)
    ids, valid, sizes = c.pad_input_ids_training([[5] * 13, [7] * 3], {0: [[1, 5]], 1: [[2, 3]]})
    assert ids.shape == (2, 16) and int(valid[0].sum()) == 13 and int(valid[1].sum()) == 3
    assert sizes == {0: [[4, 8]], 1: [[15, 16]]} and int(ids[1, 0]) == 2


def test_process_image_crops_to_multiple_of_16():
    from PIL import Image
    import numpy as np
    proc = LVMProcessor(FakeTokenizer(), max_image_size=320)
    img = Image.fromarray((np.random.RandomState(0).rand(181, 333, 3) * 255).astype(np.uint8))
    x = proc.process_image(img)
    assert x.shape[0] == 3 and x.shape[1] % 16 == 0 and x.shape[2] % 16 == 0
    assert float(x.min()) >= -1.0 and float(x.max()) <= 1.0

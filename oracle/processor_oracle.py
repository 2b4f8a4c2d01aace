"""Oracle restatement of the reference's index / mask construction (CPU, ints).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows
``LVM/processor.py`` of the reference step by step (slice assignments on a
dense matrix, Python loops) so that it is an independent check of the
product's closed-form construction in ``videogpt_b200/processor.py``.

Token ids: the released tokenizer is not available; as in SURVEY.md section 8(c)
each tag is one id after BOS strip (``processor.py:138-142``).
"""
from __future__ import annotations

import torch

IMG_OPEN, IMG_CLOSE, DIFFUSION = 32001, 32002, 32003
PAD_ID = 2  # LVMCollator(pad_token_id=2), processor.py:427


def frame_block_token_layout(n_ctx: int, n_gen: int, tokens_per_frame: int):
    """``process_multi_modal_prompt_frame_block`` (processor.py:128-179).

    Context frame: ``[<img>, 0 x N, </img>]``; generated frame:
    ``[<|diffusion|>, 0 (time slot), 0 x N]``.  Returns (ids, image ranges).
    The prompt chunk before generated frame 0 also holds the ``</img>`` of the
    last context frame; chunks are emitted in prompt order (154-177).
    """
    ids, ranges = [], []
    for i in range(n_ctx):
        # chunk i of the split prompt: "</img>" of the previous frame (if any) + "<img>"
        if i > 0:
            ids.append(IMG_CLOSE)
        ids.append(IMG_OPEN)
        s = len(ids)
        ranges.append([s, s + tokens_per_frame])
        ids.extend([0] * tokens_per_frame)
    for j in range(n_gen):
        if j == 0 and n_ctx > 0:
            ids.append(IMG_CLOSE)
        ids.append(DIFFUSION)
        ids.append(0)                      # time-embedding slot (processor.py:169)
        s = len(ids)
        ranges.append([s, s + tokens_per_frame])
        ids.extend([0] * tokens_per_frame)
    return ids, ranges


def left_pad(rows, ranges_per_row, sp_size: int = 1):
    """``pad_input_ids_training`` (processor.py:812-838): left-pad to the longest
    row rounded up to a multiple of the sequence-parallel size; shift ranges."""
    max_l = max(len(r) for r in rows)
    if max_l % sp_size != 0:
        max_l += sp_size - max_l % sp_size
    ids, valid, shifted = [], [], {}
    for b, r in enumerate(rows):
        pad = max_l - len(r)
        ids.append([PAD_ID] * pad + list(r))
        valid.append([0] * pad + [1] * len(r))
        if b in ranges_per_row:
            shifted[b] = [[s + pad, e + pad] for s, e in ranges_per_row[b]]
    return torch.LongTensor(ids), torch.ByteTensor(valid), shifted


def position_ids_frame_block(image_ranges, frame_blocks):
    """``create_position_frame_block_inference`` (processor.py:502-534)."""
    out, block_ls = [], []
    for b in image_ranges.keys():
        first = image_ranges[b][0][0]
        pad_l = first - 1 if b == 0 else first - 2      # 508-511
        token_l = image_ranges[b][-1][-1] - pad_l
        assert token_l % len(image_ranges[b]) == 0      # 514
        bl = token_l // len(image_ranges[b])
        block_ls.append(bl)
        pos = [0] * pad_l
        start = 0
        for fb in frame_blocks[b]:
            for _ in range(fb):
                pos.extend(range(start, start + bl))
                start += bl
        out.append(pos)
    return torch.LongTensor(out), block_ls


def dense_mask_frame_block(valid, block_ls, frame_blocks):
    """``create_mask_frame_block_inference`` (processor.py:682-731), row = query."""
    masks = []
    seq_len = valid.size(-1)
    for b, v in enumerate(valid):
        t = int(v.sum())
        pad = seq_len - t
        bl = block_ls[b]
        m = torch.zeros(t, t)
        r0, r1, c0, c1 = 0, bl, 0, bl
        fbs = frame_blocks[b]
        for k, fb in enumerate(fbs):
            if k != len(fbs) - 1:
                for _ in range(fb):            # context frame: 699-706
                    m[r0:, c0] = 1
                    m[r0 + 1:, c0 + 1:c1 - 1] = 1
                    m[r1 - 1:, c1 - 1] = 1
                    c0 += bl; c1 += bl; r0 += bl; r1 += bl
            else:
                for _ in range(fb):            # generated clip, first frame's rows: 708-713
                    m[r0:r1, c0] = 1
                    m[r0 + 1:r1, c0 + 1] = 1
                    m[r0 + 2:r1, c0 + 2:c1] = 1
                    c0 += bl; c1 += bl
                r0 += bl; r1 += bl
                for _ in range(fb - 1):        # copy that row pattern to the other frames: 716-720
                    m[r0:r1, c0 - fb * bl:c0] = m[r0 - bl:r1 - bl, c0 - fb * bl:c0]
                    r0 += bl; r1 += bl
        if pad > 0:                             # 722-727
            m = torch.cat([torch.zeros(t, pad), m], dim=-1)
            m = torch.cat([torch.ones(pad, seq_len), m], dim=0)
        masks.append(m.unsqueeze(0))
    return torch.cat(masks, dim=0).to(torch.bool)


def frame_block_inputs(n_ctx: int, n_gen: int, height: int, width: int,
                       use_img_cfg: bool = True, sp_size: int = 1):
    """``prompt_condition_frame_block_inference`` + ``process_mllm_input_frame_block_call``
    (processor.py:366-421, 916-1000) for already-sized frames (``height``/``width`` in
    pixels, multiples of 16).  Returns the same dict (minus pixel tensors)."""
    n_tok = height * width // 16 // 16
    rows, ranges, fbs = [], {}, {}
    ids, rg = frame_block_token_layout(n_ctx, n_gen, n_tok)
    rows.append(ids); ranges[0] = rg; fbs[0] = [n_ctx, n_gen]
    if use_img_cfg:
        ids, rg = frame_block_token_layout(0, n_gen, n_tok)   # 411-418: gen frames only
        rows.append(ids); ranges[1] = rg; fbs[1] = [0, n_gen]
    input_ids, valid, ranges = left_pad(rows, ranges, sp_size)
    position_ids, block_ls = position_ids_frame_block(ranges, fbs)
    mask = dense_mask_frame_block(valid, block_ls, fbs)
    denoise, inputs, time_inx = {}, {}, {}
    for b in ranges.keys():                     # 973-988
        n_c = fbs[b][0]
        inputs[b] = ranges[b][:n_c]
        denoise[b] = ranges[b][n_c:]
        time_inx[b] = [r[0] - 1 for r in denoise[b]]
    return {
        "input_ids": input_ids, "attention_mask": mask, "position_ids": position_ids,
        "input_image_sizes": inputs, "denoise_image_sizes": denoise,
        "time_emb_inx": time_inx, "frame_blocks": fbs,
    }


# ---------------------------------------------------------------------------
# pipeline.__call__ (one frame at a time, OmniGen-style) variant
# ---------------------------------------------------------------------------

def single_frame_inputs(n_ctx: int, height: int, width: int, use_img_cfg: bool = True,
                        sp_size: int = 1):
    """``LVMProcessor.__call__`` + ``LVMCollator.__call__`` (processor.py:282-317,
    943-962, 841-866, 783-809, 432-440, 536-573, 776-781) for ``n_ctx`` already
    sized context frames and one output frame of the same size."""
    n_tok = height * width // 16 // 16
    # process_multi_modal_prompt with the "<|diffusion|>" suffix (90-126, 276-279)
    ids, ranges = [], []
    for i in range(n_ctx):
        if i > 0:
            ids.append(IMG_CLOSE)
        ids.append(IMG_OPEN)
        s = len(ids)
        ranges.append([s, s + n_tok])
        ids.extend([0] * n_tok)
    if n_ctx > 0:
        ids.append(IMG_CLOSE)
    ids.append(DIFFUSION)
    rows = [ids]
    row_ranges = {0: ranges} if n_ctx > 0 else {}
    if use_img_cfg and n_ctx > 0:
        rows.append([DIFFUSION])               # process_multi_modal_prompt("", None): 310
    n_out = [n_tok] * len(rows)

    # pad_input_ids (783-809)
    max_l = max(len(r) + n_out[i] + 1 for i, r in enumerate(rows))
    if max_l % sp_size != 0:
        max_l += sp_size - max_l % sp_size
    padded, valid = [], []
    for i, r in enumerate(rows):
        pad = max_l - len(r) - n_out[i] - 1
        padded.append([PAD_ID] * pad + r)
        valid.append([0] * pad + [1] * len(r))
        if i in row_ranges:
            row_ranges[i] = [[s + pad, e + pad] for s, e in row_ranges[i]]
    input_ids = torch.LongTensor(padded)
    valid = torch.ByteTensor(valid)

    # create_position (432-440)
    text_len = valid.size(-1)
    img_len = max(n_out)
    pos = []
    for v in valid:
        t = int(v.sum())
        pos.append([0] * (text_len - t) + list(range(t + img_len + 1)))
    position_ids = torch.LongTensor(pos)

    # create_mask (536-573) + adjust_attention_for_input_images (776-781)
    seq_len = text_len + img_len + 1
    masks = []
    for i, v in enumerate(valid):
        t = int(v.sum())
        pad = text_len - t
        m = torch.tril(torch.ones(t + 1, t + 1))
        m = torch.cat([m, torch.zeros(t + 1, img_len)], dim=-1)
        m = torch.cat([m, torch.ones(img_len, t + img_len + 1)], dim=0)
        if pad > 0:
            m = torch.cat([torch.zeros(t + 1 + img_len, pad), m], dim=-1)
            m = torch.cat([torch.ones(pad, seq_len), m], dim=0)
        masks.append(m.unsqueeze(0))
    mask = torch.cat(masks, dim=0).to(torch.uint8)
    for b, rg in row_ranges.items():
        for s, e in rg:
            mask[b][s:e, s:e] = 1
    return {"input_ids": input_ids, "attention_mask": mask, "position_ids": position_ids,
            "input_image_sizes": row_ranges}

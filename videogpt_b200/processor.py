"""Host-side token layout, position ids, masks and index dicts for next-clip prediction.

Mirrors the interface of the reference's ``LVMProcessor`` / ``LVMCollator``
(``LVM/processor.py:23-421`` and ``426-1000``): same method names, arguments and
returned dict keys, so callers of the reference can switch without changes.  The
arithmetic is NOT the reference's Python slice-assignment loops: positions and
the dense ``[B,L,L]`` mask come from a closed form over per-token (frame, rank)
codes, evaluated with vectorised tensor ops, and the same
closed form is what the CUDA attention kernel evaluates in-register
(``csrc/attention.cu``) -- the dense mask is only materialised for API
compatibility.

Closed form, per real (non-pad) token at sequence-relative index ``i``:
``f = i // bl`` (frame), ``o = i % bl``, ``gen = f >= n_ctx`` and rank ``r``:
context frame ``<img>``->0, image->1, ``</img>``->2; generated frame
``<|diffusion|>``->0, time slot->1, image->2.  Then::

    allowed(q, k) = (not gen_k and (f_q > f_k or (f_q == f_k and r_q >= r_k)))
                 or (gen_k and gen_q and r_q >= r_k)

Pad query rows attend to everything, real rows never see pad columns
(``processor.py:722-727``).
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

__all__ = ["FrameGeometry", "LVMProcessor", "LVMCollator", "frame_block_mask",
           "frame_block_positions", "token_codes"]


@dataclass(frozen=True)
class FrameGeometry:
    """Geometry of one row of a frame-block batch (all the attention kernel needs)."""
    seq_len: int      # L, padded length of the row
    pad: int          # number of left-pad tokens
    n_ctx: int        # context frames
    n_gen: int        # frames being denoised
    block: int        # tokens per frame block (N + 2 with single-id tags)

    @property
    def tokens(self) -> int:
        return (self.n_ctx + self.n_gen) * self.block

    @property
    def t_ctx(self) -> int:
        return self.n_ctx * self.block

    @property
    def t_gen(self) -> int:
        return self.n_gen * self.block


def token_codes(geom: FrameGeometry, device=None):
    """Per-token (real, frame, rank, gen) codes of one row, as int64/bool tensors [L]."""
    idx = torch.arange(geom.seq_len, device=device) - geom.pad
    real = idx >= 0
    zero = torch.zeros_like(idx)
    f = torch.where(real, torch.div(idx, geom.block, rounding_mode="floor"), zero)
    o = torch.where(real, idx % geom.block, zero)
    gen = f >= geom.n_ctx
    r_ctx = torch.where(o == 0, 0, torch.where(o == geom.block - 1, 2, 1))
    r = torch.where(gen, o.clamp(max=2), r_ctx)
    return real, f, r, gen


def frame_block_mask(geom: FrameGeometry, device=None) -> torch.Tensor:
    """Dense bool ``[L, L]`` mask (row = query) of one row; bit-exact with
    ``create_mask_frame_block_inference`` (``processor.py:682-731``)."""
    real, f, r, gen = token_codes(geom, device)
    fq, fk = f[:, None], f[None, :]
    rq, rk = r[:, None], r[None, :]
    gq, gk = gen[:, None], gen[None, :]
    m = ((~gk) & ((fq > fk) | ((fq == fk) & (rq >= rk)))) | (gk & gq & (rq >= rk))
    return (m & real[None, :] & real[:, None]) | (~real)[:, None]


def frame_block_positions(geom: FrameGeometry, device=None) -> torch.Tensor:
    """``position_ids`` of one row: sequential from the first real token
    (``create_position_frame_block_inference``, ``processor.py:502-534``)."""
    return (torch.arange(geom.seq_len, device=device) - geom.pad).clamp(min=0)


class LVMCollator:
    """Same constructor and inference entry points as the reference's ``LVMCollator``
    (``processor.py:426-430``)."""

    def __init__(self, pad_token_id=2, hidden_size=3072, sequence_parallel_size=1):
        self.pad_token_id = pad_token_id
        self.hidden_size = hidden_size
        self.sequence_parallel_size = sequence_parallel_size

    # -- shared ------------------------------------------------------------------
    def _round_up(self, n: int) -> int:
        p = self.sequence_parallel_size
        return n if n % p == 0 else n + p - n % p

    # -- frame-block (next-clip) path ---------------------------------------------
    def pad_input_ids_training(self, input_ids, image_sizes):
        """Left-pad to the longest row rounded up to the SP size; shift the image
        ranges (``processor.py:812-838``)."""
        max_l = self._round_up(max(len(x) for x in input_ids))
        out = torch.full((len(input_ids), max_l), self.pad_token_id, dtype=torch.long)
        valid = torch.zeros((len(input_ids), max_l), dtype=torch.uint8)
        for i, ids in enumerate(input_ids):
            pad = max_l - len(ids)
            out[i, pad:] = torch.as_tensor(ids, dtype=torch.long)
            valid[i, pad:] = 1
            if i in image_sizes:
                image_sizes[i] = [[s + pad, e + pad] for s, e in image_sizes[i]]
        return out, valid, image_sizes

    def frame_geometry(self, seq_len: int, image_sizes, frame_blocks) -> Dict[int, FrameGeometry]:
        """Row geometry from the padded image ranges, derived exactly as the reference
        derives ``pad_l`` / ``block_l`` (``processor.py:508-516``)."""
        geoms = {}
        for b in image_sizes.keys():
            first = image_sizes[b][0][0]
            pad = first - 1 if b == 0 else first - 2
            token_l = image_sizes[b][-1][-1] - pad
            n_frames = len(image_sizes[b])
            assert token_l % n_frames == 0, "frame blocks must have equal length"
            fb = frame_blocks[b]
            assert len(fb) == 2 and fb[0] + fb[1] == n_frames, \
                "frame_blocks must be [n_context, n_generated]"
            geoms[b] = FrameGeometry(seq_len=seq_len, pad=pad, n_ctx=fb[0], n_gen=fb[1],
                                     block=token_l // n_frames)
        return geoms

    def create_position_frame_block_inference(self, image_sizes, frame_blocks, seq_len=None):
        if seq_len is None:
            seq_len = max(v[-1][-1] for v in image_sizes.values())
        geoms = self.frame_geometry(seq_len, image_sizes, frame_blocks)
        pos = torch.stack([frame_block_positions(geoms[b]) for b in image_sizes.keys()])
        return pos, [geoms[b].block for b in image_sizes.keys()]

    def create_mask_frame_block_inference(self, attention_mask, block_ls, frame_blocks):
        seq_len = attention_mask.size(-1)
        rows = []
        for b, valid in enumerate(attention_mask):
            t = int(valid.sum())
            fb = frame_blocks[b]
            rows.append(frame_block_mask(FrameGeometry(seq_len, seq_len - t, fb[0], fb[1], block_ls[b])))
        return torch.stack(rows)

    def process_mllm_input_frame_block_call(self, features, build_dense_mask: bool = True):
        """Collate frame-block rows (``processor.py:916-1000``).  Returns the
        reference's dict plus ``frame_geometry`` (what the CUDA path consumes)."""
        pixel_values, image_sizes, frame_blocks = [], {}, {}
        for b, x in enumerate(features):
            if x["pixel_values"] is not None:
                pixel_values.extend(x["pixel_values"])
                for size in x["image_sizes"]:
                    image_sizes.setdefault(b, []).append(size)
                    frame_blocks[b] = x["frame_blocks"]
        pixel_values = [x.unsqueeze(0) for x in pixel_values]
        input_ids, valid, image_sizes = self.pad_input_ids_training(
            [x["input_ids"] for x in features], image_sizes)
        seq_len = input_ids.shape[1]
        geoms = self.frame_geometry(seq_len, image_sizes, frame_blocks)
        position_ids = torch.stack([frame_block_positions(geoms[b]) for b in image_sizes.keys()])
        attention_mask = None
        if build_dense_mask:
            attention_mask = torch.stack([frame_block_mask(geoms[b]) for b in image_sizes.keys()])

        denoise, inputs, time_inx, input_images = {}, {}, {}, []
        idx_pix = 0
        for b in image_sizes.keys():
            n_c = frame_blocks[b][0]
            inputs[b] = image_sizes[b][:n_c]
            denoise[b] = image_sizes[b][n_c:]
            time_inx[b] = [r[0] - 1 for r in denoise[b]]
        for b in image_sizes.keys():
            n_c = frame_blocks[b][0]
            input_images.extend(pixel_values[idx_pix:idx_pix + n_c])
            idx_pix += n_c
        return {
            "input_ids": input_ids,
            "attention_mask": attention_mask,
            "position_ids": position_ids,
            "input_pixel_values": input_images,
            "input_image_sizes": inputs,
            "denoise_image_sizes": denoise,
            "output_images": [],
            "time_emb_inx": time_inx,
            "frame_blocks": frame_blocks,
            "frame_geometry": geoms,
        }

    # -- pipeline.__call__ (one frame at a time) path -------------------------------
    def pad_input_ids(self, input_ids, image_sizes, num_tokens_for_output_images):
        """``processor.py:783-809``."""
        max_l = self._round_up(max(len(x) + num_tokens_for_output_images[i] + 1
                                   for i, x in enumerate(input_ids)))
        rows, valid = [], []
        for i, ids in enumerate(input_ids):
            pad = max_l - len(ids) - num_tokens_for_output_images[i] - 1
            rows.append([self.pad_token_id] * pad + list(ids))
            valid.append([0] * pad + [1] * len(ids))
            if i in image_sizes:
                image_sizes[i] = [[s + pad, e + pad] for s, e in image_sizes[i]]
        return torch.LongTensor(rows), torch.ByteTensor(valid), image_sizes

    def create_position(self, attention_mask, num_tokens_for_output_images):
        """``processor.py:432-440``: positions restart after the pad, and run on over the
        time token and the output image tokens."""
        text_len = attention_mask.size(-1)
        total = text_len + max(num_tokens_for_output_images) + 1
        pads = text_len - attention_mask.long().sum(-1)
        return (torch.arange(total)[None, :] - pads[:, None]).clamp(min=0)

    def create_mask(self, attention_mask, num_tokens_for_output_images):
        """``processor.py:536-573``: text causal, then ``[time, image]`` rows see all real
        tokens; shorter output images are right-padded and masked."""
        text_len = attention_mask.size(-1)
        img_len = max(num_tokens_for_output_images)
        seq_len = text_len + img_len + 1
        idx = torch.arange(seq_len)
        out, padding_images = [], []
        for b, valid in enumerate(attention_mask):
            t = int(valid.sum())
            pad = text_len - t
            real = idx >= pad
            q, k = idx[:, None], idx[None, :]
            image_row = q > text_len           # rows after the time token
            m = real[None, :] & real[:, None] & ((k <= q) | image_row)
            m = m | (~real)[:, None]
            pad_img = img_len - num_tokens_for_output_images[b]
            if pad_img > 0:
                m[:, seq_len - pad_img:] = False
                padding_images.append(torch.zeros(1, pad_img, self.hidden_size))
            else:
                padding_images.append(None)
            out.append(m)
        return torch.stack(out).to(torch.uint8), padding_images

    def adjust_attention_for_input_images(self, attention_mask, image_sizes):
        for b in image_sizes.keys():
            for s, e in image_sizes[b]:
                attention_mask[b][s:e, s:e] = 1
        return attention_mask

    def process_mllm_input(self, mllm_inputs, target_img_size):
        n_out = [s[0] * s[1] // 16 // 16 for s in target_img_size]
        pixel_values, image_sizes = [], {}
        for b, x in enumerate(mllm_inputs):
            if x["pixel_values"] is not None:
                pixel_values.extend(x["pixel_values"])
                for size in x["image_sizes"]:
                    image_sizes.setdefault(b, []).append(size)
        pixel_values = [x.unsqueeze(0) for x in pixel_values]
        input_ids, valid, image_sizes = self.pad_input_ids(
            [x["input_ids"] for x in mllm_inputs], image_sizes, n_out)
        position_ids = self.create_position(valid, n_out)
        mask, padding_images = self.create_mask(valid, n_out)
        mask = self.adjust_attention_for_input_images(mask, image_sizes)
        return input_ids, position_ids, mask, padding_images, pixel_values, image_sizes

    def __call__(self, features):
        mllm_inputs = [f[0] for f in features]
        cfg_inputs = [f[1] for f in features]
        target = [f[2] for f in features]
        if cfg_inputs[0] is not None:
            mllm_inputs = mllm_inputs + cfg_inputs
            target = target + target
        ids, pos, mask, padding, pixels, sizes = self.process_mllm_input(mllm_inputs, target)
        return {"input_ids": ids, "attention_mask": mask, "position_ids": pos,
                "input_pixel_values": pixels, "input_image_sizes": sizes,
                "padding_images": padding}


class LVMProcessor:
    """Same constructor / methods as the reference's ``LVMProcessor``
    (``processor.py:23-421``)."""

    _IMAGE_TAG = re.compile(r"<\|image_\d+\|>")

    def __init__(self, text_tokenizer, max_image_size: int = 1024, sequence_parallel_size: int = 1):
        self.text_tokenizer = text_tokenizer
        self.max_image_size = max_image_size
        self.sequence_parallel_size = sequence_parallel_size
        self.collator = LVMCollator(sequence_parallel_size=sequence_parallel_size)

    @classmethod
    def from_pretrained(cls, model_name, sequence_parallel_size=1):
        from transformers import AutoTokenizer
        return cls(AutoTokenizer.from_pretrained(model_name),
                   sequence_parallel_size=sequence_parallel_size)

    # -- images -------------------------------------------------------------------
    def crop_arr(self, pil_image):
        """Resize to at most ``max_image_size`` and centre-crop both sides to a
        multiple of 16 (``processor.py:41-67``)."""
        from PIL import Image
        while min(*pil_image.size) >= 2 * self.max_image_size:
            pil_image = pil_image.resize(tuple(x // 2 for x in pil_image.size), resample=Image.BOX)
        if max(*pil_image.size) > self.max_image_size:
            scale = self.max_image_size / max(*pil_image.size)
            pil_image = pil_image.resize(tuple(round(x * scale) for x in pil_image.size),
                                         resample=Image.BICUBIC)
        if min(*pil_image.size) < 16:
            scale = 16 / min(*pil_image.size)
            pil_image = pil_image.resize(tuple(round(x * scale) for x in pil_image.size),
                                         resample=Image.BICUBIC)
        arr = np.array(pil_image)
        y1 = (arr.shape[0] % 16) // 2
        y2 = arr.shape[0] % 16 - y1
        x1 = (arr.shape[1] % 16) // 2
        x2 = arr.shape[1] % 16 - x1
        return Image.fromarray(arr[y1:arr.shape[0] - y2, x1:arr.shape[1] - x2])

    def process_image(self, image):
        """PIL image (or path) -> float tensor ``[3,H,W]`` in [-1, 1] (``processor.py:80-88``)."""
        from PIL import Image
        if isinstance(image, str):
            image = Image.open(image).convert("RGB")
        elif not isinstance(image, Image.Image):
            raise ValueError("Input must be a PIL.Image object")
        arr = np.asarray(self.crop_arr(image), dtype=np.uint8)
        if arr.ndim == 2:
            arr = arr[:, :, None]
        x = torch.from_numpy(arr.copy()).permute(2, 0, 1).to(torch.float32).div_(255.0)
        return x.sub_(0.5).div_(0.5)

    # -- prompts ------------------------------------------------------------------
    def _chunks(self, text):
        chunks = [list(self.text_tokenizer(c).input_ids) for c in self._IMAGE_TAG.split(text)]
        chunks = [c[1:] if len(c) > 0 and c[0] == 1 else c for c in chunks]
        tags = self._IMAGE_TAG.findall(text)
        image_ids = [int(s.split("|")[1].split("_")[-1]) for s in tags]
        unique = sorted(set(image_ids))
        assert unique == list(range(1, len(unique) + 1)), \
            f"image_ids must start from 1, and must be continuous int, e.g. [1, 2, 3], cannot be {unique}"
        return chunks, image_ids, unique

    def add_prefix_instruction(self, prompt):
        return f"{prompt}<|diffusion|>"

    def _text_only(self, text):
        ids = list(self.text_tokenizer(text).input_ids)
        if ids and ids[0] == 1:
            ids = ids[1:]
        return {"input_ids": ids, "pixel_values": None, "image_sizes": None}

    def process_multi_modal_prompt(self, text, input_images):
        """``processor.py:90-126``."""
        text = self.add_prefix_instruction(text)
        if input_images is None or len(input_images) == 0:
            return self._text_only(text)
        chunks, image_ids, unique = self._chunks(text)
        assert len(unique) == len(input_images), \
            f"total images must be the same as the number of image tags, got {len(unique)} image tags and {len(input_images)} images"
        input_images = [input_images[x - 1] for x in image_ids]
        ids, ranges = [], []
        for i, chunk in enumerate(chunks):
            ids.extend(chunk)
            if i != len(chunks) - 1:
                size = input_images[i].size(-2) * input_images[i].size(-1) // 16 // 16
                ranges.append([len(ids), len(ids) + size])
                ids.extend([0] * size)
        return {"input_ids": ids, "pixel_values": input_images, "image_sizes": ranges}

    def process_multi_modal_prompt_frame_block(self, text, input_images, frame_blocks,
                                               height=None, width=None):
        """``processor.py:128-179``: context frames ``[tags, 0 x N]``, generated frames
        ``[tags, 0 (time slot), 0 x N]``."""
        if input_images is None:
            input_images = []
        if len(input_images) == 0 and (height is None and width is None):
            return self._text_only(text)
        chunks, image_ids, unique = self._chunks(text)
        assert len(unique) == len(input_images) + frame_blocks[-1], \
            f"total images must be the same as the number of image tags, got {len(unique)} image tags and {len(input_images)} images"
        input_images = [input_images[x - 1] for x in image_ids[:frame_blocks[0]]]
        ids, ranges, idx = [], [], 0
        for k, fb in enumerate(frame_blocks):
            last = k == len(frame_blocks) - 1
            for _ in range(fb):
                ids.extend(chunks[idx])
                if last:
                    ids.append(0)
                    if height is not None and width is not None:
                        size = height * width // 16 // 16
                    else:
                        size = input_images[0].size(-2) * input_images[0].size(-1) // 16 // 16
                else:
                    size = input_images[idx].size(-2) * input_images[idx].size(-1) // 16 // 16
                ranges.append([len(ids), len(ids) + size])
                ids.extend([0] * size)
                idx += 1
        return {"input_ids": ids, "pixel_values": input_images, "image_sizes": ranges}

    def _images(self, imgs):
        if imgs is not None and len(imgs) > 0:
            return [x if torch.is_tensor(x) else self.process_image(x) for x in imgs]
        return None

    def __call__(self, instructions: List[str], input_images=None, height: int = 1024,
                 width: int = 1024, use_img_cfg: bool = True,
                 use_input_image_size_as_output: bool = False) -> Dict:
        """``processor.py:282-317``."""
        if input_images is None:
            use_img_cfg = False
        if isinstance(instructions, str):
            instructions = [instructions]
            input_images = [input_images]
        data = []
        for i, text in enumerate(instructions):
            imgs = self._images(None if input_images is None else input_images[i])
            if imgs is None:
                assert "<img><|image_1|></img>" not in text
            mllm = self.process_multi_modal_prompt(text, imgs)
            cfg = self.process_multi_modal_prompt("", None) if use_img_cfg else None
            if use_input_image_size_as_output:
                size = [mllm["pixel_values"][0].size(-2), mllm["pixel_values"][0].size(-1)]
            else:
                size = [height, width]
            data.append((mllm, cfg, size))
        return self.collator(data)

    def prompt_condition_frame_block_inference(self, instructions: List[str], input_images=None,
                                               height: int = 1024, width: int = 1024,
                                               use_img_cfg: bool = True,
                                               use_input_image_size_as_output: bool = False,
                                               frame_blocks: Optional[List[int]] = None,
                                               build_dense_mask: bool = True) -> Dict:
        """``processor.py:366-421``.  ``input_images`` entries may be PIL images / paths
        (processed like the reference) or already-processed ``[3,H,W]`` tensors."""
        if input_images is None:
            use_img_cfg = False
        if isinstance(instructions, str):
            instructions = [instructions]
            input_images = [input_images]
        imgs = self._images(None if input_images is None else input_images[0])
        if imgs is None:
            assert "<img><|image_1|></img>" not in instructions[0]
        mllm = self.process_multi_modal_prompt_frame_block(instructions[0], imgs, frame_blocks)
        mllm["frame_blocks"] = frame_blocks
        data = [mllm]
        if use_img_cfg:
            imgs1 = self._images(None if input_images is None else input_images[1])
            if imgs1 is None:
                assert "<img><|image_1|></img>" not in instructions[1]
            cfg = self.process_multi_modal_prompt_frame_block(
                instructions[1], imgs1, [0, frame_blocks[-1]],
                height=mllm["pixel_values"][0].size(-2), width=mllm["pixel_values"][0].size(-1))
            cfg["frame_blocks"] = [0, frame_blocks[-1]]
            data.append(cfg)
        return self.collator.process_mllm_input_frame_block_call(data, build_dense_mask=build_dense_mask)

"""``LVMPipeline``: drop-in for the reference's user API (``LVM/pipeline.py:46-595``).

``prompt_condition_frame_block_autoregressive_inference`` keeps the reference's signature and
host flow (prompt strings, seeded noise, VAE encode of the context, scheduler, VAE decode to
PIL).  The VAE stays whatever ``diffusers``-style ``AutoencoderKL`` the caller passes in (it
is outside the hot path and excluded from the headline number, BASELINE.json); everything
between the latents is the B200 engine.

``next_clip_latents`` is the same computation in latent space -- context latents in,
generated latents out -- for callers (and ``bench.py``) that hold latents already; it is what
the autoregressive method calls between its two VAE passes.
"""
from __future__ import annotations

import gc
from typing import List, Optional, Union

import torch

from .model import LVM
from .parallel_states import hccl_info
from .processor import LVMProcessor
from .scheduler import LVMScheduler


def frame_block_prompts(n_ctx: int, gen_num: int):
    """Prompt strings of one round (pipeline.py:426-448)."""
    prompt = "".join(f"<img><|image_{i + 1}|></img>" for i in range(n_ctx))
    prompt += "".join(f"<|diffusion|><|image_{n_ctx + i + 1}|>" for i in range(gen_num))
    prompt_ = "".join(f"<|diffusion|><|image_{i + 1}|>" for i in range(gen_num))
    return prompt, prompt_


def replicate_frame_block_inputs(data: dict, n_videos: int) -> dict:
    """Index dicts of ONE video (rows: conditional [, unconditional]) -> those of ``n_videos`` videos
    of the same geometry in one batch: rows ``0..n-1`` are the conditional rows, rows ``n..2n-1``
    the unconditional ones, so that latents and predictions keep the ``[cond..., uncond...]`` order
    the sampler's CFG expects (``LVM/scheduler.py:186-199``) and the model's running counters
    over rows (``LVM/model.py:436-453``) number context and noisy latents video by video."""
    rows = data["input_ids"].shape[0]
    order = [r for r in range(rows) for _ in range(n_videos)]           # source row of every batch row
    out = dict(data)
    for k in ("input_ids", "position_ids", "attention_mask"):
        if data.get(k) is not None:
            out[k] = data[k][order].contiguous()
    for k in ("input_image_sizes", "denoise_image_sizes", "time_emb_inx"):
        out[k] = {b: list(data[k][src]) for b, src in enumerate(order) if src in data[k]}
    if "frame_blocks" in data:
        out["frame_blocks"] = {b: data["frame_blocks"][src] for b, src in enumerate(order)}
    return out


class LVMPipeline:
    def __init__(self, vae, model: LVM, processor: LVMProcessor, device: Union[str, torch.device] = None):
        self.vae = vae
        self.model = model
        self.processor = processor
        self.device = device
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("videogpt_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            self.device = torch.device("cuda")
        self.model.eval()
        if self.vae is not None:
            self.vae.eval()
        self.model_cpu_offload = False

    @classmethod
    def from_pretrained(cls, model_name, vae_path: str = None, load_llm_ckpt=True):
        """pipeline.py:74-95 (needs ``diffusers`` for the VAE, as the reference does)."""
        import os
        model = LVM.from_pretrained(model_name, load_llm_ckpt=load_llm_ckpt)
        processor = LVMProcessor.from_pretrained(model_name, sequence_parallel_size=max(hccl_info.world_size, 1))
        from diffusers.models import AutoencoderKL
        if os.path.exists(os.path.join(model_name, "vae")):
            vae = AutoencoderKL.from_pretrained(os.path.join(model_name, "vae"))
        else:
            vae = AutoencoderKL.from_pretrained(vae_path if vae_path is not None else "stabilityai/sdxl-vae")
        return cls(vae, model, processor)

    def to(self, device: Union[str, torch.device]):
        if isinstance(device, str):
            device = torch.device(device)
        self.model.to(device)
        if self.vae is not None:
            self.vae.to(device)
        self.device = device

    def vae_encode(self, x, dtype):
        """pipeline.py:110-117."""
        if self.vae.config.shift_factor is not None:
            x = self.vae.encode(x).latent_dist.sample()
            x = (x - self.vae.config.shift_factor) * self.vae.config.scaling_factor
        else:
            x = self.vae.encode(x).latent_dist.sample().mul_(self.vae.config.scaling_factor)
        return x.to(dtype)

    def vae_decode_to_pil(self, latent):
        """pipeline.py:559-570 / 572-590."""
        from PIL import Image
        latent = latent.to(torch.float32)
        if self.vae.config.shift_factor is not None:
            latent = latent / self.vae.config.scaling_factor + self.vae.config.shift_factor
        else:
            latent = latent / self.vae.config.scaling_factor
        img = self.vae.decode(latent).sample
        img = (img * 0.5 + 0.5).clamp(0, 1)
        img = (img * 255).to("cpu", dtype=torch.uint8).permute(0, 2, 3, 1).numpy()
        return Image.fromarray(img[0])

    def move_to_device(self, data):
        if isinstance(data, list):
            return [x.to(self.device) for x in data]
        return data.to(self.device)

    # ---- latent-space next-clip prediction (the hot path) ----------------------------------------
    @torch.no_grad()
    def next_clip_latents(self, context_latents: List[torch.Tensor], gen_num: int,
                          num_inference_steps: int = 50, img_guidance_scale: float = 1.6,
                          use_img_guidance: bool = True, seed: Optional[int] = None,
                          time_shifting_factor: float = 1.0, prediction_type: str = "v",
                          dtype: torch.dtype = torch.bfloat16, initial_noise: Optional[List[torch.Tensor]] = None,
                          scheduler: Optional[LVMScheduler] = None) -> List[torch.Tensor]:
        """Context latents ``[1,4,h,w]`` (host or device) -> ``gen_num`` generated latents.

        Follows one round of the reference (pipeline.py:418-549): token layout and index dicts
        from the processor, ``randn`` noise per generated frame from ``torch.Generator(device)``
        seeded with ``seed`` (473-481) unless ``initial_noise`` is given, the cond list duplicated
        for the unconditional branch (482), ``LVMScheduler`` over ``frame_block_forward_with_cfg``,
        first half of the samples returned (549)."""
        if img_guidance_scale == 1:
            use_img_guidance = False
        n_ctx = len(context_latents)
        lat_h, lat_w = context_latents[0].shape[-2:]
        height, width = lat_h * 8, lat_w * 8
        prompt, prompt_ = frame_block_prompts(n_ctx, gen_num)
        placeholders = [torch.empty(3, height, width, device="meta") for _ in range(n_ctx)]
        instructions = [prompt, prompt_] if use_img_guidance else [prompt]
        images = [placeholders, []] if use_img_guidance else [placeholders]
        self.model.to(dtype)
        data = self.processor.prompt_condition_frame_block_inference(
            instructions, images, height=height, width=width, use_img_cfg=use_img_guidance,
            use_input_image_size_as_output=True, frame_blocks=[n_ctx, gen_num], build_dense_mask=False)
        num_cfg = 1 if use_img_guidance else 0
        if initial_noise is not None:
            latents = [x.to(self.device, dtype) for x in initial_noise]
        else:
            generator = torch.Generator(device=self.device).manual_seed(seed) if seed is not None else None
            latents = [torch.randn(1, 4, lat_h, lat_w, device=self.device, generator=generator).to(dtype)
                       for _ in range(gen_num)]
        latents = latents * (1 + num_cfg)
        ctx = [x.to(self.device, dtype, non_blocking=True) for x in context_latents]
        model_kwargs = dict(
            input_ids=data["input_ids"], input_img_latents=ctx, input_image_sizes=data["input_image_sizes"],
            attention_mask=None, position_ids=data["position_ids"],
            denoise_image_sizes=data["denoise_image_sizes"], time_emb_inx=data["time_emb_inx"],
            img_cfg_scale=img_guidance_scale, use_img_cfg=use_img_guidance, use_kv_cache=False,
            offload_model=False, vae=self.vae)
        scheduler = scheduler or LVMScheduler(num_steps=num_inference_steps, time_shifting_factor=time_shifting_factor)
        samples = scheduler(latents, self.model.frame_block_forward_with_cfg, model_kwargs,
                            use_kv_cache=False, offload_kv_cache=False, prediction_type=prediction_type,
                            vae=self.vae)
        return samples[:len(samples) // 2] if use_img_guidance else samples

    @torch.no_grad()
    def next_clip_latents_batch(self, context_latents: List[List[torch.Tensor]], gen_num: int,
                                num_inference_steps: int = 50, img_guidance_scale: float = 1.6,
                                use_img_guidance: bool = True, seed: Optional[int] = None,
                                time_shifting_factor: float = 1.0, prediction_type: str = "v",
                                dtype: torch.dtype = torch.bfloat16,
                                initial_noise: Optional[List[List[torch.Tensor]]] = None,
                                scheduler: Optional[LVMScheduler] = None) -> List[List[torch.Tensor]]:
        """``next_clip_latents`` for several independent videos of the same geometry in ONE pass
        (BASELINE.json configs[3]): ``context_latents[v]`` are the context latents of video ``v``;
        returns ``gen_num`` generated latents per video.

        The reference's pipeline handles one video per call; its model, however, takes any number
        of rows (``LVM/model.py:436-453``), and that is what this uses: all rows of all videos
        and both CFG branches are packed into one ``[M, hidden]`` matrix, so every projection is
        one GEMM over ``n_videos`` x more rows.  Every row-wise kernel and the attention kernel
        give a row the same bits whatever else is in the batch, so video ``v`` of a batch equals
        the same video run alone with the same noise."""
        n_videos = len(context_latents)
        if n_videos == 0:
            return []
        if img_guidance_scale == 1:
            use_img_guidance = False
        n_ctx = len(context_latents[0])
        lat_h, lat_w = context_latents[0][0].shape[-2:]
        for ctx in context_latents:
            if len(ctx) != n_ctx or any(tuple(x.shape[-2:]) != (lat_h, lat_w) for x in ctx):
                raise ValueError("all videos of a batch must have the same number and size of context frames")
        height, width = lat_h * 8, lat_w * 8
        prompt, prompt_ = frame_block_prompts(n_ctx, gen_num)
        placeholders = [torch.empty(3, height, width, device="meta") for _ in range(n_ctx)]
        instructions = [prompt, prompt_] if use_img_guidance else [prompt]
        images = [placeholders, []] if use_img_guidance else [placeholders]
        self.model.to(dtype)
        data = replicate_frame_block_inputs(self.processor.prompt_condition_frame_block_inference(
            instructions, images, height=height, width=width, use_img_cfg=use_img_guidance,
            use_input_image_size_as_output=True, frame_blocks=[n_ctx, gen_num], build_dense_mask=False), n_videos)
        if initial_noise is not None:
            if len(initial_noise) != n_videos or any(len(n) != gen_num for n in initial_noise):
                raise ValueError("initial_noise must hold gen_num latents for every video")
            latents = [x.to(self.device, dtype) for video in initial_noise for x in video]
        else:
            generator = torch.Generator(device=self.device).manual_seed(seed) if seed is not None else None
            latents = [torch.randn(1, 4, lat_h, lat_w, device=self.device, generator=generator).to(dtype)
                       for _ in range(n_videos * gen_num)]
        latents = latents * (2 if use_img_guidance else 1)
        ctx = [x.to(self.device, dtype, non_blocking=True) for video in context_latents for x in video]
        model_kwargs = dict(
            input_ids=data["input_ids"], input_img_latents=ctx, input_image_sizes=data["input_image_sizes"],
            attention_mask=None, position_ids=data["position_ids"],
            denoise_image_sizes=data["denoise_image_sizes"], time_emb_inx=data["time_emb_inx"],
            img_cfg_scale=img_guidance_scale, use_img_cfg=use_img_guidance, use_kv_cache=False,
            offload_model=False, vae=self.vae)
        scheduler = scheduler or LVMScheduler(num_steps=num_inference_steps, time_shifting_factor=time_shifting_factor)
        samples = scheduler(latents, self.model.frame_block_forward_with_cfg, model_kwargs,
                            use_kv_cache=False, offload_kv_cache=False, prediction_type=prediction_type,
                            vae=self.vae)
        return [samples[v * gen_num:(v + 1) * gen_num] for v in range(n_videos)]

    @torch.no_grad()
    def next_frame_latents(self, context_latents: Optional[List[torch.Tensor]], height: int = None,
                           width: int = None, num_inference_steps: int = 50, img_guidance_scale: float = 1.6,
                           use_img_guidance: bool = True, seed: Optional[int] = None,
                           time_shifting_factor: float = 1.0, prediction_type: str = "v",
                           dtype: torch.dtype = torch.bfloat16, initial_noise: Optional[torch.Tensor] = None,
                           use_kv_cache: bool = True) -> torch.Tensor:
        """One iteration of ``LVMPipeline.__call__`` (pipeline.py:215-296) in latent space: context
        latents ``[1,4,h,w]`` (or none: unconditional generation of a ``height`` x ``width`` frame)
        -> ONE generated latent ``[1,4,h,w]``, through the one-frame-at-a-time layout
        (``LVMProcessor.__call__``, ``LVM.forward_with_cfg``)."""
        ctx_in = list(context_latents or [])
        if img_guidance_scale == 1 or not ctx_in:
            use_img_guidance = False                                                     # 207-208, 217-218
        if ctx_in and (height is None or width is None):
            height, width = ctx_in[0].shape[-2] * 8, ctx_in[0].shape[-1] * 8
        assert height is not None and width is not None and height % 16 == 0 and width % 16 == 0, \
            "The height and width must be a multiple of 16."
        lat_h, lat_w = height // 8, width // 8
        prompt = "".join(f"<img><|image_{i + 1}|></img>" for i in range(len(ctx_in)))    # 222-225
        images = [[torch.empty(3, x.shape[-2] * 8, x.shape[-1] * 8, device="meta") for x in ctx_in]] if ctx_in else None
        self.model.to(dtype)
        data = self.processor([prompt], images, height=height, width=width, use_img_cfg=use_img_guidance,
                              use_input_image_size_as_output=False)
        num_cfg = 1 if use_img_guidance else 0
        if initial_noise is not None:
            latents = initial_noise.to(self.device).reshape(1, 4, lat_h, lat_w)
        else:
            generator = torch.Generator(device=self.device).manual_seed(seed) if seed is not None else None
            latents = torch.randn(1, 4, lat_h, lat_w, device=self.device, generator=generator)   # 250
        latents = torch.cat([latents] * (1 + num_cfg), 0).to(dtype)                               # 251
        model_kwargs = dict(
            input_ids=self.move_to_device(data["input_ids"]),
            input_img_latents=[x.to(self.device, dtype, non_blocking=True) for x in ctx_in],
            input_image_sizes=data["input_image_sizes"],
            attention_mask=self.move_to_device(data["attention_mask"]),
            position_ids=self.move_to_device(data["position_ids"]), img_cfg_scale=img_guidance_scale,
            use_img_cfg=use_img_guidance, use_kv_cache=use_kv_cache, offload_model=False)
        scheduler = LVMScheduler(num_steps=num_inference_steps, time_shifting_factor=time_shifting_factor)
        samples = scheduler(latents, self.model.forward_with_cfg, model_kwargs, use_kv_cache=use_kv_cache,
                            offload_kv_cache=False, prediction_type=prediction_type, vae=self.vae)
        return samples.chunk(1 + num_cfg, dim=0)[0]                                               # 297

    @torch.no_grad()
    def rollout_latents(self, context_latents: List[torch.Tensor], gen_nums: List[int],
                        num_inference_steps: int = 50, img_guidance_scale: float = 1.6,
                        use_img_guidance: bool = True, seed: Optional[int] = None,
                        time_shifting_factor: float = 1.0, prediction_type: str = "v",
                        dtype: torch.dtype = torch.bfloat16, clean_image_noise_level: float = 0.0,
                        max_frame_window: int = 16, persistent_cache: bool = True) -> List[torch.Tensor]:
        """The autoregressive loop of ``prompt_condition_frame_block_autoregressive_inference``
        (pipeline.py:418-424, 485-500) in LATENT space: no VAE decode / uint8 / re-encode between
        rounds.  Returns the latents of all generated frames.

        ``persistent_cache`` (needs equal ``gen_nums``): every context frame is prefilled once and
        its K/V stay in the paged pool across rounds (``rollout.LatentRollout`` -- identical to the
        reference's round while the window holds the whole history, windowed attention over the
        cached K/V afterwards).  Otherwise every round recomputes its window from scratch exactly
        like the reference (positions restart at the window start)."""
        self.model.to(dtype)
        generated: List[torch.Tensor] = []
        if persistent_cache and len(set(gen_nums)) == 1:
            from .rollout import LatentRollout
            ro = LatentRollout(self.model, self.processor, gen_nums[0], max_frame_window, num_inference_steps,
                               img_guidance_scale, use_img_guidance, time_shifting_factor, prediction_type,
                               clean_image_noise_level, rounds_hint=len(gen_nums)).start(context_latents)
            for _ in gen_nums:
                generated += ro.next_clip(seed=seed)
            return generated
        frames = [x.to(self.device, dtype) for x in context_latents]
        for k, gen_num in enumerate(gen_nums):
            if len(frames) + gen_num > max_frame_window:                                  # 421-422
                frames = frames[gen_num + len(frames) - max_frame_window:]
            out = self.next_clip_latents(frames, gen_num, num_inference_steps=num_inference_steps,
                                         img_guidance_scale=img_guidance_scale, use_img_guidance=use_img_guidance,
                                         seed=seed, time_shifting_factor=time_shifting_factor,
                                         prediction_type=prediction_type, dtype=dtype)
            generated += out
            a = clean_image_noise_level or 0.0
            frames = frames + [((1 - a) * x + a * torch.randn_like(x)) if a else x for x in out]
        return generated

    # ---- reference user API --------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, input_images=None, height: int = 1024, width: int = 1024, gen_num: int = 1,
                 num_inference_steps: int = 50, use_img_guidance: bool = True, img_guidance_scale: float = 1.6,
                 max_input_image_size: int = 1024, offload_model: bool = False, use_kv_cache: bool = True,
                 offload_kv_cache: bool = True, use_input_image_size_as_output: bool = False,
                 dtype: torch.dtype = torch.bfloat16, seed: int = None, output_type: str = "pil",
                 time_shifting_factor: float = 1.0, prediction_type: str = "v",
                 clean_image_noise_level: float = None):
        """``LVMPipeline.__call__`` (pipeline.py:136-343): ``gen_num`` frames, one at a time, each
        conditioned on the input images plus the frames generated so far (re-encoded through the
        VAE every iteration and noised by ``clean_image_noise_level``, 256-260).  Returns the PIL
        list: reconstructions of the input images, then the generated frames.

        The arithmetic between the two VAE passes is ``next_frame_latents``.  As in the reference
        the output frame takes the size of the (first) input image when
        ``use_input_image_size_as_output`` is set; context frames of a size other than the output
        frame's are rejected by the model (the engine's plan assumes one latent geometry)."""
        if offload_model:
            raise NotImplementedError("offload_model is out of scope on B200 (SURVEY.md 2b)")
        prompt_img_len = len(input_images) if input_images is not None else 0
        if not use_input_image_size_as_output:
            assert height % 16 == 0 and width % 16 == 0, "The height and width must be a multiple of 16."
        if max_input_image_size != self.processor.max_image_size:
            self.processor = LVMProcessor(self.processor.text_tokenizer, max_image_size=max_input_image_size,
                                          sequence_parallel_size=max(hccl_info.world_size, 1))
        self.model.to(dtype)
        frames = list(input_images) if input_images is not None else None               # ori_input_images
        output_images = []
        for gen_idx in range(gen_num):
            if output_images:                                                            # 211-215
                frames = [output_images[-1]] if frames is None else frames + [output_images[-1]]
            ctx = []
            for idx, img in enumerate(frames or []):                                     # 254-260
                pixels = self.processor.process_image(img)
                lat = self.vae_encode(pixels.unsqueeze(0).to(self.device), dtype)
                if idx >= prompt_img_len:
                    lat = (1 - clean_image_noise_level) * lat + clean_image_noise_level * torch.randn_like(lat)
                ctx.append(lat)
            if ctx and use_input_image_size_as_output:                                   # 246-247
                height, width = ctx[0].shape[-2] * 8, ctx[0].shape[-1] * 8
            sample = self.next_frame_latents(
                ctx, height=height, width=width, num_inference_steps=num_inference_steps,
                img_guidance_scale=img_guidance_scale, use_img_guidance=use_img_guidance, seed=seed,
                time_shifting_factor=time_shifting_factor, prediction_type=prediction_type, dtype=dtype,
                use_kv_cache=use_kv_cache)
            if gen_idx == 0:
                output_images.extend(self.vae_decode_to_pil(lat) for lat in ctx)         # 305-316
            output_images.append(self.vae_decode_to_pil(sample))                         # 318-336
        gc.collect()
        return output_images

    @torch.no_grad()
    def prompt_condition_frame_block_autoregressive_inference(
            self, input_images=None, height: int = 1024, width: int = 1024, gen_nums: list = [1],
            num_inference_steps: int = 50, use_img_guidance: bool = True, img_guidance_scale: float = 1.6,
            max_input_image_size: int = 1024, offload_model: bool = False, use_kv_cache: bool = True,
            offload_kv_cache: bool = True, use_input_image_size_as_output: bool = False,
            dtype: torch.dtype = torch.bfloat16, seed: int = None, output_type: str = "pil",
            time_shifting_factor: float = 1.0, prediction_type: str = "v",
            clean_image_noise_level: float = None, max_frame_window: int = 16):
        """pipeline.py:346-595.  Returns the list of PIL frames (context reconstructions first).

        One deliberate fix: with guidance off the reference still returns ``samples[:len//2]``
        (pipeline.py:549) and silently drops half the generated frames; here
        all generated frames are returned."""
        if offload_model:
            raise NotImplementedError("offload_model is out of scope on B200 (SURVEY.md 2b)")
        if input_images is None:
            raise ValueError("next-clip prediction needs context frames")
        if not use_input_image_size_as_output:
            assert height % 16 == 0 and width % 16 == 0, "The height and width must be a multiple of 16."
        if max_input_image_size != self.processor.max_image_size:
            self.processor = LVMProcessor(self.processor.text_tokenizer, max_image_size=max_input_image_size,
                                          sequence_parallel_size=max(hccl_info.world_size, 1))
        self.model.to(dtype)
        output_images = []
        for k, gen_num in enumerate(gen_nums):
            if k > 0:
                input_images = output_images
            if len(input_images) + gen_num > max_frame_window:                           # 421-422
                input_images = input_images[gen_num + len(input_images) - max_frame_window:]
            pixels = [self.processor.process_image(x) for x in input_images]
            ctx = []
            for img in pixels:                                                            # 491-498
                lat = self.vae_encode(img.unsqueeze(0).to(self.device), dtype)
                if k > 0:
                    lat = (1 - clean_image_noise_level) * lat + clean_image_noise_level * torch.randn_like(lat)
                ctx.append(lat)
            samples = self.next_clip_latents(
                ctx, gen_num, num_inference_steps=num_inference_steps, img_guidance_scale=img_guidance_scale,
                use_img_guidance=use_img_guidance, seed=seed, time_shifting_factor=time_shifting_factor,
                prediction_type=prediction_type, dtype=dtype)
            if k == 0:
                output_images.extend(self.vae_decode_to_pil(lat) for lat in ctx)          # 558-570
            output_images.extend(self.vae_decode_to_pil(s) for s in samples)              # 572-590
        gc.collect()
        return output_images

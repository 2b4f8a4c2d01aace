"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): CFG branches split over two
ranks give bit-identical latents to the single-GPU run."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    from transformers import Phi3Config
    from oracle import processor_oracle as po
    from videogpt_b200 import LVM, LVMScheduler, parallel, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        dims = synth.REDUCED
        model = LVM(Phi3Config(**dims.phi3_kwargs()), device=dev)
        model.load_state_dict(synth.init_state_dict(dims, seed=0))
        model.to(torch.bfloat16).eval()
        n_ctx, n_gen, H, W, steps = 3, 2, 64, 96, 4
        d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
        lat = [x.to(dev, torch.bfloat16) for x in synth.synthetic_latents(n_ctx + n_gen, H, W, seed=42)]
        mk = dict(input_ids=d["input_ids"].to(dev), input_img_latents=lat[:n_ctx],
                  input_image_sizes=d["input_image_sizes"], attention_mask=None, position_ids=d["position_ids"].to(dev),
                  denoise_image_sizes=d["denoise_image_sizes"], time_emb_inx=d["time_emb_inx"], img_cfg_scale=1.5,
                  use_img_cfg=True, use_kv_cache=False, offload_model=False, vae=None)
        grp = parallel.CfgBranchGroup()
        out = {}
        for pt in ("x1", "v"):
            split = parallel.sample_cfg_split(model, LVMScheduler(steps), [x.clone() for x in lat[n_ctx:]] * 2, mk, grp, pt)
            single = LVMScheduler(steps)([x.clone() for x in lat[n_ctx:]] * 2, model.frame_block_forward_with_cfg, mk,
                                         prediction_type=pt)
            torch.cuda.synchronize()
            out[pt] = all(torch.equal(a, b) for a, b in zip(split, single[:n_gen]))
        ret[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_cfg_branch_split_matches_single_gpu():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert dict(ret) == {0: {"x1": True, "v": True}, 1: {"x1": True, "v": True}}

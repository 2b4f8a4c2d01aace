#!/usr/bin/env python
"""Probe of the peer-memory plumbing on N GPUs of one box (run under torchrun):
IPC export/import, NVLink stores into every peer, the flag barrier kernel (eager and from a CUDA
graph), barrier latency and peer-store bandwidth.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_probe.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videogpt_b200 import peer  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    grp = peer.PeerGroup(list(range(world)))
    n = 64 << 20
    buf = grp.alloc(n)
    views = []
    for r in range(world):                       # uint8 views of every rank's buffer in this process
        views.append(buf.local if r == rank else torch.as_tensor(peer._Blob(buf.ptrs[r], n, grp), device=dev))
    chunk = n // world
    ok = True
    for it in range(20):                         # every rank writes its chunk into ALL ranks, barrier, check
        for r in range(world):
            views[r][rank * chunk:(rank + 1) * chunk].fill_((it * 7 + rank + 1) % 251)
        grp.barrier()
        for src in range(world):
            want = (it * 7 + src + 1) % 251
            got = buf.local[src * chunk:(src + 1) * chunk]
            ok &= bool((got == want).all().item())
        grp.barrier()                            # nobody overwrites before everyone has checked
    grp.check()
    # barrier latency (eager launches back to back) and from a CUDA graph
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        grp.barrier()
    e1.record(); torch.cuda.synchronize()
    lat = e0.elapsed_time(e1) / 200 * 1e3
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        grp.barrier(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(50):
                grp.barrier()
    dist.barrier()
    e0.record()
    for _ in range(4):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    lat_g = e0.elapsed_time(e1) / 200 * 1e3
    grp.check()
    # peer store bandwidth: copy 48 MB into the next rank
    src = torch.ones(48 << 20, dtype=torch.uint8, device=dev)
    dst = views[(rank + 1) % world][:48 << 20]
    dst.copy_(src); torch.cuda.synchronize(); dist.barrier()
    e0.record()
    for _ in range(10):
        dst.copy_(src)
    e1.record(); torch.cuda.synchronize()
    bw = 10 * (48 << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9
    print(f"rank {rank}/{world}: data ok={ok}  barrier {lat:.1f} us eager, {lat_g:.1f} us in a graph;  "
          f"peer store {bw:.0f} GB/s", flush=True)
    dist.barrier()
    grp.close()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    t0 = time.time()
    main()

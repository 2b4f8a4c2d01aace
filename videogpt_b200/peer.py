"""Peer memory between the GPUs that share one video (sequence parallelism).

One process per GPU.  A ``PeerGroup`` lets every rank of a sequence-parallel group allocate a
buffer that all the other ranks of the group map into their own address space (CUDA IPC over
NVLink 5 / NVSwitch; ``csrc/peer.cu``).  The kernels that produce data every rank needs --
post-RoPE K / V rows (``vgpt_rope_kv_append_peers``) and the final-layer prediction
(``vgpt_final_layer_rows``) -- store it into all peers directly, and ``barrier()`` enqueues the
flag-exchange kernel that orders those stores before the consumers.  ``torch.distributed`` is
used once per allocation, on the host, to exchange the 64-byte IPC handles; nothing on the data
path calls NCCL.  This replaces the reference's DeepSpeed-Ulysses all-to-alls
(``LVM/transform/sdpa_transform.py:126-156``) and hidden-state all-gather (``LVM/model.py:466-474``).

``LocalPeerGroup`` gives the same interface to several *virtual* ranks inside one process on one
GPU (buffers are ordinary allocations, the barrier is stream order); the GPU tests use it to
prove that a row-sharded engine reproduces the unsharded one bit for bit.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch

from . import _lib


class _Blob:
    """``__cuda_array_interface__`` view of a raw device allocation (so torch can wrap it)."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 3, "strides": None}


class SharedBuffer:
    """``local``: this rank's bytes as a uint8 tensor; ``ptrs[r]``: device address of rank r's copy
    in THIS process (``ptrs[rank] == local.data_ptr()``)."""

    def __init__(self, local: torch.Tensor, ptrs: List[int], rank: int):
        self.local, self.ptrs, self.rank = local, list(ptrs), rank

    def ptr_array(self, offset_bytes: int = 0):
        """ctypes ``void*[world]`` of every rank's buffer + ``offset_bytes`` (kernel argument)."""
        return (ctypes.c_void_p * len(self.ptrs))(*[p + offset_bytes for p in self.ptrs])


def _raw_alloc(nbytes: int) -> int:
    out = ctypes.c_void_p()
    _lib.call("vgpt_peer_alloc", ctypes.byref(out), ctypes.c_uint64(nbytes))
    return int(out.value)


class PeerGroup:
    """The ranks ``ranks`` (global ranks of the default process group, this process included) that
    cooperate on one video.  Every method marked *collective* must be called by all of them in the
    same order."""

    def __init__(self, ranks: Sequence[int], group=None, device=None, host_group=None):
        """``host_group``: process group for the HOST-side collectives of this class (handle exchange,
        ``host_barrier``); defaults to ``group``.  A gloo group keeps them off the GPUs entirely."""
        import torch.distributed as dist
        self.ranks = list(ranks)
        self.world = len(self.ranks)
        if self.world > 8:
            raise ValueError("at most 8 ranks per sequence-parallel group (one NVSwitch domain)")
        self.rank = self.ranks.index(dist.get_rank())
        self.group = group if host_group is None else host_group
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._owned: List[int] = []
        self._imported: List[int] = []
        self._flags = self.alloc(4 * 8)
        self._state = torch.zeros(4, dtype=torch.int32, device=self.device)   # epoch, timed out, uint64 ns in barriers
        self._flag_ptrs = self._flags.ptr_array()

    # ---- collective -------------------------------------------------------------------------------
    def alloc(self, nbytes: int) -> SharedBuffer:
        """*collective*: zero-filled ``nbytes`` on every rank, mapped into every other rank."""
        import torch.distributed as dist
        nbytes = (int(nbytes) + 255) // 256 * 256
        with self._device_ctx():
            ptr = self._raw_alloc(nbytes)
            self._owned.append(ptr)
            handle = self._export(ptr)
            gathered: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(gathered, (handle, nbytes), group=self.group)
            # every rank validates before anybody maps anything (a size mismatch raises everywhere)
            for r, (_, n) in enumerate(gathered):
                if n != nbytes:
                    raise RuntimeError(f"PeerGroup.alloc: rank {r} asked for {n} bytes, this rank for {nbytes}")
            ptrs = []
            for r, (h, _) in enumerate(gathered):
                if r == self.rank:
                    ptrs.append(ptr)
                    continue
                mapped = self._import(h)
                self._imported.append(mapped)
                ptrs.append(mapped)
            local = self._wrap(ptr, nbytes)
        dist.barrier(group=self.group)
        return SharedBuffer(local, ptrs, self.rank)

    # ---- the four driver-facing steps (overridden by the CPU test double) -------------------------
    def _device_ctx(self):
        return torch.cuda.device(self.device)

    def _raw_alloc(self, nbytes: int) -> int:
        return _raw_alloc(nbytes)

    def _export(self, ptr: int) -> bytes:
        handle = (ctypes.c_ubyte * 64)()
        _lib.call("vgpt_peer_export", ctypes.c_void_p(ptr), handle)
        return bytes(handle)

    def _import(self, handle: bytes) -> int:
        out = ctypes.c_void_p()
        _lib.call("vgpt_peer_import", (ctypes.c_ubyte * 64).from_buffer_copy(handle), ctypes.byref(out))
        return int(out.value)

    def _wrap(self, ptr: int, nbytes: int) -> torch.Tensor:
        return torch.as_tensor(_Blob(ptr, nbytes, self), device=self.device)

    def _unmap(self, ptr: int):
        _lib.call("vgpt_peer_close", ctypes.c_void_p(ptr))

    def _free(self, ptr: int):
        _lib.call("vgpt_peer_free", ctypes.c_void_p(ptr))

    def _sync(self):
        torch.cuda.synchronize(self.device)

    def free(self, buf: SharedBuffer):
        """*collective*: release ONE buffer of ``alloc`` -- every rank unmaps its peers' copies, then (after a
        rendezvous: nobody frees while a peer still maps it) frees its own.  The engine calls it before it
        re-allocates a pool that a later plan has outgrown; without it every growth step of a rollout leaked a
        multi-GB pool on every rank until ``close``."""
        import torch.distributed as dist
        self._sync()
        dist.barrier(group=self.group)
        for r, p in enumerate(buf.ptrs):
            if r != self.rank and p in self._imported:
                self._unmap(p)
                self._imported.remove(p)
        dist.barrier(group=self.group)
        mine = buf.ptrs[self.rank]
        if mine in self._owned:
            self._free(mine)
            self._owned.remove(mine)
        buf.local, buf.ptrs = None, []

    def close(self):
        """*collective*: unmap the peers' buffers, then free this rank's."""
        import torch.distributed as dist
        self._sync()
        dist.barrier(group=self.group)
        for p in self._imported:
            self._unmap(p)
        self._imported = []
        dist.barrier(group=self.group)           # nobody frees while a peer still maps it
        for p in self._owned:
            self._free(p)
        self._owned = []

    # ---- data path --------------------------------------------------------------------------------
    def barrier(self):
        """Enqueue the cross-GPU barrier kernel on the current stream (graph capturable)."""
        _lib.call("vgpt_peer_barrier", self._flag_ptrs, self.world, self.rank,
                  ctypes.c_void_p(self._state.data_ptr()), torch.cuda.current_stream().cuda_stream)

    def host_barrier(self):
        """*collective*: drain this GPU, then rendezvous on the host.  Brackets every region in
        which a rank may block inside the CUDA driver (allocation, graph capture / instantiation):
        such calls can wait for a PEER device to go idle, and a peer that is spinning in
        ``vgpt_peer_barrier`` for kernels this rank has not enqueued yet never does."""
        import torch.distributed as dist
        self._sync()
        dist.barrier(group=self.group)

    def check(self):
        """Raise if any barrier timed out (a peer died); synchronises the stream."""
        st = self._state.cpu().tolist()
        if st[1] != 0:
            flags = self._flags.local[:4 * self.world].view(torch.int32).cpu().tolist()
            raise RuntimeError(f"vgpt_peer_barrier timed out on rank {self.rank} of {self.world}: this rank is at barrier "
                               f"epoch {st[0]}, the peers' last arrivals are {flags} (a rank of the sequence-parallel "
                               f"group did not arrive within ~10 s; results of this clip are invalid)")

    def barrier_seconds(self) -> float:
        """Device time spent inside barrier kernels so far (waiting for the slowest peer + the NVLink flag
        round trip); synchronises the stream."""
        return float(self._state[2:4].view(torch.int64).item()) * 1e-9

    lockstep = False


class LocalPeerGroup:
    """``world`` virtual ranks in one process on one GPU.  ``LocalPeerGroup.create(world)`` returns
    one member per virtual rank; the k-th ``alloc`` of every member refers to the same set of
    buffers.  ``barrier()`` does nothing: the caller issues the members' kernels phase by phase on
    one stream (``engine.run_lockstep``), so stream order is the barrier."""

    lockstep = True

    def __init__(self, rank: int, world: int, device, registry: list):
        self.rank, self.world, self.device = rank, world, torch.device(device)
        self._registry, self._calls = registry, 0

    @classmethod
    def create(cls, world: int, device=None) -> List["LocalPeerGroup"]:
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        registry: list = []
        return [cls(r, world, device, registry) for r in range(world)]

    def alloc(self, nbytes: int) -> SharedBuffer:
        nbytes = (int(nbytes) + 255) // 256 * 256
        k = self._calls
        self._calls += 1
        if k == len(self._registry):
            self._registry.append([torch.zeros(nbytes, dtype=torch.uint8, device=self.device) for _ in range(self.world)])
        bufs = self._registry[k]
        if bufs[0].numel() != nbytes:
            raise RuntimeError("LocalPeerGroup.alloc: members disagree on the allocation size")
        return SharedBuffer(bufs[self.rank], [b.data_ptr() for b in bufs], self.rank)

    def free(self, buf: SharedBuffer):
        buf.local, buf.ptrs = None, []        # torch owns the virtual ranks' buffers

    def barrier(self):
        pass

    def host_barrier(self):
        pass

    def check(self):
        pass

    def close(self):
        pass

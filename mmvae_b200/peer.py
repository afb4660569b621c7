"""Symmetric peer-memory buffers for the data-parallel step (one process per GPU, NVLink / NVSwitch).

``PeerComm.alloc(name, nbytes)`` is a collective: every rank allocates the same number of bytes, exports the
allocation through CUDA IPC (torch's own tensor-sharing plumbing, ``torch.multiprocessing.reductions``) and maps
every peer's allocation into its address space.  The result is a ``SymBuf``: ``local`` (this rank's bytes), ``ptr[r]``
(device address of rank r's copy as seen from THIS process -- peer memory for r != rank).  Kernels of
``csrc/peer.cu`` / the routed epilogues store straight into ``ptr[r]``.  torch.distributed is used for the handle
exchange and barriers only; no collective ever carries step data.

Flags: one uint32 per (channel, source rank) in every rank's ``flags`` buffer.  A producer raises
``flags[channel][my_rank]`` on every consumer to the current step number; consumers wait until all N entries of the
channel have reached it.  Step numbers only grow, so nothing is ever reset.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class SymBuf:
    def __init__(self, name: str, local: torch.Tensor, peers: List[torch.Tensor], rank: int):
        self.name, self.local, self.peers, self.rank = name, local, peers, rank
        self.ptr = [int(t.data_ptr()) for t in peers]
        self.nbytes = local.numel()

    def view(self, dtype, offset: int = 0, count: Optional[int] = None) -> torch.Tensor:
        """typed view of this rank's bytes"""
        t = self.local[offset:] if count is None else self.local[offset:offset + count * torch.empty(0, dtype=dtype).element_size()]
        return t.view(dtype)


class PeerComm:
    N_CHANNELS = 64

    def __init__(self, device, group=None):
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > 16:
            raise RuntimeError("peer-memory data parallelism supports up to 16 GPUs of one NVSwitch domain")
        self.bufs: Dict[str, SymBuf] = {}
        self._keep = []
        self.flags = self.alloc("flags", self.N_CHANNELS * 16 * 4, zero=True)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._channels: Dict[str, int] = {}

    def alloc(self, name: str, nbytes: int, zero: bool = False) -> SymBuf:
        """collective: same name / size / order on every rank"""
        nbytes = (int(nbytes) + 255) // 256 * 256
        # a dedicated cudaMalloc block per buffer (not a slice of a cached segment): the IPC handle then covers
        # exactly this buffer and the exported offset is 0
        local = torch.empty(max(nbytes, 2 << 20), dtype=torch.uint8, device=self.device)[:nbytes]
        if zero:
            local.zero_()
        torch.cuda.synchronize(self.device)
        if self.world == 1:
            peers = [local]
        else:
            from torch.multiprocessing.reductions import reduce_tensor
            fn, args = reduce_tensor(local)
            gathered = [None] * self.world
            dist.all_gather_object(gathered, (name, nbytes, args), group=self.group)
            peers = []
            for r, (n_r, b_r, a_r) in enumerate(gathered):
                if n_r != name or b_r != nbytes:
                    raise RuntimeError(f"PeerComm.alloc mismatch: rank {r} allocates {n_r}/{b_r}, rank {self.rank} "
                                       f"{name}/{nbytes}")
                if r == self.rank:
                    peers.append(local)
                else:
                    a_r = list(a_r)
                    a_r[6] = self.device.index if self.device.index is not None else torch.cuda.current_device()
                    peers.append(fn(*a_r))       # cudaIpcOpenMemHandle under the hood; peer access enabled lazily
            dist.barrier(group=self.group)
        buf = SymBuf(name, local, peers, self.rank)
        self.bufs[name] = buf
        return buf

    # ---- flags ----
    def channel(self, name: str) -> int:
        if name not in self._channels:
            if len(self._channels) >= self.N_CHANNELS:
                raise RuntimeError("out of flag channels")
            self._channels[name] = len(self._channels)
        return self._channels[name]

    def flag_ptrs(self, channel: str) -> List[int]:
        """address, on every rank r, of flags[channel][my_rank]"""
        c = self.channel(channel)
        return [p + (c * 16 + self.rank) * 4 for p in self.flags.ptr]

    def local_flags(self, channel: str) -> torch.Tensor:
        c = self.channel(channel)
        return self.flags.local[(c * 16) * 4:(c * 16 + 16) * 4].view(torch.int32)

    def barrier(self):
        if self.world > 1:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)

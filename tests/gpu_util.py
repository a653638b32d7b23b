"""GPU-side test scaffolding: device twins of oracle layouts, driven through the C ABI via ctypes.

torch is used only as the device allocator / stream provider (plumbing)."""
from __future__ import annotations

import numpy as np
import torch

from dynamo_b200 import kernels as K
from oracle import oracle as O


def stream_ptr(stream=None) -> int:
    s = stream if stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)


def dev_u8(n: int, device="cuda:0", fill=0) -> torch.Tensor:
    return torch.full((max(1, n),), fill, dtype=torch.uint8, device=device)


def dev_ptr_table(ptrs, device="cuda:0") -> torch.Tensor:
    return torch.tensor([int(p) for p in ptrs], dtype=torch.int64, device=device)


def pinned_u8(arr: np.ndarray) -> torch.Tensor:
    t = torch.empty(arr.size, dtype=torch.uint8).pin_memory()
    t.numpy()[:] = arr.reshape(-1)
    return t


class DevicePool:
    """Device twin of an oracle Layout: same geometry, one device allocation per host buffer."""

    def __init__(self, host: O.Layout, device="cuda:0", copy=True):
        self.host = host
        self.device = device
        self.bufs = [torch.from_numpy(b.copy()).to(device) if copy else dev_u8(b.size, device) for b in host.buffers]
        nl = host.num_layers
        if host.kind == O.FC:
            base = self.bufs[0].data_ptr()
            bases = [base + l * host.layer_stride for l in range(nl)]
        else:
            bases = [b.data_ptr() for b in self.bufs]
        self.layer_base = torch.tensor(bases, dtype=torch.int64, device=device)
        self.desc = K.PagedLayout(self.layer_base.data_ptr(), host.block_stride, host.outer_stride,
                                  host.region_size, nl, host.outer_dim, host.num_blocks)

    def region_addr(self, b, l, o) -> int:
        a, _ = self.host.memory_region(b, l, o)
        if self.host.kind == O.FC:
            return self.bufs[0].data_ptr() + (a - self.host.buffers[0].ctypes.data)
        return self.bufs[l].data_ptr() + (a - self.host.buffers[l].ctypes.data)

    def download(self):
        """Copy device contents back into the host twin's numpy buffers."""
        for hb, db in zip(self.host.buffers, self.bufs):
            hb[:] = db.cpu().numpy()

    def snapshot(self):
        return [db.cpu().numpy().copy() for db in self.bufs]


def make_layout(kind, nb, nl=2, no=2, page=16, inner=128, dt=2, block_dim=O.BLOCK_IS_FIRST_DIM, fill=0,
                allow_fp8=False) -> O.Layout:
    return O.Layout(kind, nb, nl, no, page, inner, dt, block_dim=block_dim, fill=fill, allow_fp8=allow_fp8)


def randomize(layout: O.Layout, seed: int):
    rng = np.random.default_rng(seed)
    for b in layout.buffers:
        b[:] = rng.integers(0, 256, b.size, dtype=np.uint8)


def ids_dev(ids, device="cuda:0") -> torch.Tensor:
    return torch.tensor([int(i) for i in ids], dtype=torch.int32, device=device)


def paged_dst(pool: DevicePool, src_ids_t: torch.Tensor, dst_ids_t: torch.Tensor, done_flag=0, layer_done=0):
    return K.PagedDst(pool.desc, src_ids_t.data_ptr(), dst_ids_t.data_ptr(), done_flag, layer_done)

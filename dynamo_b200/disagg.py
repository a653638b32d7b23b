"""Prefill -> decode hand-off orchestration across processes (one process per GPU / worker).

This is the control-plane sliver the data path needs and nothing more: who pushes to whom, and the
exchange of layout metadata (the reference exchanges `SerializedLayout` blobs the same way:
lib/kvbm-physical/src/manager/metadata.rs:87-159, manager/mod.rs:112-147).  It rides on whatever
`torch.distributed` backend the job already has (nccl on GPUs, gloo in CPU tests); no data-path collective
is used -- blocks are independent units pushed one-sided.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

from .physical import TransferCompleteNotification, TransferManager, TransferOptions


@dataclass
class Roles:
    sources: List[int]
    destinations: Dict[int, List[int]]    # source rank -> destination ranks

    def is_source(self, rank: int) -> bool:
        return rank in self.destinations

    def is_destination(self, rank: int) -> bool:
        return any(rank in d for d in self.destinations.values())

    def source_of(self, rank: int) -> Optional[int]:
        for s, ds in self.destinations.items():
            if rank in ds:
                return s
        return None


def assign_roles(world: int, topology: str = "fanout") -> Roles:
    """fanout: rank 0 is the prefill GPU, ranks 1..world-1 decode (1 -> N-1; world 1 = loop-back).
    pairs:  rank r < world/2 pushes to rank r + world/2 (TP-sharded prefill -> decode, BASELINE configs[3])."""
    if world < 1:
        raise ValueError("world must be >= 1")
    if topology == "fanout":
        return Roles([0], {0: list(range(1, world)) if world > 1 else [0]})
    if topology == "pairs":
        if world == 1:
            return Roles([0], {0: [0]})
        if world % 2:
            raise ValueError("pairs topology needs an even world size")
        half = world // 2
        return Roles(list(range(half)), {r: [r + half] for r in range(half)})
    raise ValueError(f"unknown topology {topology!r}")


class HandoffGroup:
    """Exchanges layout metadata so every source holds a (peer-mapped) handle to its destinations' pools."""

    def __init__(self, mgr: TransferManager, rank: int, world: int, topology: str = "fanout", group=None):
        self.mgr, self.rank, self.world, self.group = mgr, rank, world, group
        self.roles = assign_roles(world, topology)
        self.remote: Dict[int, int] = {}     # destination rank -> handle usable from this process

    def publish(self, local_dst_handle: Optional[int], local_views: Optional[Dict[int, Sequence[int]]] = None) -> Dict[int, int]:
        """Collective: every rank contributes the metadata of the pool it receives into (or None).
        `local_views[rank]` = this process's own mappings (one address per allocation) of that rank's HOST pool when the pools
        live in shared memory; device pools travel as CUDA IPC handles inside the metadata and need nothing here."""
        blob = self.mgr.export_metadata(local_dst_handle) if local_dst_handle is not None else b""
        if self.world == 1:
            blobs = [blob]
        else:
            import torch.distributed as dist
            blobs = [None] * self.world
            dist.all_gather_object(blobs, blob, group=self.group)
        self.blobs = blobs
        for dst in self.roles.destinations.get(self.rank, []):
            if dst == self.rank:
                self.remote[dst] = local_dst_handle
            else:
                if not blobs[dst]:
                    raise RuntimeError(f"rank {dst} published no layout")
                self.remote[dst] = self.mgr.import_metadata(blobs[dst], (local_views or {}).get(dst))
        return self.remote

    def push(self, src_handle: int, src_block_ids: Sequence[Sequence[int]], dst_block_ids: Sequence[Sequence[int]],
             replicate: bool = False, options: Optional[TransferOptions] = None) -> TransferCompleteNotification:
        """Source side: one launch to all of this rank's destinations (list i <-> i-th destination rank)."""
        dsts = [self.remote[d] for d in self.roles.destinations[self.rank]]
        if len(dsts) == 1 and not replicate:
            return self.mgr.execute_transfer(src_handle, src_block_ids[0], dsts[0], dst_block_ids[0], options)
        return self.mgr.execute_fanout(src_handle, dsts, src_block_ids, dst_block_ids, replicate, options)


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing reduction the bench contract asks for (max over ranks)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def share_fd(rank: int, world: int, fd: Optional[int], name: str, timeout_s: float = 60.0) -> int:
    """Hand one POSIX file descriptor from rank 0 to every other rank of the node (SCM_RIGHTS over an abstract
    AF_UNIX socket named after `name`).  Used for the shareable handle of a multicast object
    (`MulticastGroup.export_fd` -> `MulticastGroup.from_fd`); torch.distributed cannot carry descriptors."""
    import socket
    import time
    addr = "\0kvbm-fd-" + name
    if rank == 0:
        if fd is None:
            raise ValueError("rank 0 must provide the descriptor")
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        srv.bind(addr)
        srv.listen(world)
        srv.settimeout(timeout_s)
        try:
            for _ in range(world - 1):
                conn, _ = srv.accept()
                with conn:
                    socket.send_fds(conn, [b"fd"], [fd])
        finally:
            srv.close()
        return fd
    deadline = time.monotonic() + timeout_s
    while True:
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        try:
            c.connect(addr)
        except (FileNotFoundError, ConnectionRefusedError):
            c.close()
            if time.monotonic() > deadline:
                raise TimeoutError(f"rank 0 never offered descriptor {name!r}")
            time.sleep(0.02)
            continue
        with c:
            c.settimeout(timeout_s)
            _, fds, _, _ = socket.recv_fds(c, 16, 1)
        if not fds:
            raise RuntimeError("no descriptor received")
        return fds[0]


# ----------------------------------------------------------------------------------------------------------------------
# Staged NVLS broadcast: identical KV blocks to N decode workers that each allocate their OWN destination blocks
# ----------------------------------------------------------------------------------------------------------------------
def staged_send(mgr: TransferManager, src_handle: int, src_block_ids: Sequence[int], staging_mc_handle: int,
                receiver_ready_flags: Sequence[int], epoch: int, stream: int,
                receiver_free_flags: Optional[Sequence[int]] = None) -> TransferCompleteNotification:
    """Root half of `CollectiveOps::broadcast` (lib/kvbm-engine/src/collectives/mod.rs:99-106, nccl.rs:421-462) for
    receivers with distinct block tables.  The multicast mode needs the same offsets in every bound pool, which a real
    decode worker's own allocations never give; so the payload is multicast ONCE into a small staging pool every receiver
    bound to the group (blocks 0..n-1), and each receiver scatters it locally (`staged_receive`).

    `staging_mc_handle`   layout registered over the group's multicast mapping (`MulticastGroup.map`)
    `receiver_ready_flags[r]`  address (peer-mapped into this process) of receiver r's per-layer flag array: the kernel
                          sets entry l to `epoch` when layer l has landed in EVERY staging pool
    `receiver_free_flags[r]`   address on THIS GPU that receiver r's local scatter sets to the epoch it has consumed:
                          the next send waits for them on the device before it overwrites the staging pool"""
    from . import kernels as K
    n = len(src_block_ids)
    nr = len(receiver_ready_flags)
    if receiver_free_flags is not None and epoch > 1:
        for p in receiver_free_flags:
            K.check(K.wait_flag(int(p), epoch - 1, stream), "wait_flag(free)")
    stage_ids = list(range(n))
    opts = TransferOptions(multicast=1, epoch=epoch, cuda_stream=stream, per_dst_layer_done_flags=[int(p) for p in receiver_ready_flags])
    return mgr.execute_fanout(src_handle, [staging_mc_handle] * nr, [src_block_ids] * nr, [stage_ids] * nr, True, opts)


def staged_receive(mgr: TransferManager, staging_local_handle: int, dst_handle: int, dst_block_ids: Sequence[int],
                   ready_flags: int, epoch: int, free_flag: int = 0, max_ctas: int = 0) -> TransferCompleteNotification:
    """Receiver half: ONE gated launch on the receiver's own GPU copies staging block i -> dst_block_ids[i], layer by
    layer as the root's flags arrive (HBM -> HBM, hidden behind the multicast), then reports `epoch` to `free_flag`
    (an address in the root's memory, peer-mapped here)."""
    stage_ids = list(range(len(dst_block_ids)))
    opts = TransferOptions(layer_ready_flags=int(ready_flags), epoch=epoch, done_flag=int(free_flag), max_ctas=max_ctas)
    return mgr.execute_transfer(staging_local_handle, stage_ids, dst_handle, dst_block_ids, opts)

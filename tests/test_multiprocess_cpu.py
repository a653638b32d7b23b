"""world_size-2 gloo test of the N>1 hand-off on CPU (BASELINE configs[0]: 1 prefill + 1 decode worker,
CPU-only memcpy KV hand-off -- plumbing, runs without a GPU).

Two forked processes rendezvous over gloo (127.0.0.1).  The decode rank registers its pool and publishes
layout metadata; the prefill rank imports it and pushes 8 blocks with the Memcpy strategy; the decode rank
verifies BLAKE3 block checksums against the oracle's Sequential fill.  The pools live in MAP_SHARED
anonymous mappings created before fork; the prefill rank names its own view of the decode pool when it imports
the metadata (`import_metadata(blob, local_bases)`), the CPU twin of the CUDA IPC peer mapping of the GPU path."""
import mmap
import multiprocessing as mp
import os
import socket

import numpy as np
import pytest

from dynamo_b200.disagg import HandoffGroup, assign_roles
from oracle import oracle as O

NB, NL, NO, PAGE, INNER, DT = 16, 3, 2, 16, 128, 2
REGION = PAGE * INNER * DT
PER_LAYER = NB * NO * REGION


def test_role_assignment():
    r = assign_roles(1)
    assert r.sources == [0] and r.destinations == {0: [0]}
    r = assign_roles(8)
    assert r.destinations == {0: [1, 2, 3, 4, 5, 6, 7]} and r.is_source(0) and r.is_destination(7) and r.source_of(3) == 0
    r = assign_roles(8, "pairs")
    assert r.destinations == {0: [4], 1: [5], 2: [6], 3: [7]} and r.source_of(6) == 2 and not r.is_source(5)
    with pytest.raises(ValueError):
        assign_roles(3, "pairs")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, port, src_addrs, dst_addrs, q):
    try:
        import torch.distributed as dist
        from dynamo_b200.physical import BlockDimension, LayoutConfig, StorageKind, TransferManager
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=2)
        mgr = TransferManager(device=-1, worker_id=rank + 1)
        cfg = LayoutConfig(NB, NL, NO, PAGE, INNER, dtype_width_bytes=DT)
        grp = HandoffGroup(mgr, rank, 2, "fanout")
        sid, did = [0, 5, 2, 9, 11, 3, 7, 15], [8, 1, 14, 0, 6, 12, 4, 10]
        if rank == 0:   # prefill worker
            h_src = mgr.register_layer_separate(cfg, src_addrs, [PER_LAYER] * NL, BlockDimension.BlockIsSecondDim, StorageKind.System)
            O.Layout(O.LW, NB, NL, NO, PAGE, INNER, DT, block_dim=O.BLOCK_IS_SECOND_DIM, bases=src_addrs).fill_blocks(range(NB), -1)
            grp.publish(None, local_views={1: dst_addrs})     # the decode pool is shared memory, mapped here at dst_addrs
            note = grp.push(h_src, [sid], [did])
            note.wait()
            dist.barrier()
            q.put((rank, "ok", mgr.bytes_moved()))
        else:           # decode worker
            h_dst = mgr.register_layer_separate(cfg, dst_addrs, [PER_LAYER] * NL, BlockDimension.BlockIsSecondDim, StorageKind.Pinned)
            grp.publish(h_dst)
            dist.barrier()   # prefill has pushed
            twin = O.Layout(O.LW, NB, NL, NO, PAGE, INNER, DT, block_dim=O.BLOCK_IS_SECOND_DIM, bases=dst_addrs)
            ref = O.Layout(O.LW, NB, NL, NO, PAGE, INNER, DT, block_dim=O.BLOCK_IS_SECOND_DIM)
            ref.fill_blocks(range(NB), -1)
            good = all(twin.block_checksum(d) == ref.block_checksum(s) for s, d in zip(sid, did))
            untouched = sorted(set(range(NB)) - set(did))
            clean = all(not twin.region_bytes(b, 0, 0).any() for b in untouched)
            q.put((rank, "ok" if good and clean else "mismatch", 0))
        dist.destroy_process_group()
        mgr.close()
    except Exception as e:  # pragma: no cover
        q.put((rank, f"error: {e!r}", 0))


def test_two_process_gloo_handoff():
    ctx = mp.get_context("fork")
    maps = [mmap.mmap(-1, PER_LAYER) for _ in range(2 * NL)]     # MAP_SHARED | MAP_ANONYMOUS
    addrs = [np.frombuffer(m, dtype=np.uint8).ctypes.data for m in maps]
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, addrs[:NL], addrs[NL:], q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict()
    for _ in range(2):
        rank, status, moved = q.get(timeout=120)
        results[rank] = (status, moved)
    for p in procs:
        p.join(timeout=30)
    assert results[0][0] == "ok" and results[1][0] == "ok", results
    assert results[0][1] == 8 * NL * NO * REGION


def _fd_worker(rank, world, name, path, q):
    from dynamo_b200.disagg import share_fd
    if rank == 0:
        fd = os.open(path, os.O_RDONLY)
        share_fd(0, world, fd, name)
        q.put((0, "sent"))
    else:
        fd = share_fd(rank, world, None, name)
        q.put((rank, os.pread(fd, 64, 0)))


def test_share_fd_hands_a_descriptor_to_every_rank(tmp_path):
    """The multicast object's shareable handle is a POSIX fd: it has to cross processes over SCM_RIGHTS."""
    import multiprocessing as mp
    p = tmp_path / "token"
    p.write_bytes(b"multicast-handle-standin")
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    name = f"test-{os.getpid()}"
    procs = [ctx.Process(target=_fd_worker, args=(r, 3, name, str(p), q)) for r in range(3)]
    for pr in procs:
        pr.start()
    got = dict(q.get(timeout=30) for _ in range(3))
    for pr in procs:
        pr.join(30)
        assert pr.exitcode == 0
    assert got[0] == "sent" and got[1] == got[2] == b"multicast-handle-standin"

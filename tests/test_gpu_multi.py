"""Peer (NVLink) tests: need >= 2 GPUs (`gpurun --gpus 2`).  Same-process peer mappings and the
cross-process CUDA-IPC path that replaces the NIXL/UCX hop."""
import os
import socket

import numpy as np
import pytest
import torch

from dynamo_b200 import kernels as K
from dynamo_b200.physical import BlockDimension, LayoutConfig, StorageKind, TransferManager, TransferOptions
from oracle import oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

NB, NL, NO, PAGE, INNER, DT = 64, 4, 2, 16, 1024, 2
REGION = PAGE * INNER * DT


def _pool(dev):
    return [torch.zeros(NO * NB * REGION, dtype=torch.uint8, device=f"cuda:{dev}") for _ in range(NL)]


def _register(mgr, bufs, dev):
    cfg = LayoutConfig(NB, NL, NO, PAGE, INNER, dtype_width_bytes=DT)
    return mgr.register_layer_separate(cfg, [b.data_ptr() for b in bufs], [b.numel() for b in bufs],
                                       BlockDimension.BlockIsSecondDim, StorageKind.Device, dev)


def _twin(bufs):
    t = O.Layout(O.LW, NB, NL, NO, PAGE, INNER, DT, block_dim=O.BLOCK_IS_SECOND_DIM)
    for hb, db in zip(t.buffers, bufs):
        hb[:] = db.cpu().numpy()
    return t


def test_same_process_peer_push_and_fanout():
    ndev = torch.cuda.device_count()
    mgr = TransferManager(device=0, worker_id=1)
    src = _pool(0)
    g = torch.Generator(device="cuda:0").manual_seed(3)
    for b in src:
        b.copy_(torch.randint(0, 256, b.shape, dtype=torch.uint8, device="cuda:0", generator=g))
    h_src = _register(mgr, src, 0)
    dsts, hs = [], []
    for d in range(1, ndev):
        mgr.enable_peer_access(d)
        bufs = _pool(d)
        dsts.append(bufs)
        hs.append(_register(mgr, bufs, d))
    rng = np.random.default_rng(0)
    n = 24
    sids = [list(map(int, rng.permutation(NB)[:n])) for _ in hs]
    dids = [list(map(int, rng.permutation(NB)[:n])) for _ in hs]
    if len(hs) == 1:
        mgr.execute_transfer(h_src, sids[0], hs[0], dids[0]).wait()
    else:
        mgr.execute_fanout(h_src, hs, sids, dids).wait()
    src_t = _twin(src)
    for bufs, sid, did in zip(dsts, sids, dids):
        ref = O.Layout(O.LW, NB, NL, NO, PAGE, INNER, DT, block_dim=O.BLOCK_IS_SECOND_DIM)
        O.execute_memcpy_transfer(src_t, ref, sid, did)
        got = _twin(bufs)
        for a, b in zip(got.buffers, ref.buffers):
            assert np.array_equal(a, b)
    # replicate (broadcast) to every peer: one HBM read, N stores
    for bufs in dsts:
        for b in bufs:
            b.zero_()
    mgr.broadcast(h_src, hs, sids[0], dids[0]).wait()
    for bufs in dsts:
        ref = O.Layout(O.LW, NB, NL, NO, PAGE, INNER, DT, block_dim=O.BLOCK_IS_SECOND_DIM)
        O.execute_memcpy_transfer(src_t, ref, sids[0], dids[0])
        for a, b in zip(_twin(bufs).buffers, ref.buffers):
            assert np.array_equal(a, b)
    mgr.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ipc_worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        from dynamo_b200.disagg import HandoffGroup
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        mgr = TransferManager(device=rank, worker_id=rank + 1)
        grp = HandoffGroup(mgr, rank, world, "fanout")
        n = 20
        sid = [list(map(int, np.random.default_rng(10 + d).permutation(NB)[:n])) for d in range(world - 1)]
        did = [list(map(int, np.random.default_rng(50 + d).permutation(NB)[:n])) for d in range(world - 1)]
        done = torch.zeros(NL, dtype=torch.int32, device=f"cuda:{rank}")
        if rank == 0:
            src = _pool(0)
            gen = torch.Generator(device="cuda:0").manual_seed(77)
            for b in src:
                b.copy_(torch.randint(0, 256, b.shape, dtype=torch.uint8, device="cuda:0", generator=gen))
            h_src = _register(mgr, src, 0)
            grp.publish(None)
            grp.push(h_src, sid, did).wait()
            sums = [[int(src[l].view(NO, NB, REGION)[:, s].sum(dtype=torch.int64)) for s in sid[d]] for d in range(world - 1) for l in (0, NL - 1)]
            box = [sums]
            dist.broadcast_object_list(box, src=0)
            dist.barrier()
            q.put((rank, "ok"))
        else:
            dst = _pool(rank)
            h = _register(mgr, dst, rank)
            grp.publish(h)
            box = [None]
            dist.broadcast_object_list(box, src=0)     # arrives after rank 0's transfer completed
            d = rank - 1
            torch.cuda.synchronize()
            good = True
            for k, l in enumerate((0, NL - 1)):
                mine = [int(dst[l].view(NO, NB, REGION)[:, b].sum(dtype=torch.int64)) for b in did[d]]
                good = good and mine == box[0][d * 2 + k]
            untouched = sorted(set(range(NB)) - set(did[d]))
            good = good and not bool(dst[0].view(NO, NB, REGION)[:, untouched].any())
            dist.barrier()
            q.put((rank, "ok" if good else "mismatch"))
        dist.destroy_process_group()
        mgr.close()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "error: " + traceback.format_exc()))


def test_cross_process_ipc_push():
    """One process per GPU; the decode process exports CUDA IPC metadata, the prefill process maps it and the
    kernel stores straight into the other process's KV pool."""
    import torch.multiprocessing as tmp
    world = min(torch.cuda.device_count(), 4)
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, s = q.get(timeout=300)
        res[r] = s
    for p in procs:
        p.join(timeout=60)
    assert all(v == "ok" for v in res.values()), res


def test_pull_direction_decode_side_reads_prefill_pool():
    """NIXL-READ style (what vLLM's NixlConnector does, components/src/dynamo/vllm/handlers.py:2076-2083): the manager
    on the DECODE GPU maps the prefill pool and the same kernel, launched there, loads over NVLink and stores locally."""
    mgr = TransferManager(device=1, worker_id=2)
    mgr.enable_peer_access(0)
    src = _pool(0)
    g = torch.Generator(device="cuda:0").manual_seed(9)
    for b in src:
        b.copy_(torch.randint(0, 256, b.shape, dtype=torch.uint8, device="cuda:0", generator=g))
    dst = _pool(1)
    h_src, h_dst = _register(mgr, src, 0), _register(mgr, dst, 1)
    rng = np.random.default_rng(4)
    sid, did = list(map(int, rng.permutation(NB)[:30])), list(map(int, rng.permutation(NB)[:30]))
    torch.cuda.synchronize(0)
    mgr.execute_transfer(h_src, sid, h_dst, did).wait()
    ref = O.Layout(O.LW, NB, NL, NO, PAGE, INNER, DT, block_dim=O.BLOCK_IS_SECOND_DIM)
    O.execute_memcpy_transfer(_twin(src), ref, sid, did)
    for a, b in zip(_twin(dst).buffers, ref.buffers):
        assert np.array_equal(a, b)
    mgr.close()


def test_pull_with_fused_upcast_ships_fp8_over_nvlink():
    """Receiver-side up-cast: the decode GPU pulls the prefill GPU's fp8 pool and widens it to bf16 while
    storing locally, so only HALF the bytes cross NVLink.  Same kernel, launched on the destination, cast mode 1."""
    mgr = TransferManager(device=1, worker_id=3)
    mgr.enable_peer_access(0)
    region8 = PAGE * INNER          # fp8 source regions are half the size
    src = [torch.zeros(NO * NB * region8, dtype=torch.uint8, device="cuda:0") for _ in range(NL)]
    g = torch.Generator(device="cuda:0").manual_seed(10)
    for b in src:
        b.copy_(torch.randint(0, 256, b.shape, dtype=torch.uint8, device="cuda:0", generator=g))
    dst = _pool(1)
    cfg8 = LayoutConfig(NB, NL, NO, PAGE, INNER, dtype_width_bytes=1, allow_fp8=True)
    h_src = mgr.register_layer_separate(cfg8, [b.data_ptr() for b in src], [b.numel() for b in src],
                                        BlockDimension.BlockIsSecondDim, StorageKind.Device, 0)
    h_dst = _register(mgr, dst, 1)
    rng = np.random.default_rng(5)
    sid, did = list(map(int, rng.permutation(NB)[:30])), list(map(int, rng.permutation(NB)[:30]))
    torch.cuda.synchronize(0)
    mgr.execute_transfer(h_src, sid, h_dst, did, TransferOptions(cast_mode=K.CastMode.FP8E4M3_TO_BF16)).wait()
    src_t = O.Layout(O.LW, NB, NL, NO, PAGE, INNER, 1, block_dim=O.BLOCK_IS_SECOND_DIM, allow_fp8=True)
    for hb, db in zip(src_t.buffers, src):
        hb[:] = db.cpu().numpy()
    ref = O.Layout(O.LW, NB, NL, NO, PAGE, INNER, DT, block_dim=O.BLOCK_IS_SECOND_DIM)
    O.execute_memcpy_transfer(src_t, ref, sid, did, cast_mode=1)
    for a, b in zip(_twin(dst).buffers, ref.buffers):
        assert np.array_equal(a, b)
    mgr.close()

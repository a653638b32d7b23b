"""NVLS multicast replicate (needs >= 2 GPUs behind an NVSwitch with multicast support): identical KV blocks written
ONCE by the source GPU land in every pool bound to the multicast group, byte-exact with the oracle's transfer --
the replacement of the grouped ncclBcast per region (lib/kvbm-engine/src/collectives/nccl.rs:421-462)."""
import numpy as np
import pytest
import torch

from dynamo_b200 import kernels as K
from dynamo_b200.physical import (BlockDimension, KvbmError, LayoutConfig, MulticastGroup, StorageKind, TransferManager,
                                  TransferOptions, multicast_supported)
from oracle import oracle as O
from tests.gpu_util import stream_ptr

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu, pytest.mark.timeout(200)]

NB, NL, NO, PAGE, INNER, DT = 64, 4, 2, 16, 1024, 2
REGION = PAGE * INNER * DT
PER_LAYER = NO * NB * REGION
TOTAL = NL * PER_LAYER


class _Raw:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _view(ptr, nbytes, dev):
    with torch.cuda.device(dev):
        return torch.as_tensor(_Raw(ptr, nbytes), device=f"cuda:{dev}")


@pytest.fixture(scope="module")
def group():
    ndev = torch.cuda.device_count()
    if ndev < 2 or not all(multicast_supported(d) for d in range(ndev)):
        pytest.skip("needs >= 2 GPUs with NVLink multicast support")
    for d in range(ndev):
        torch.zeros(1, device=f"cuda:{d}")
    torch.cuda.set_device(0)
    g = MulticastGroup.create(ndev, TOTAL)
    for d in range(ndev):
        g.add_device(d)
    pools = []
    for d in range(ndev):
        t = _view(g.bind_local(d), TOTAL, d)
        t.zero_()
        pools.append(t)
    for d in range(ndev):
        torch.cuda.synchronize(d)
    mc = g.map(0)
    yield g, pools, mc
    for d in range(ndev):
        torch.cuda.synchronize(d)
    del pools
    g.close()   # note: tearing a multicast object down takes the driver tens of seconds; one group serves the module


def _src_pool(seed):
    gen = torch.Generator(device="cuda:0").manual_seed(seed)
    return [torch.randint(0, 256, (PER_LAYER,), dtype=torch.uint8, device="cuda:0", generator=gen) for _ in range(NL)]


def _twin(bufs=None):
    t = O.Layout(O.LW, NB, NL, NO, PAGE, INNER, DT, block_dim=O.BLOCK_IS_SECOND_DIM)
    if bufs is not None:
        for hb, db in zip(t.buffers, bufs):
            hb[:] = db.cpu().numpy()
    return t


def _cfg():
    return LayoutConfig(NB, NL, NO, PAGE, INNER, dtype_width_bytes=DT)


def _expect(src, sid, did, layers=None):
    want = _twin()
    O.execute_memcpy_transfer(_twin(src), want, sid, did, layer_range=layers)
    return np.concatenate(want.buffers)


def _reset(pools):
    for d, p in enumerate(pools):
        p.zero_()
        torch.cuda.synchronize(d)


def test_manager_multicast_transfer_reaches_every_bound_pool(group):
    g, pools, mc = group
    _reset(pools)
    mgr = TransferManager(device=0, worker_id=7)
    src = _src_pool(11)
    h_src = mgr.register_layer_separate(_cfg(), [b.data_ptr() for b in src], [b.numel() for b in src],
                                        BlockDimension.BlockIsSecondDim, StorageKind.Device, 0)
    h_mc = mgr.register_layer_separate(_cfg(), [mc + l * PER_LAYER for l in range(NL)], [PER_LAYER] * NL,
                                       BlockDimension.BlockIsSecondDim, StorageKind.Device, 0)
    rng = np.random.default_rng(5)
    n = 24
    sid, did = rng.permutation(NB)[:n].tolist(), rng.permutation(NB)[:n].tolist()
    mgr.execute_transfer(h_src, sid, h_mc, did, TransferOptions(multicast=1)).wait()
    want = _expect(src, sid, did)
    for d, p in enumerate(pools):
        torch.cuda.synchronize(d)
        assert np.array_equal(p.cpu().numpy(), want), f"pool on cuda:{d} differs from the oracle"
    # a cast cannot ride on the multicast path: loud error, nothing launched
    with pytest.raises(KvbmError):
        mgr.execute_transfer(h_src, sid, h_mc, did, TransferOptions(multicast=1, cast_mode=K.CastMode.FP8E4M3_TO_BF16)).wait()
    mgr.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_kernel_abi_multicast_with_per_receiver_flags_and_layer_range(group, mode):
    g, pools, mc = group
    _reset(pools)
    ndev = len(pools)
    mgr = TransferManager(device=0, worker_id=8)
    for d in range(1, ndev):
        mgr.enable_peer_access(d)
    src = _src_pool(12)
    sbase = torch.tensor([b.data_ptr() for b in src], dtype=torch.int64, device="cuda:0")
    dbase = torch.tensor([mc + l * PER_LAYER for l in range(NL)], dtype=torch.int64, device="cuda:0")
    s_desc = K.PagedLayout(sbase.data_ptr(), REGION, REGION * NB, REGION, NL, NO, NB)
    d_desc = K.PagedLayout(dbase.data_ptr(), REGION, REGION * NB, REGION, NL, NO, NB)
    rng = np.random.default_rng(6)
    n = 31
    sid, did = rng.permutation(NB)[:n], rng.permutation(NB)[:n]
    s_t = torch.tensor(sid, dtype=torch.int32, device="cuda:0")
    d_t = torch.tensor(did, dtype=torch.int32, device="cuda:0")
    flags = [torch.zeros(1 + NL, dtype=torch.int32, device=f"cuda:{d}") for d in range(ndev)]
    ws = torch.zeros(NL + 4, dtype=torch.int32, device="cuda:0")
    dsts = [K.PagedDst(d_desc, s_t.data_ptr(), d_t.data_ptr(), f.data_ptr(), f[1:].data_ptr()) for f in flags]
    opts = K.PagedCopyOpts(epoch=3, sync_workspace=ws.data_ptr(), multicast=mode)
    K.check(K.paged_copy(s_desc, dsts, n, 1, 3, 0, opts, int(torch.cuda.current_stream().cuda_stream)), "paged_copy")
    torch.cuda.synchronize(0)
    want = _expect(src, sid, did, layers=range(1, 3))
    for d, p in enumerate(pools):
        torch.cuda.synchronize(d)
        assert flags[d].tolist() == [3, 0, 3, 3, 0], f"flags on cuda:{d}"
        assert np.array_equal(p.cpu().numpy(), want), f"pool on cuda:{d} differs from the oracle (multicast={mode})"
    mgr.close()


def test_staged_broadcast_serves_receivers_with_their_own_block_tables(group):
    """The multicast write needs identical offsets in every bound pool; real decode workers allocate their own blocks.
    Staged form (disagg.staged_send / staged_receive): multicast once into blocks 0..n-1 of the group-bound pools, every
    receiver scatters locally into ITS table, gated layer by layer on the flags the root's kernel sets in its memory; the
    root reuses the staging pool only after every receiver reported the previous epoch.  Byte-exact vs the oracle."""
    from dynamo_b200.disagg import staged_receive, staged_send
    g, pools, mc = group
    nd = len(pools)
    _reset(pools)
    src = _src_pool(4242)
    root = TransferManager(device=0, worker_id=70)
    receivers = list(range(1, nd))
    for d in receivers:
        root.enable_peer_access(d)
    h_src = root.register_layer_separate(_cfg(), [b.data_ptr() for b in src], [b.numel() for b in src], BlockDimension.BlockIsSecondDim,
                                         StorageKind.Device, 0)
    h_mc = root.register_layer_separate(_cfg(), [mc + l * PER_LAYER for l in range(NL)], [PER_LAYER] * NL,
                                        BlockDimension.BlockIsSecondDim, StorageKind.Device, 0)
    free = torch.zeros(nd, dtype=torch.int32, device="cuda:0")
    side = torch.cuda.Stream(device="cuda:0")
    mgrs, h_stage, h_dst, dsts, ready = {}, {}, {}, {}, {}
    for d in receivers:
        with torch.cuda.device(d):
            m = TransferManager(device=d, worker_id=70 + d)
            m.enable_peer_access(0)
            mgrs[d] = m
            h_stage[d] = m.register_layer_separate(_cfg(), [pools[d].data_ptr() + l * PER_LAYER for l in range(NL)], [PER_LAYER] * NL,
                                                   BlockDimension.BlockIsSecondDim, StorageKind.Device, d)
            dsts[d] = [torch.zeros(PER_LAYER, dtype=torch.uint8, device=f"cuda:{d}") for _ in range(NL)]
            h_dst[d] = m.register_layer_separate(_cfg(), [b.data_ptr() for b in dsts[d]], [PER_LAYER] * NL,
                                                 BlockDimension.BlockIsSecondDim, StorageKind.Device, d)
            ready[d] = torch.zeros(NL, dtype=torch.int32, device=f"cuda:{d}")
    for d in range(nd):
        torch.cuda.synchronize(d)
    n = 20
    want = {d: _twin() for d in receivers}
    for epoch in (1, 2, 3):
        rng = np.random.default_rng(100 + epoch)
        sid = rng.permutation(NB)[:n]
        dids = {d: rng.permutation(NB)[:n] for d in receivers}            # every receiver scatters into its OWN blocks
        notes = []
        for d in receivers:                                                 # receivers first: they wait (on their GPU) for the flags
            with torch.cuda.device(d):
                notes.append(staged_receive(mgrs[d], h_stage[d], h_dst[d], list(dids[d]), ready[d].data_ptr(), epoch,
                                            free_flag=free[d:].data_ptr(), max_ctas=8))
        staged_send(root, h_src, list(sid), h_mc, [ready[d].data_ptr() for d in receivers], epoch, stream_ptr(side),
                    receiver_free_flags=[free[d:].data_ptr() for d in receivers])
        try:
            for nt in notes:
                nt.wait(60.0)
        except KvbmError as e:      # say what the flags looked like: which half of the hand-shake did not happen?
            import time
            time.sleep(0.2)
            state = {d: ready[d].tolist() for d in receivers}
            raise AssertionError(f"epoch {epoch}: {e}; ready flags {state}; free {free.tolist()}; root stream idle: {side.query()}")
        side.synchronize()
        for d in receivers:
            O.execute_memcpy_transfer(_twin(src), want[d], sid, dids[d])
            got = np.concatenate([b.cpu().numpy() for b in dsts[d]])
            assert np.array_equal(got, np.concatenate(want[d].buffers)), f"receiver {d} epoch {epoch}"
            assert ready[d].tolist() == [epoch] * NL
        assert free[1:].tolist() == [epoch] * len(receivers)
    for d in range(nd):
        torch.cuda.synchronize(d)
    for m in mgrs.values():
        m.close()
    root.close()
    _reset(pools)


def test_bind_addr_binds_memory_the_engine_allocated_itself():
    """kvbm_mc_group_bind_addr = cuMulticastBindAddr: the engine's OWN cuMemCreate-backed pool (allocated here through
    cuda-python, standing in for an engine allocator) becomes a member of the group -- no pool from the group's allocator."""
    ndev = torch.cuda.device_count()
    if ndev < 2 or not all(multicast_supported(d) for d in range(ndev)):
        pytest.skip("needs >= 2 GPUs with NVLink multicast support")
    drv = pytest.importorskip("cuda.bindings.driver")
    nl, nb = 2, 32
    per_layer = NO * nb * REGION
    total = nl * per_layer                       # 4 MiB
    for d in range(ndev):
        torch.zeros(1, device=f"cuda:{d}")
    torch.cuda.set_device(0)
    g = MulticastGroup.create(ndev, total)
    size = g.size
    pools, keep = [], []

    def ck(res):
        assert int(res[0]) == 0, res[0]
        return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)

    for d in range(ndev):
        g.add_device(d)
    for d in range(ndev):
        prop = drv.CUmemAllocationProp()
        prop.type = drv.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
        prop.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        prop.location.id = d
        gran = ck(drv.cuMemGetAllocationGranularity(prop, drv.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_RECOMMENDED))
        assert size % int(gran) == 0
        handle = ck(drv.cuMemCreate(size, prop, 0))
        va = ck(drv.cuMemAddressReserve(size, max(int(gran), 512 << 20), 0, 0))
        ck(drv.cuMemMap(va, size, 0, handle, 0))
        descs = []
        for p in range(ndev):
            ad = drv.CUmemAccessDesc()
            ad.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
            ad.location.id = p
            ad.flags = drv.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
            descs.append(ad)
        ck(drv.cuMemSetAccess(va, size, descs, len(descs)))
        g.bind_addr(d, int(va), size)            # <- the call under test
        t = _view(int(va), size, d)
        t.zero_()
        pools.append(t)
        keep.append((handle, va))
    for d in range(ndev):
        torch.cuda.synchronize(d)
    mc = g.map(0)
    cfg = LayoutConfig(nb, nl, NO, PAGE, INNER, dtype_width_bytes=DT)
    mgr = TransferManager(device=0, worker_id=91)
    src = [torch.randint(0, 256, (per_layer,), dtype=torch.uint8, device="cuda:0") for _ in range(nl)]
    h_src = mgr.register_layer_separate(cfg, [b.data_ptr() for b in src], [per_layer] * nl, BlockDimension.BlockIsSecondDim, StorageKind.Device, 0)
    h_mc = mgr.register_layer_separate(cfg, [mc + l * per_layer for l in range(nl)], [per_layer] * nl, BlockDimension.BlockIsSecondDim,
                                       StorageKind.Device, 0)
    sid, did = [3, 9, 30, 0, 17], [1, 2, 8, 31, 5]
    mgr.execute_transfer(h_src, sid, h_mc, did, TransferOptions(multicast=1)).wait(30.0)
    for d in range(ndev):
        torch.cuda.synchronize(d)
    for d in range(ndev):
        for l in range(nl):
            got = pools[d][l * per_layer:(l + 1) * per_layer].view(NO, nb, REGION)[:, did].cpu()
            assert torch.equal(got, src[l].view(NO, nb, REGION)[:, sid].cpu()), f"engine-owned pool on cuda:{d}, layer {l}"
    mgr.close()
    del pools
    g.detach()     # the driver reclaims the object at exit; unbinding takes seconds per member

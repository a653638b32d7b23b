"""GPU tests of the host API (TransferManager over libkvbm_physical.so) -- the reference's transfer
integration tests with Device storage: lib/kvbm-physical/src/transfer/tests/local_transfers.rs
(test_p2p :108-171, test_roundtrip :173-241, test_large_block_counts :286-320, guard blocks :460-540,
layer composition :922-...) checked with BLAKE3 block checksums against the CPU oracle."""
import itertools

import numpy as np
import pytest
import torch

from dynamo_b200 import kernels as K
from dynamo_b200.physical import (BlockDimension, ErrorCode, KvbmError, LayoutConfig, StorageKind, TransferManager,
                                  TransferOptions)
from oracle import oracle as O

pytestmark = pytest.mark.gpu
KINDS = ["FC", "LWf", "LWs"]


@pytest.fixture()
def mgr():
    m = TransferManager(device=0, worker_id=1)
    yield m
    torch.cuda.synchronize()
    m.close()


class Pool:
    """A registered layout + its oracle twin.  storage: Device (torch cuda), Pinned (torch pinned), System (numpy)."""

    def __init__(self, mgr, kind, nb, storage, nl=2, no=2, page=16, inner=128, dt=2, device=0, allow_fp8=False):
        self.mgr, self.storage = mgr, storage
        bd = O.BLOCK_IS_SECOND_DIM if kind == "LWs" else O.BLOCK_IS_FIRST_DIM
        self.twin = O.Layout(O.FC if kind == "FC" else O.LW, nb, nl, no, page, inner, dt, block_dim=bd, allow_fp8=allow_fp8)
        self.cfg = LayoutConfig(nb, nl, no, page, inner, dtype_width_bytes=dt, allow_fp8=allow_fp8)
        if storage == StorageKind.Device:
            self.mem = [torch.zeros(b.size, dtype=torch.uint8, device=f"cuda:{device}") for b in self.twin.buffers]
        elif storage == StorageKind.Pinned:
            self.mem = [torch.zeros(b.size, dtype=torch.uint8).pin_memory() for b in self.twin.buffers]
        else:
            self.mem = [torch.from_numpy(b) for b in self.twin.buffers]
        ptrs, sizes = [t.data_ptr() for t in self.mem], [t.numel() for t in self.mem]
        if kind == "FC":
            self.h = mgr.register_fully_contiguous(self.cfg, ptrs[0], sizes[0], storage, device)
        else:
            self.h = mgr.register_layer_separate(self.cfg, ptrs, sizes, BlockDimension(bd), storage, device)

    def upload(self):      # twin -> registered memory
        for t, b in zip(self.mem, self.twin.buffers):
            t.copy_(torch.from_numpy(b))
        torch.cuda.synchronize()

    def download(self):    # registered memory -> twin
        torch.cuda.synchronize()
        for t, b in zip(self.mem, self.twin.buffers):
            b[:] = t.cpu().numpy()

    def bytes(self):
        torch.cuda.synchronize()
        return [t.cpu().numpy().copy() for t in self.mem]


@pytest.mark.parametrize("sk,dk", list(itertools.product(KINDS, repeat=2)))
@pytest.mark.parametrize("mode", [None, range(0, 1), range(1, 2)], ids=["full", "layer0", "layer1"])
def test_device_to_device_matrix_with_guards(mgr, sk, dk, mode):
    src, dst = Pool(mgr, sk, 6, StorageKind.Device), Pool(mgr, dk, 6, StorageKind.Device)
    ref = O.Layout(dst.twin.kind, 6, 2, 2, 16, 128, 2, block_dim=dst.twin.c.block_dim)
    src.twin.fill_blocks([0, 1], -1)
    for t in (dst.twin, ref):
        t.fill_blocks([2, 5], 0xFF)
    src.upload()
    dst.upload()
    note = mgr.execute_transfer(src.h, [0, 1], dst.h, [3, 4], TransferOptions(layer_range=mode))
    note.wait()
    assert note.is_complete()
    O.execute_memcpy_transfer(src.twin, ref, [0, 1], [3, 4], mode)
    for got, want in zip(dst.bytes(), ref.buffers):
        assert np.array_equal(got, want)
    dst.download()
    want = src.twin.block_checksums([0, 1], mode)
    got = dst.twin.block_checksums([3, 4], mode)
    assert [got[3], got[4]] == [want[0], want[1]]


@pytest.mark.parametrize("sk,ik,dk", list(itertools.product(["FC", "LWs"], repeat=3)))
def test_roundtrip_pinned_device_pinned(mgr, sk, ik, dk):
    # local_transfers.rs:173-241: Pinned[0,1] -> Device[0,1] -> Pinned[2,3]   (CudaAsyncH2D then CudaAsyncD2H)
    src, dev, dst = Pool(mgr, sk, 4, StorageKind.Pinned), Pool(mgr, ik, 4, StorageKind.Device), Pool(mgr, dk, 4, StorageKind.Pinned)
    src.twin.fill_blocks([0, 1], -1)
    src.upload()
    want = src.twin.block_checksums([0, 1])
    mgr.execute_transfer(src.h, [0, 1], dev.h, [0, 1]).wait()
    mgr.execute_transfer(dev.h, [0, 1], dst.h, [2, 3]).wait()
    dst.download()
    got = dst.twin.block_checksums([2, 3])
    assert [got[2], got[3]] == [want[0], want[1]]


def test_system_device_is_rejected(mgr):
    sysl, dev = Pool(mgr, "FC", 4, StorageKind.System), Pool(mgr, "FC", 4, StorageKind.Device)
    with pytest.raises(KvbmError) as e:   # strategy.rs:161: "System to Device transfers are not supported"
        mgr.execute_transfer(sysl.h, [0], dev.h, [1])
    assert e.value.code == ErrorCode.UNSUPPORTED


@pytest.mark.parametrize("block_count", [1024, 4096, 16384])
def test_large_block_counts(mgr, block_count):
    # local_transfers.rs:286-320 (Pinned -> Device, identity block lists, all blocks), small regions
    src = Pool(mgr, "FC", block_count, StorageKind.Pinned, nl=2, no=2, page=4, inner=32, dt=2)
    dev = Pool(mgr, "LWs", block_count, StorageKind.Device, nl=2, no=2, page=4, inner=32, dt=2)
    rng = np.random.default_rng(block_count)
    src.twin.buffers[0][:] = rng.integers(0, 256, src.twin.buffers[0].size, dtype=np.uint8)
    src.upload()
    ids = list(range(block_count))
    mgr.execute_transfer(src.h, ids, dev.h, ids).wait()
    ref = O.Layout(O.LW, block_count, 2, 2, 4, 32, 2, block_dim=O.BLOCK_IS_SECOND_DIM)
    O.execute_memcpy_transfer(src.twin, ref, ids, ids)
    for got, want in zip(dev.bytes(), ref.buffers):
        assert np.array_equal(got, want)


def test_many_transfers_in_flight_recycle_slots(mgr):
    src, dst = Pool(mgr, "LWs", 64, StorageKind.Device), Pool(mgr, "LWs", 64, StorageKind.Device)
    rng = np.random.default_rng(9)
    for b in src.twin.buffers:
        b[:] = rng.integers(0, 256, b.size, dtype=np.uint8)
    src.upload()
    notes = []
    for i in range(200):   # > 64 slots: the ring must recycle without corrupting earlier transfers
        notes.append(mgr.execute_transfer(src.h, [i % 64], dst.h, [(i * 7) % 64]))
    for n in notes:
        n.wait()
    ref = O.Layout(O.LW, 64, 2, 2, 16, 128, 2, block_dim=O.BLOCK_IS_SECOND_DIM)
    for i in range(200):
        O.execute_memcpy_transfer(src.twin, ref, [i % 64], [(i * 7) % 64])
    for got, want in zip(dst.bytes(), ref.buffers):
        assert np.array_equal(got, want)


def test_caller_stream_returns_completed_notification(mgr):
    src, dst = Pool(mgr, "LWs", 8, StorageKind.Device), Pool(mgr, "LWs", 8, StorageKind.Device)
    src.twin.fill_blocks([0, 1, 2], -1)
    src.upload()
    s = torch.cuda.Stream()
    note = mgr.execute_transfer(src.h, [0, 1, 2], dst.h, [5, 6, 7], TransferOptions(cuda_stream=int(s.cuda_stream)))
    assert note.token == 0 and note.is_complete()      # cuda.rs:139-141: caller manages sync
    s.synchronize()
    dst.download()
    assert dst.twin.block_checksums([5, 6, 7]) == {5: src.twin.block_checksum(0), 6: src.twin.block_checksum(1), 7: src.twin.block_checksum(2)}


@pytest.mark.parametrize("replicate", [False, True])
def test_fanout_and_broadcast(mgr, replicate):
    nb, n, nd = 32, 10, 4
    src = Pool(mgr, "LWs", nb, StorageKind.Device, nl=3, inner=256)
    dsts = [Pool(mgr, "LWs", nb, StorageKind.Device, nl=3, inner=256) for _ in range(nd)]
    rng = np.random.default_rng(17)
    for b in src.twin.buffers:
        b[:] = rng.integers(0, 256, b.size, dtype=np.uint8)
    src.upload()
    sids = [list(rng.permutation(nb)[:n]) for _ in range(nd)]
    dids = [list(rng.permutation(nb)[:n]) for _ in range(nd)]
    if replicate:
        mgr.broadcast(src.h, [d.h for d in dsts], sids[0], dids[0]).wait()
        sids, dids = [sids[0]] * nd, [dids[0]] * nd
    else:
        mgr.execute_fanout(src.h, [d.h for d in dsts], sids, dids).wait()
    for d, sid, did in zip(dsts, sids, dids):
        ref = O.Layout(O.LW, nb, 3, 2, 16, 256, 2, block_dim=O.BLOCK_IS_SECOND_DIM)
        O.execute_memcpy_transfer(src.twin, ref, sid, did)
        for got, want in zip(d.bytes(), ref.buffers):
            assert np.array_equal(got, want)


def test_cast_through_manager(mgr):
    nb, n = 16, 6
    src = Pool(mgr, "LWs", nb, StorageKind.Device, nl=2, inner=512, dt=1, allow_fp8=True)
    dst = Pool(mgr, "LWs", nb, StorageKind.Device, nl=2, inner=512, dt=2)
    rng = np.random.default_rng(23)
    for b in src.twin.buffers:
        b[:] = rng.integers(0, 256, b.size, dtype=np.uint8)
    src.upload()
    sid, did = list(rng.permutation(nb)[:n]), list(rng.permutation(nb)[:n])
    mgr.execute_transfer(src.h, sid, dst.h, did, TransferOptions(cast_mode=K.CastMode.FP8E4M3_TO_BF16)).wait()
    ref = O.Layout(O.LW, nb, 2, 2, 16, 512, 2, block_dim=O.BLOCK_IS_SECOND_DIM)
    O.execute_memcpy_transfer(src.twin, ref, sid, did, cast_mode=1)
    for got, want in zip(dst.bytes(), ref.buffers):
        assert np.array_equal(got, want)
    with pytest.raises(KvbmError) as e:      # widths must be 1 -> 2
        mgr.execute_transfer(dst.h, sid, dst.h, did, TransferOptions(cast_mode=K.CastMode.FP8E4M3_TO_BF16))
    assert e.value.code in (ErrorCode.INCOMPATIBLE, ErrorCode.OVERLAP)


def test_layer_streaming_with_ready_and_done_flags(mgr):
    # physical.rs:277-346 pattern, but ONE launch: the producer releases layers, the kernel publishes layers
    nb, n, nl = 32, 16, 8
    src = Pool(mgr, "LWs", nb, StorageKind.Device, nl=nl, inner=256)
    dst = Pool(mgr, "LWs", nb, StorageKind.Device, nl=nl, inner=256)
    rng = np.random.default_rng(29)
    for b in src.twin.buffers:
        b[:] = rng.integers(0, 256, b.size, dtype=np.uint8)
    src.upload()
    ready = torch.zeros(nl, dtype=torch.int32, device="cuda")
    done = torch.zeros(nl, dtype=torch.int32, device="cuda")
    sid, did = list(range(n)), list(range(n, 2 * n))
    ctl = torch.cuda.Stream()     # control-plane stream of the "engine": releases layers, waits for their arrival
    note = mgr.execute_transfer(src.h, sid, dst.h, did, TransferOptions(layer_ready_flags=ready.data_ptr(),
                                                                        layer_done_flags=done.data_ptr(), epoch=5, max_ctas=16))
    try:
        assert not note.is_complete()          # gated on layer 0's ready flag
        for l in range(nl):
            K.check(K.set_flags(ready.data_ptr(), l, 1, 5, int(ctl.cuda_stream)))
            K.check(K.wait_flag(done[l:].data_ptr(), 5, int(ctl.cuda_stream)))   # layer l seen before l+1 is released
        note.wait(20.0)
    finally:   # whatever happened, let every spinning kernel finish so teardown cannot hang
        rescue = torch.cuda.Stream()
        K.set_flags(ready.data_ptr(), 0, nl, 5, int(rescue.cuda_stream))
        if not note.is_complete():
            K.set_flags(done.data_ptr(), 0, nl, 5, int(rescue.cuda_stream))
    torch.cuda.synchronize()
    assert done.tolist() == [5] * nl
    ref = O.Layout(O.LW, nb, nl, 2, 16, 256, 2, block_dim=O.BLOCK_IS_SECOND_DIM)
    O.execute_memcpy_transfer(src.twin, ref, sid, did)
    for got, want in zip(dst.bytes(), ref.buffers):
        assert np.array_equal(got, want)


def test_done_flag_option_delivers_the_epoch_to_destination_memory(mgr):
    """TransferOptions::nixl_write_notification (options.rs:36-43) restated for peer stores: a word in the destination's
    memory receives the value once every byte has landed."""
    src, dst = Pool(mgr, "LWs", 8, StorageKind.Device), Pool(mgr, "LWs", 8, StorageKind.Device)
    src.twin.fill_blocks([0, 1, 2], -1)
    src.upload()
    flag = torch.zeros(4, dtype=torch.int32, device="cuda")
    note = mgr.execute_transfer(src.h, [0, 1, 2], dst.h, [5, 6, 7], TransferOptions(done_flag=flag[1:].data_ptr(), epoch=77))
    note.wait(20.0)
    torch.cuda.synchronize()
    assert flag.tolist() == [0, 77, 0, 0]
    ref = O.Layout(dst.twin.kind, 8, 2, 2, 16, 128, 2, block_dim=dst.twin.c.block_dim)
    O.execute_memcpy_transfer(src.twin, ref, [0, 1, 2], [5, 6, 7])
    for got, want in zip(dst.bytes(), ref.buffers):
        assert np.array_equal(got, want)


@pytest.mark.parametrize("bounce_blocks", [1, 2, 5, 64], ids=lambda b: f"bounce{b}")
def test_two_hop_through_a_pinned_bounce_buffer(mgr, bounce_blocks):
    """TwoHop{CudaAsyncD2H, Pinned, CudaAsyncH2D} (strategy.rs:222-233) executed like executor/mod.rs:357-416,477-571: the
    bounce blocks are split into two groups, chunks alternate between them, one bounce block = one block at a time.  The two
    device pools are registered as living on different GPUs (device ids 0 / 1) so that disallowing GPU RDMA selects the plan;
    the bytes are checked against the oracle and the bounce pool must hold nothing but staged copies of source blocks."""
    nb, n = 40, 23
    src, dst = Pool(mgr, "LWs", nb, StorageKind.Device, nl=3), Pool(mgr, "FC", nb, StorageKind.Device, nl=3)
    far = mgr.register_fully_contiguous(dst.cfg, dst.mem[0].data_ptr(), dst.mem[0].numel(), StorageKind.Device, 1)
    bounce = Pool(mgr, "FC", max(bounce_blocks, 2), StorageKind.Pinned, nl=3)
    rng = np.random.default_rng(bounce_blocks)
    for b in src.twin.buffers:
        b[:] = rng.integers(0, 256, b.size, dtype=np.uint8)
    src.upload()
    sid, did = rng.permutation(nb)[:n], rng.permutation(nb)[:n]
    bids = list(rng.permutation(max(bounce_blocks, 2))[:bounce_blocks])
    mgr.set_capabilities(allow_gpu_rdma=False)
    try:
        with pytest.raises(KvbmError) as e:                                   # executor/mod.rs:514-519
            mgr.execute_transfer(src.h, list(sid), far, list(did))
        assert "Two-hop transfers require a bounce buffer." in e.value.msg
        dev_bounce = Pool(mgr, "FC", 4, StorageKind.Device, nl=3)
        with pytest.raises(KvbmError) as e:                                   # :521-527
            mgr.execute_transfer(src.h, list(sid), far, list(did), TransferOptions(bounce_buffer=(dev_bounce.h, [0, 1])))
        assert "Bounce buffer layout does not match bounce location." in e.value.msg
        with pytest.raises(KvbmError):
            mgr.execute_transfer(src.h, list(sid), far, list(did), TransferOptions(bounce_buffer=(bounce.h, [0, 0])))   # duplicate ids
        before = K.launch_count()
        note = mgr.execute_transfer(src.h, list(sid), far, list(did), TransferOptions(bounce_buffer=(bounce.h, bids)))
        note.wait(30.0)
        nbq = min(bounce_blocks, n)
        lens = [nbq // 2, nbq - nbq // 2] if nbq >= 2 else [nbq]
        pos = chunks = 0
        while pos < n:
            pos += min(lens[chunks % len(lens)], n - pos)
            chunks += 1
        assert K.launch_count() - before == 2 * chunks                          # two launches (hops) per chunk
        # same device id on both sides stays a direct D2D even without GPU RDMA
        mgr.execute_transfer(src.h, list(sid[:3]), dst.h, [int(x) for x in np.setdiff1d(np.arange(nb), did)[:3]]).wait(30.0)
    finally:
        mgr.set_capabilities(allow_gpu_rdma=True)
    dst.download()
    want = src.twin.block_checksums(sid)
    for s, d in zip(sid, did):
        assert dst.twin.block_checksum(int(d)) == want[int(s)]
    bounce.download()
    staged = {bounce.twin.block_checksum(int(b)) for b in bids[:min(len(bids), n)]}   # only min(bounce, n) blocks are ever used
    assert staged <= set(want.values())

"""Property tests (hypothesis) of the host half against the oracle: random layouts, block tables and layer ranges.
CPU only -- the Memcpy strategy moves the bytes, the same validation / address code fronts the CUDA strategies."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from dynamo_b200 import physical as P
from dynamo_b200.physical import (BlockDimension, ErrorCode, KvbmError, LayoutConfig, StorageKind, TransferManager,
                                  TransferOptions)
from oracle import oracle as O

KINDS = ["FC", "LWf", "LWs"]
_SETTINGS = dict(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


@pytest.fixture(scope="module")
def mgr():
    m = TransferManager(device=-1, worker_id=9)
    yield m
    m.close()


def _twin(kind, nb, nl, no, page, inner, dt, fill=0):
    k, bd = {"FC": (O.FC, O.BLOCK_IS_FIRST_DIM), "LWf": (O.LW, O.BLOCK_IS_FIRST_DIM), "LWs": (O.LW, O.BLOCK_IS_SECOND_DIM)}[kind]
    return O.Layout(k, nb, nl, no, page, inner, dt, block_dim=bd, fill=fill)


def _register(mgr, kind, twin, cfg):
    ptrs, sizes = [b.ctypes.data for b in twin.buffers], [b.size for b in twin.buffers]
    if kind == "FC":
        return mgr.register_fully_contiguous(cfg, ptrs[0], sizes[0], StorageKind.System)
    bd = BlockDimension.BlockIsSecondDim if kind == "LWs" else BlockDimension.BlockIsFirstDim
    return mgr.register_layer_separate(cfg, ptrs, sizes, bd, StorageKind.System)


geometry = st.tuples(st.integers(1, 9), st.integers(1, 4), st.integers(1, 2), st.sampled_from([1, 4, 16]), st.sampled_from([8, 24, 128]),
                     st.sampled_from([2, 4, 8]))


@settings(**_SETTINGS)
@given(geom=geometry, sk=st.sampled_from(KINDS), dk=st.sampled_from(KINDS), data=st.data())
def test_host_transfers_equal_the_oracle_for_random_layouts_tables_and_layer_ranges(mgr, geom, sk, dk, data):
    nb, nl, no, page, inner, dt = geom
    src_t, dst_t, ref_t = _twin(sk, nb, nl, no, page, inner, dt), _twin(dk, nb, nl, no, page, inner, dt, 0x5C), _twin(dk, nb, nl, no, page, inner, dt, 0x5C)
    rng = np.random.default_rng(data.draw(st.integers(0, 2**32 - 1)))
    for b in src_t.buffers:
        b[:] = rng.integers(0, 256, b.size, dtype=np.uint8)
    cfg = LayoutConfig(nb, nl, no, page, inner, dtype_width_bytes=dt)
    hs, hd = _register(mgr, sk, src_t, cfg), _register(mgr, dk, dst_t, cfg)
    try:
        n = data.draw(st.integers(0, nb))
        sids = [int(x) for x in rng.integers(0, nb, n)]                   # sources may repeat
        dids = [int(x) for x in rng.permutation(nb)[:n]]                  # destinations are unique
        lb = data.draw(st.integers(0, nl))
        le = data.draw(st.integers(lb, nl))
        lr = data.draw(st.sampled_from([None, range(lb, le)]))
        mgr.execute_transfer(hs, sids, hd, dids, TransferOptions(layer_range=lr) if lr is not None else None).wait()
        O.execute_memcpy_transfer(src_t, ref_t, sids, dids, lr)
        for got, want in zip(dst_t.buffers, ref_t.buffers):
            assert np.array_equal(got, want)
        for b, l, o in [(0, 0, 0), (nb - 1, nl - 1, no - 1)]:            # address arithmetic of both libraries
            assert mgr.memory_region(hs, b, l, o) == src_t.memory_region(b, l, o)
            assert mgr.memory_region(hd, b, l, o) == dst_t.memory_region(b, l, o)
    finally:
        mgr.unregister(hs)
        mgr.unregister(hd)


_CODES = {0: None, O.ERR_LENGTH_MISMATCH: ErrorCode.LENGTH_MISMATCH, O.ERR_DUP_DST: ErrorCode.DUPLICATE_DST,
          O.ERR_OVERLAP: ErrorCode.OVERLAP, O.ERR_RANGE: ErrorCode.RANGE}


@settings(**_SETTINGS)
@given(nb_s=st.integers(1, 12), nb_d=st.integers(1, 12), same=st.booleans(),
       sids=st.lists(st.integers(0, 14), max_size=10), dids=st.lists(st.integers(0, 14), max_size=10))
def test_validation_agrees_with_the_oracle_on_arbitrary_id_lists(nb_s, nb_d, same, sids, dids):
    """validation.rs: mismatched lengths, duplicate destinations, overlap on the same layout, out-of-range ids -- the first
    failing check (in the reference's order) decides the error."""
    if same:
        nb_d = nb_s
    src = O.Layout(O.FC, nb_s, 1, 1, 1, 8, 2)
    dst = src if same else O.Layout(O.FC, nb_d, 1, 1, 1, 8, 2)
    want = _CODES[O.validate_block_transfer(sids, dids, src, dst)]
    try:
        P.validate_block_transfer(sids, dids, nb_s, nb_d, same)
        got = None
    except KvbmError as e:
        got = ErrorCode(e.code)
    assert got == want


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(geom=geometry, kind=st.sampled_from(KINDS), heads=st.sampled_from([None, 1, 2, 4]))
def test_serialized_layout_round_trip_keeps_the_geometry_for_random_configs(geom, kind, heads):
    """manager/metadata.rs:87-159 + layout/serialize.rs: export -> import in a second manager describes the same regions."""
    nb, nl, no, page, inner, dt = geom
    twin = _twin(kind, nb, nl, no, page, inner, dt)
    cfg = LayoutConfig(nb, nl, no, page, inner, dtype_width_bytes=dt, num_heads=heads)
    a, b = TransferManager(device=-1, worker_id=21), TransferManager(device=-1, worker_id=22)
    try:
        h = _register(a, kind, twin, cfg)
        handles = b.import_serialized_layout(a.export_serialized_layout())
        assert len(handles) == 1
        assert b.is_fully_contiguous(handles[0]) == (kind == "FC")
        for blk, l, o in [(0, 0, 0), (nb - 1, nl - 1, no - 1), (nb // 2, nl // 2, 0)]:
            assert b.memory_region(handles[0], blk, l, o) == a.memory_region(h, blk, l, o) == twin.memory_region(blk, l, o)
        again = a.import_descriptor_json(a.layout_descriptor_json(h))
        assert a.memory_region(again, nb - 1, nl - 1, no - 1) == twin.memory_region(nb - 1, nl - 1, no - 1)
    finally:
        a.close()
        b.close()

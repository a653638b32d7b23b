"""Pins the oracle's memcpy transfer / fill / checksum / validation against the reference's tests.

Sources (relative to /root/reference/lib/kvbm-physical/src):
  transfer/tests/mod.rs:128-138         standard_config: nl=2 no=2 page=16 inner=128 dtype=2
  transfer/tests/local_transfers.rs:108-171   test_p2p (src [0,1] -> dst [2,3], Sequential fill, BLAKE3 by position)
  transfer/tests/local_transfers.rs:460-540   guard blocks [2,5] = 0xFF stay unchanged, bounce 3-hop
  transfer/tests/local_transfers.rs:922-...   layer composition == full block
  transfer/fill.rs:236-293              fill KATs
  transfer/checksum.rs:163-195          checksum of constant pattern == blake3(vec![42; region])
  transfer/validation.rs                duplicate dst / overlap / range / length
"""
import itertools

import blake3
import numpy as np
import pytest

from oracle import oracle as O

KINDS = [("FC", dict(kind=O.FC)), ("LWf", dict(kind=O.LW, block_dim=O.BLOCK_IS_FIRST_DIM)),
         ("LWs", dict(kind=O.LW, block_dim=O.BLOCK_IS_SECOND_DIM))]


def std(num_blocks, kw, fill=0):
    return O.Layout(num_blocks=num_blocks, num_layers=2, outer_dim=2, page_size=16, inner_dim=128,
                    dtype_width_bytes=2, fill=fill, **kw)


def test_fill_sequential_kat():
    # fill.rs:254-279: block 0 layer 0 starts 0,1; block 1 layer 1 starts 2,3
    L = std(2, dict(kind=O.FC))
    L.fill_blocks([0, 1], -1)
    r = L.region_bytes(0, 0, 0)
    assert r[0] == 0 and r[1] == 1
    r = L.region_bytes(1, 1, 0)
    assert r[0] == 2 and r[1] == 3
    assert r[300] == (1 + 1 + 300) % 256


def test_fill_layers_kat():
    # fill.rs:281-293
    L = std(2, dict(kind=O.FC))
    L.fill_layers([0], 0, 1, 0)
    L.fill_layers([0], 1, 2, 1)
    L.fill_layers([1], 0, 1, 100)
    L.fill_layers([1], 1, 2, 101)
    got = [L.region_bytes(b, l, 0)[0] for b, l in [(0, 0), (0, 1), (1, 0), (1, 1)]]
    assert got == [0, 1, 100, 101]


def test_checksum_constant_pattern():
    # checksum.rs:163-195
    L = std(2, dict(kind=O.FC))
    L.fill_blocks([0, 1], 42)
    cs = L.block_checksums([0, 1])
    assert cs[0] == cs[1]
    region = L.region_bytes(0, 0, 0)
    assert (region == 42).all()
    assert blake3.blake3(region.tobytes()).hexdigest() == blake3.blake3(bytes([42]) * L.region_size).hexdigest()
    # whole-block digest is BLAKE3 over the 4 regions in layer-major, outer-minor order
    h = blake3.blake3()
    for l in range(2):
        for o in range(2):
            h.update(L.region_bytes(0, l, o).tobytes())
    assert cs[0] == h.hexdigest()


@pytest.mark.parametrize("src_kw,dst_kw", list(itertools.product([k[1] for k in KINDS], repeat=2)),
                         ids=[f"{a[0]}-{b[0]}" for a, b in itertools.product(KINDS, repeat=2)])
def test_p2p_checksums_by_position(src_kw, dst_kw):
    # local_transfers.rs:108-171 (host<->host rows of the matrix)
    src, dst = std(4, src_kw), std(4, dst_kw)
    src.fill_blocks([0, 1], -1)
    want = src.block_checksums([0, 1])
    O.execute_memcpy_transfer(src, dst, [0, 1], [2, 3])
    got = dst.block_checksums([2, 3])
    assert [got[2], got[3]] == [want[0], want[1]]
    # untouched destination blocks remain zero
    assert not dst.region_bytes(0, 0, 0).any() and not dst.region_bytes(1, 1, 1).any()


@pytest.mark.parametrize("host_kw,bounce_kw,mode", list(itertools.product(
    [k[1] for k in KINDS], [k[1] for k in KINDS], [None, range(0, 1), range(1, 2)])))
def test_bounce_with_guards(host_kw, bounce_kw, mode):
    # local_transfers.rs:460-540: host[0,1] -> bounce[0,1] -> host[3,4]; guards host[2,5]=0xFF
    host, bounce = std(6, host_kw), std(6, bounce_kw)
    if mode is None:
        host.fill_blocks([0, 1], -1)
    else:
        host.fill_layers([0, 1], mode.start, mode.stop, -1)
    host.fill_blocks([2, 5], 0xFF)
    want = host.block_checksums([0, 1], mode)
    guards = host.block_checksums([2, 5])
    O.execute_memcpy_transfer(host, bounce, [0, 1], [0, 1], mode)
    O.execute_memcpy_transfer(bounce, host, [0, 1], [3, 4], mode)
    got = host.block_checksums([3, 4], mode)
    assert [got[3], got[4]] == [want[0], want[1]]
    assert host.block_checksums([2, 5]) == guards
    if mode is not None:  # the other layer of the destination must be untouched (zeros)
        other = 1 - mode.start
        assert not host.region_bytes(3, other, 0).any()


@pytest.mark.parametrize("src_kw,dst_kw", list(itertools.product([k[1] for k in KINDS], repeat=2)))
def test_layer_composition_equals_full_block(src_kw, dst_kw):
    # local_transfers.rs:922-...
    src, full, layered = std(4, src_kw), std(4, dst_kw), std(4, dst_kw)
    src.fill_blocks([0, 1], -1)
    O.execute_memcpy_transfer(src, full, [0, 1], [2, 3])
    O.execute_memcpy_transfer(src, layered, [0, 1], [2, 3], range(0, 1))
    O.execute_memcpy_transfer(src, layered, [0, 1], [2, 3], range(1, 2))
    assert full.block_checksums([2, 3]) == layered.block_checksums([2, 3])


def test_whole_block_predicate():
    # transfer/mod.rs:150-173
    fc, lw = std(2, dict(kind=O.FC)), std(2, dict(kind=O.LW))
    fc2 = std(2, dict(kind=O.FC))
    assert O.can_use_whole_block_transfer(fc, fc2, None)
    assert O.can_use_whole_block_transfer(fc, fc2, range(0, 2))
    assert not O.can_use_whole_block_transfer(fc, fc2, range(0, 1))
    assert not O.can_use_whole_block_transfer(fc, lw, None)
    assert not O.can_use_whole_block_transfer(lw, fc, None)


def test_validation_codes():
    a, b = std(4, dict(kind=O.FC)), std(4, dict(kind=O.FC))
    assert O.validate_block_transfer([0, 1], [2, 3], a, b) == O.OK
    assert O.validate_block_transfer([0, 1], [2], a, b) == O.ERR_LENGTH_MISMATCH
    assert O.validate_block_transfer([0, 1], [2, 2], a, b) == O.ERR_DUP_DST
    assert O.validate_block_transfer([0, 1], [1, 2], a, a) == O.ERR_OVERLAP   # same layout only
    assert O.validate_block_transfer([0, 1], [1, 2], a, b) == O.OK
    assert O.validate_block_transfer([0, 4], [1, 2], a, b) == O.ERR_RANGE
    assert O.validate_block_transfer([0, 1], [1, 9], a, b) == O.ERR_RANGE
    assert O.validate_block_transfer([], [], a, b) == O.OK


def test_incompatible_layouts():
    # memcpy.rs:49-63
    a = std(4, dict(kind=O.FC))
    b = O.Layout(O.FC, 4, 3, 2, 16, 128, 2)
    with pytest.raises(O.OracleError) as e:
        O.execute_memcpy_transfer(a, b, [0], [1])
    assert e.value.code == O.ERR_INCOMPATIBLE
    c = O.Layout(O.LW, 4, 2, 2, 16, 64, 2)   # different region size -> memcpy.rs:143-153
    with pytest.raises(O.OracleError) as e:
        O.execute_memcpy_transfer(a, c, [0], [1])
    assert e.value.code == O.ERR_SIZE_MISMATCH


def test_config1_cpu_handoff_multithreaded_matches_single():
    # BASELINE.json configs[0] at reduced block count: LW/BlockIsSecondDim (vLLM [2,nb,16,8,128]), random ids
    nb, n = 64, 32
    kw = dict(kind=O.LW, block_dim=O.BLOCK_IS_SECOND_DIM)
    mk = lambda: O.Layout(num_blocks=nb, num_layers=4, outer_dim=2, page_size=16, inner_dim=1024,
                          dtype_width_bytes=2, **kw)
    src, d1, d2 = mk(), mk(), mk()
    rng = np.random.default_rng(1234)
    for buf in src.buffers:
        buf[:] = rng.integers(0, 256, buf.size, dtype=np.uint8)
    sid = np.random.default_rng(0).permutation(nb)[:n]
    did = np.random.default_rng(1).permutation(nb)[:n]
    O.execute_memcpy_transfer(src, d1, sid, did)
    O.execute_memcpy_transfer(src, d2, sid, did, nthreads=4)
    want = src.block_checksums(sid)
    for s, d in zip(sid, did):
        assert d1.block_checksum(int(d)) == want[int(s)] == d2.block_checksum(int(d))
    untouched = sorted(set(range(nb)) - set(int(x) for x in did))
    assert not d1.region_bytes(untouched[0], 0, 0).any()

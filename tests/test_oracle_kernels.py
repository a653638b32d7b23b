"""Pins the oracle's restatement of the kernel semantics (K1 copy, K2/K3 permute, cast)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import kats

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name,size,pairs,gen", kats.COPY_KATS, ids=[k[0] for k in kats.COPY_KATS])
def test_vectorized_copy_kats(name, size, pairs, gen):
    src = kats.copy_kat_data(size, pairs, gen)
    dst = [np.zeros(size, dtype=np.uint8) for _ in range(pairs)]
    O.vectorized_copy(src, dst, size)
    for s, d in zip(src, dst):
        assert np.array_equal(s, d)


def test_vectorized_copy_unaligned_offsets():
    # K1 promises any alignment (tensor_kernels.cu:511-540): exercise every (src%16, dst%16) phase
    base = np.arange(4096, dtype=np.int64).astype(np.uint8)
    for so in (0, 1, 3, 4, 8, 15):
        for do in (0, 2, 4, 7, 8):
            dst = np.zeros(2048, dtype=np.uint8)
            O.vectorized_copy([base[so:so + 999]], [dst[do:do + 999]], 999)
            assert np.array_equal(dst[do:do + 999], base[so:so + 999])
            assert not dst[:do].any() and not dst[do + 999:].any()


@pytest.mark.parametrize("layout", [kats.NHD, kats.HND])
def test_permute_position_encoded_kat(layout):
    # kernel_roundtrip.rs:418-493: first-principles offsets
    d = kats.PERMUTE_DIMS
    nh, nl, no, nt, hd = d["nh"], d["nl"], d["no"], d["nt"], d["hd"]
    uni = kats.position_encoded_universal(**d)
    chunks = [np.full(nt * nh * hd, -1, dtype=np.float32) for _ in range(nl * no)]
    O.block_from_universal([uni.reshape(-1)], chunks, nh, nl, no, nt, hd, 4, layout)
    for l in range(nl):
        for o in range(no):
            blk = chunks[l * no + o]
            for t in range(nt):
                for h in range(nh):
                    for x in range(hd):
                        off = (t * nh + h) * hd + x if layout == kats.NHD else (h * nt + t) * hd + x
                        want = ((((h * nl + l) * no + o) * nt + t) * hd + x)
                        assert blk[off] == want
    # and it agrees with the ndarray-style reference make_blocks
    for got, want in zip(chunks, kats.make_blocks(uni, layout)):
        assert np.array_equal(got, want)


@pytest.mark.parametrize("dtype", [0, 1, 2, 3])
@pytest.mark.parametrize("layout", [kats.NHD, kats.HND])
def test_permute_roundtrip(dtype, layout):
    # kernel_roundtrip.rs:241-356 (random tensors, nb=3, poison 0xDE before reverse)
    d = kats.PERMUTE_DIMS
    nh, nl, no, nt, hd = d["nh"], d["nl"], d["no"], d["nt"], d["hd"]
    rng = np.random.default_rng(7)
    npd, elem = kats.DTYPES[dtype], kats.ELEM[dtype]
    unis = [(rng.random((nh, nl, no, nt, hd)) * 2 - 1).astype(npd) if dtype != 1
            else rng.integers(0, 65536, (nh, nl, no, nt, hd)).astype(np.uint16) for _ in range(kats.PERMUTE_NB)]
    ref_chunks = [c for u in unis for c in kats.make_blocks(u, layout)]
    out_unis = [np.zeros(u.size, dtype=npd) for u in unis]
    O.universal_from_block(out_unis, ref_chunks, nh, nl, no, nt, hd, elem, layout)
    for got, want in zip(out_unis, unis):
        assert np.array_equal(got, want.reshape(-1))
    poisoned = [np.frombuffer(bytes([0xDE]) * c.nbytes, dtype=npd).copy() for c in ref_chunks]
    O.block_from_universal(out_unis, poisoned, nh, nl, no, nt, hd, elem, layout)
    for got, want in zip(poisoned, ref_chunks):
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


def test_cast_up_matches_torch_table():
    table = np.load(os.path.join(GOLD, "fp8_e4m3_to_bf16_torch.npy"))
    got = O.cast_e4m3_to_bf16(np.arange(256, dtype=np.uint8))
    assert np.array_equal(got, table)


def test_cast_down_matches_torch_where_torch_is_finite():
    table = np.load(os.path.join(GOLD, "bf16_to_fp8_e4m3_torch.npy"))
    bits = np.arange(65536, dtype=np.uint16)
    got = O.cast_bf16_to_e4m3(bits)
    mag = bits & 0x7FFF
    is_nan_in = mag > 0x7F80
    torch_nan = (table & 0x7F) == 0x7F
    overflow = torch_nan & ~is_nan_in                 # torch: overflow -> NaN; ours: saturate to +-448
    same = ~overflow
    assert np.array_equal(got[same], table[same])
    assert np.array_equal(got[overflow] & 0x7F, np.full(overflow.sum(), 0x7E, dtype=np.uint8))
    assert np.array_equal(got[overflow] & 0x80, ((bits[overflow] >> 8) & 0x80).astype(np.uint8))
    # the overflow set is exactly |x| > 464 (464 itself ties-to-even to 448 in torch as well)
    assert mag[overflow].min() == 0x43E9  # first bf16 above 464.0 (0x43E8)


def test_cast_roundtrip_is_identity_on_fp8_codes():
    codes = np.array([c for c in range(256) if (c & 0x7F) != 0x7F], dtype=np.uint8)
    assert np.array_equal(O.cast_bf16_to_e4m3(O.cast_e4m3_to_bf16(codes)), codes)


def test_cast_transfer_config3_small():
    # configs[2] in miniature: fp8 LW source -> bf16 LW destination, distinct block tables
    nb, n = 16, 8
    src = O.Layout(O.LW, nb, 3, 2, 16, 64, 1, block_dim=O.BLOCK_IS_SECOND_DIM, allow_fp8=True)
    dst = O.Layout(O.LW, nb, 3, 2, 16, 64, 2, block_dim=O.BLOCK_IS_SECOND_DIM)
    rng = np.random.default_rng(3)
    for b in src.buffers:
        b[:] = rng.integers(0, 256, b.size, dtype=np.uint8)
    sid, did = rng.permutation(nb)[:n], rng.permutation(nb)[:n]
    O.execute_memcpy_transfer(src, dst, sid, did, cast_mode=1)
    table = np.load(os.path.join(GOLD, "fp8_e4m3_to_bf16_torch.npy"))
    for s, d in zip(sid, did):
        for l in range(3):
            for o in range(2):
                want = table[src.region_bytes(int(s), l, o)]
                got = dst.region_bytes(int(d), l, o).view(np.uint16)
                assert np.array_equal(got, want)


# ---- KvBlockLayout transforms (kv_block_layout.rs:85-95): the numpy definition vs the pinned K2 / K3 restatement ----
@pytest.mark.parametrize("layout,kv", [(kats.NHD, O.KV_OPERATIONAL_NHD), (kats.HND, O.KV_OPERATIONAL_HND)])
@pytest.mark.parametrize("dims", [(3, 2, 4, 5, 8, 2), (2, 1, 16, 8, 64, 2), (1, 2, 3, 1, 16, 4)])
def test_kv_layout_permute_agrees_with_k2_k3(layout, kv, dims):
    nl, no, nt, nh, hd, elem = dims
    row = hd * elem
    rng = np.random.default_rng(nl * 100 + nh)
    blk = rng.integers(0, 256, nl * no * nt * nh * row, dtype=np.uint8)
    chunk = nt * nh * row
    chunks = [blk[i * chunk:(i + 1) * chunk].copy() for i in range(nl * no)]
    uni = np.zeros_like(blk)
    O.universal_from_block([uni], chunks, nh, nl, no, nt, hd, elem, layout)
    assert np.array_equal(uni, O.kv_layout_permute(blk, kv, O.KV_UNIVERSAL_TP, nl, no, nt, nh, row))
    back = [np.zeros(chunk, dtype=np.uint8) for _ in range(nl * no)]
    O.block_from_universal([uni], back, nh, nl, no, nt, hd, elem, layout)
    assert np.array_equal(np.concatenate(back), O.kv_layout_permute(uni, O.KV_UNIVERSAL_TP, kv, nl, no, nt, nh, row))


def test_kv_layout_permute_first_principles_and_layer_ranges():
    nl, no, nt, nh, row = 3, 2, 4, 5, 16
    n = nl * no * nt * nh
    # element e of the NHD block carries its own (l, o, t, h) in its first four bytes
    blk = np.zeros((nl, no, nt, nh, row), dtype=np.uint8)
    for l in range(nl):
        for o in range(no):
            for t in range(nt):
                for h in range(nh):
                    blk[l, o, t, h, :4] = (l, o, t, h)
    flat = blk.reshape(-1)
    pp = O.kv_layout_permute(flat, O.KV_OPERATIONAL_NHD, O.KV_UNIVERSAL_PP, nl, no, nt, nh, row).reshape(nl, nh, no, nt, row)
    hnd = O.kv_layout_permute(flat, O.KV_OPERATIONAL_NHD, O.KV_OPERATIONAL_HND, nl, no, nt, nh, row).reshape(nl, no, nh, nt, row)
    tp = O.kv_layout_permute(flat, O.KV_OPERATIONAL_NHD, O.KV_UNIVERSAL_TP, nl, no, nt, nh, row).reshape(nh, nl, no, nt, row)
    for l in range(nl):
        for o in range(no):
            for t in range(nt):
                for h in range(nh):
                    assert tuple(pp[l, h, o, t, :4]) == (l, o, t, h)
                    assert tuple(hnd[l, o, h, t, :4]) == (l, o, t, h)
                    assert tuple(tp[h, l, o, t, :4]) == (l, o, t, h)
    # every pair composes to the identity
    for a in O.KV_DIM_ORDER:
        for b in O.KV_DIM_ORDER:
            x = O.kv_layout_permute(flat, O.KV_OPERATIONAL_NHD, a, nl, no, nt, nh, row)
            y = O.kv_layout_permute(x, a, b, nl, no, nt, nh, row)
            assert np.array_equal(O.kv_layout_permute(y, b, O.KV_OPERATIONAL_NHD, nl, no, nt, nh, row), flat)
    # a layer range moves only that layer's elements
    old = np.full(n * row, 0xEE, dtype=np.uint8)
    part = O.kv_layout_permute(flat, O.KV_OPERATIONAL_NHD, O.KV_UNIVERSAL_TP, nl, no, nt, nh, row, layers=range(1, 2), dst_old=old)
    part = part.reshape(nh, nl, no, nt, row)
    assert (part[:, 0] == 0xEE).all() and (part[:, 2] == 0xEE).all()
    assert np.array_equal(part[:, 1], tp[:, 1])

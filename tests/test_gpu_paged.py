"""GPU parity tests for the v2 block-table kernel (kvbm_kernels_paged_copy_v2) against the CPU oracle.

The cases are the reference's transfer tests (lib/kvbm-physical/src/transfer/tests/local_transfers.rs):
layout matrix FC/LW x FC/LW, layer ranges, guard blocks, layer composition -- with Device storage on
both sides, checked by per-block BLAKE3 (transfer/checksum.rs:91-158) and whole-pool equality.
"""
import itertools
import os

import numpy as np
import pytest
import torch

from dynamo_b200 import kernels as K
from oracle import oracle as O
from tests.gpu_util import DevicePool, dev_u8, ids_dev, make_layout, paged_dst, randomize, stream_ptr

pytestmark = pytest.mark.gpu

KINDS = [("FC", dict(kind=O.FC)), ("LWf", dict(kind=O.LW, block_dim=O.BLOCK_IS_FIRST_DIM)),
         ("LWs", dict(kind=O.LW, block_dim=O.BLOCK_IS_SECOND_DIM))]
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def run_paged(src_pool, dst_pools, sid_list, did_list, layers=None, cast=0, opts=None, replicate=False):
    nl = src_pool.host.num_layers
    lb, le = (layers.start, layers.stop) if layers is not None else (0, nl)
    shared = ids_dev(sid_list[0]) if replicate else None
    keep, dsts = [], []
    for pool, sid, did in zip(dst_pools, sid_list, did_list):
        s = shared if replicate else ids_dev(sid)
        d = ids_dev(did)
        keep += [s, d]
        dsts.append(paged_dst(pool, s, d))
    rc = K.paged_copy(src_pool.desc, dsts, len(sid_list[0]), lb, le, cast, opts, stream_ptr())
    torch.cuda.synchronize()
    return rc


def check_against_oracle(src_h, dst_h, dst_pool, sid, did, layers=None, cast=0):
    """Run the oracle on the host twins, then require the device pool to equal it byte for byte."""
    O.execute_memcpy_transfer(src_h, dst_h, sid, did, layers, cast_mode=cast)
    want_sums = dst_h.block_checksums(did, layers)
    want = [b.copy() for b in dst_h.buffers]
    got = dst_pool.snapshot()
    for w, g in zip(want, got):
        assert np.array_equal(w, g)
    dst_pool.download()
    assert dst_h.block_checksums(did, layers) == want_sums


@pytest.mark.parametrize("src_kw,dst_kw", list(itertools.product([k[1] for k in KINDS], repeat=2)),
                         ids=[f"{a[0]}-{b[0]}" for a, b in itertools.product(KINDS, repeat=2)])
@pytest.mark.parametrize("layers", [None, range(0, 1), range(1, 2)], ids=["full", "layer0", "layer1"])
def test_p2p_layout_matrix(src_kw, dst_kw, layers):
    # local_transfers.rs:108-171 (+ TransferMode layer0/layer1 of tests/mod.rs:100-113), Device -> Device
    src_h, dst_h = make_layout(nb=6, **src_kw), make_layout(nb=6, fill=0, **dst_kw)
    src_h.fill_blocks([0, 1], -1)
    dst_h.fill_blocks([2, 5], 0xFF)               # guard blocks (local_transfers.rs:498-505)
    src_p, dst_p = DevicePool(src_h), DevicePool(dst_h)
    assert run_paged(src_p, [dst_p], [[0, 1]], [[3, 4]], layers) == 0
    check_against_oracle(src_h, dst_h, dst_p, [0, 1], [3, 4], layers)
    assert (dst_h.region_bytes(2, 0, 0) == 0xFF).all() and (dst_h.region_bytes(5, 1, 1) == 0xFF).all()


@pytest.mark.parametrize("src_kw,dst_kw", list(itertools.product([k[1] for k in KINDS], repeat=2)))
def test_layer_composition_equals_full_block(src_kw, dst_kw):
    # local_transfers.rs:922-...
    src_h = make_layout(nb=4, **src_kw)
    src_h.fill_blocks([0, 1], -1)
    full_h, lay_h = make_layout(nb=4, **dst_kw), make_layout(nb=4, **dst_kw)
    src_p, full_p, lay_p = DevicePool(src_h), DevicePool(full_h), DevicePool(lay_h)
    assert run_paged(src_p, [full_p], [[0, 1]], [[2, 3]]) == 0
    assert run_paged(src_p, [lay_p], [[0, 1]], [[2, 3]], range(0, 1)) == 0
    assert run_paged(src_p, [lay_p], [[0, 1]], [[2, 3]], range(1, 2)) == 0
    for a, b in zip(full_p.snapshot(), lay_p.snapshot()):
        assert np.array_equal(a, b)
    full_p.download()
    assert full_h.block_checksums([2, 3]) == {2: src_h.block_checksum(0), 3: src_h.block_checksum(1)}


@pytest.mark.parametrize("geom", [
    dict(nl=4, no=2, page=16, inner=1024, dt=2),     # Llama-3-8B bf16: 32 KiB regions (2 tiles each)
    dict(nl=5, no=2, page=16, inner=256, dt=2),      # Llama-3-70B TP4: 8 KiB regions
    dict(nl=3, no=1, page=16, inner=576, dt=2),      # MLA-style outer_dim=1, 18 KiB regions (tile tail)
    dict(nl=2, no=2, page=3, inner=37, dt=2),        # 222 B regions: not a multiple of 16 -> SIMT ladder
    dict(nl=2, no=2, page=16, inner=100, dt=4),      # 6400 B
], ids=["8b", "70b-tp4", "mla", "odd222", "f32"])
@pytest.mark.parametrize("force_simt", [0, 1], ids=["tma", "simt"])
def test_random_block_tables_vllm_layout(geom, force_simt):
    # SURVEY §8(d) cfg2 in miniature: LW/BlockIsSecondDim pools, random non-contiguous ids, random bytes
    nb, n = 48, 20
    mk = lambda: make_layout(O.LW, nb, block_dim=O.BLOCK_IS_SECOND_DIM, **geom)
    src_h, dst_h = mk(), mk()
    randomize(src_h, 1234)
    randomize(dst_h, 99)
    sid = np.random.default_rng(0).permutation(nb)[:n]
    did = np.random.default_rng(1).permutation(nb)[:n]
    src_p, dst_p = DevicePool(src_h), DevicePool(dst_h)
    opts = K.PagedCopyOpts(force_simt=force_simt)
    assert run_paged(src_p, [dst_p], [sid], [did], opts=opts) == 0
    check_against_oracle(src_h, dst_h, dst_p, sid, did)


@pytest.mark.parametrize("warps,stages,tile", [(1, 2, 4096), (2, 5, 8192), (4, 3, 16384), (8, 2, 2048), (3, 16, 512), (2, 1, 4096),
                                                (16, 3, 1024)])
def test_ring_geometries(warps, stages, tile):
    nb, n = 32, 16
    mk = lambda: make_layout(O.LW, nb, nl=3, no=2, page=16, inner=512, dt=2, block_dim=O.BLOCK_IS_SECOND_DIM)
    src_h, dst_h = mk(), mk()
    randomize(src_h, 7)
    sid, did = np.arange(n)[::-1].copy(), np.arange(n) + 10
    src_p, dst_p = DevicePool(src_h), DevicePool(dst_h)
    opts = K.PagedCopyOpts(warps_per_cta=warps, stages=stages, tile_bytes=tile, max_ctas=7)
    assert run_paged(src_p, [dst_p], [sid], [did], opts=opts) == 0
    check_against_oracle(src_h, dst_h, dst_p, sid, did)


@pytest.mark.parametrize("replicate", [False, True], ids=["distinct", "replicate"])
@pytest.mark.parametrize("ndst", [2, 4, 7])
def test_fan_out_to_many_destinations(ndst, replicate):
    # 1 -> N, here N pools on one device (peer mappings are plain pointers to the kernel)
    nb, n = 40, 12
    mk = lambda: make_layout(O.LW, nb, nl=3, no=2, page=16, inner=256, dt=2, block_dim=O.BLOCK_IS_SECOND_DIM)
    src_h = mk()
    randomize(src_h, 21)
    dst_hs = [mk() for _ in range(ndst)]
    for i, d in enumerate(dst_hs):
        randomize(d, 100 + i)
    rng = np.random.default_rng(5)
    sids = [rng.permutation(nb)[:n] for _ in range(ndst)]
    if replicate:
        sids = [sids[0]] * ndst
    dids = [np.random.default_rng(50 + i).permutation(nb)[:n] for i in range(ndst)]
    src_p, dst_ps = DevicePool(src_h), [DevicePool(d) for d in dst_hs]
    assert run_paged(src_p, dst_ps, sids, dids, replicate=replicate) == 0
    for d_h, d_p, sid, did in zip(dst_hs, dst_ps, sids, dids):
        check_against_oracle(src_h, d_h, d_p, sid, did)


@pytest.mark.parametrize("force_simt", [0, 1], ids=["tma", "simt"])
def test_fused_cast_fp8_to_bf16_all_codes(force_simt):
    # BASELINE configs[2] in miniature; source bytes cover all 256 e4m3 codes
    nb, n = 24, 10
    src_h = make_layout(O.LW, nb, nl=3, no=2, page=16, inner=256, dt=1, block_dim=O.BLOCK_IS_SECOND_DIM, allow_fp8=True)
    dst_h = make_layout(O.LW, nb, nl=3, no=2, page=16, inner=256, dt=2, block_dim=O.BLOCK_IS_SECOND_DIM)
    randomize(src_h, 3)
    src_h.buffers[0][:256] = np.arange(256, dtype=np.uint8)
    sid = np.array([0] + list(np.random.default_rng(2).permutation(np.arange(1, nb))[:n - 1]))
    did = np.random.default_rng(4).permutation(nb)[:n]
    src_p, dst_p = DevicePool(src_h), DevicePool(dst_h)
    opts = K.PagedCopyOpts(force_simt=force_simt)
    assert run_paged(src_p, [dst_p], [sid], [did], cast=K.CastMode.FP8E4M3_TO_BF16, opts=opts) == 0
    check_against_oracle(src_h, dst_h, dst_p, sid, did, cast=1)
    # and directly against the torch-generated golden table
    table = np.load(os.path.join(GOLD, "fp8_e4m3_to_bf16_torch.npy"))
    got = dst_h.region_bytes(int(did[0]), 0, 0).view(np.uint16)[:256]
    assert np.array_equal(got, table)


@pytest.mark.parametrize("force_simt", [0, 1], ids=["tma", "simt"])
def test_fused_cast_bf16_to_fp8_all_bit_patterns(force_simt):
    # 65 536 bf16 patterns = 128 KiB = exactly 4 regions of 32 KiB
    nb = 8
    src_h = make_layout(O.LW, nb, nl=2, no=2, page=16, inner=1024, dt=2, block_dim=O.BLOCK_IS_SECOND_DIM)
    dst_h = make_layout(O.LW, nb, nl=2, no=2, page=16, inner=1024, dt=1, block_dim=O.BLOCK_IS_SECOND_DIM, allow_fp8=True)
    randomize(src_h, 8)
    allbits = np.arange(65536, dtype=np.uint16)
    src_h.region_bytes(3, 0, 0).view(np.uint16)[:] = allbits[:16384]
    src_h.region_bytes(3, 0, 1).view(np.uint16)[:] = allbits[16384:32768]
    src_h.region_bytes(3, 1, 0).view(np.uint16)[:] = allbits[32768:49152]
    src_h.region_bytes(3, 1, 1).view(np.uint16)[:] = allbits[49152:]
    sid, did = [3, 5, 1], [0, 7, 2]
    src_p, dst_p = DevicePool(src_h), DevicePool(dst_h)
    opts = K.PagedCopyOpts(force_simt=force_simt)
    assert run_paged(src_p, [dst_p], [sid], [did], cast=K.CastMode.BF16_TO_FP8E4M3, opts=opts) == 0
    check_against_oracle(src_h, dst_h, dst_p, sid, did, cast=2)


def test_argument_validation():
    h = make_layout(O.LW, 8, block_dim=O.BLOCK_IS_SECOND_DIM)
    other = make_layout(O.LW, 8, nl=3, block_dim=O.BLOCK_IS_SECOND_DIM)
    small = make_layout(O.LW, 8, inner=64, block_dim=O.BLOCK_IS_SECOND_DIM)
    p, q, r = DevicePool(h), DevicePool(other), DevicePool(small)
    ids = ids_dev([0, 1])
    sp = stream_ptr()
    ok = paged_dst(p, ids, ids)
    assert K.paged_copy(p.desc, [ok], 0, 0, 2, 0, None, sp) == 0                      # empty -> no-op
    assert K.paged_copy(p.desc, [ok], 2, 1, 1, 0, None, sp) == 0                      # empty layer range
    assert K.paged_copy(p.desc, [ok], 2, 0, 3, 0, None, sp) == K.CUDA_ERROR_INVALID_VALUE   # layer range too long
    assert K.paged_copy(p.desc, [paged_dst(q, ids, ids)], 2, 0, 2, 0, None, sp) == K.CUDA_ERROR_INVALID_VALUE  # cuda.rs:52-58
    assert K.paged_copy(p.desc, [paged_dst(r, ids, ids)], 2, 0, 2, 0, None, sp) == K.CUDA_ERROR_INVALID_VALUE  # memcpy.rs:143
    assert K.paged_copy(p.desc, [ok] * 9, 2, 0, 2, 0, None, sp) == K.CUDA_ERROR_INVALID_VALUE
    assert K.paged_copy(p.desc, [ok], 2, 0, 2, 7, None, sp) == K.CUDA_ERROR_INVALID_VALUE
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    need_ws = K.PagedDst(p.desc, ids.data_ptr(), ids.data_ptr(), flag.data_ptr(), 0)
    assert K.paged_copy(p.desc, [need_ws], 2, 0, 2, 0, None, sp) == K.CUDA_ERROR_INVALID_VALUE  # flag without workspace


def test_completion_and_layer_flags():
    nb, n, nl = 32, 16, 6
    mk = lambda: make_layout(O.LW, nb, nl=nl, no=2, page=16, inner=256, dt=2, block_dim=O.BLOCK_IS_SECOND_DIM)
    src_h, d0_h, d1_h = mk(), mk(), mk()
    randomize(src_h, 31)
    sid, did0, did1 = np.arange(n), np.arange(n) + 8, np.arange(n)[::-1].copy() + 3
    src_p, d0_p, d1_p = DevicePool(src_h), DevicePool(d0_h), DevicePool(d1_h)
    flags = torch.zeros(2, dtype=torch.int32, device="cuda")
    layer_flags = torch.zeros(2, nl, dtype=torch.int32, device="cuda")
    ws = torch.zeros(nl + 4, dtype=torch.int32, device="cuda")
    ready = torch.zeros(nl, dtype=torch.int32, device="cuda")
    s_ids = ids_dev(sid)
    a, b = ids_dev(did0), ids_dev(did1)
    dsts = [K.PagedDst(d0_p.desc, s_ids.data_ptr(), a.data_ptr(), flags[0:].data_ptr(), layer_flags[0].data_ptr()),
            K.PagedDst(d1_p.desc, s_ids.data_ptr(), b.data_ptr(), flags[1:].data_ptr(), layer_flags[1].data_ptr())]
    side = torch.cuda.Stream()
    for epoch in (1, 2):      # run twice: the kernel must leave the workspace zeroed
        opts = K.PagedCopyOpts(epoch=epoch, layer_ready_flags=ready.data_ptr(), sync_workspace=ws.data_ptr(), max_ctas=8)
        torch.cuda.synchronize()
        assert K.paged_copy(src_p.desc, dsts, n, 0, nl, 0, opts, stream_ptr(side)) == 0
        # the transfer kernel is now spinning on ready[0]; release the layers one at a time from another stream
        for l in range(nl):
            assert K.set_flags(ready.data_ptr(), l, 1, epoch, stream_ptr()) == 0
        assert K.wait_flag(flags[1:].data_ptr(), epoch, stream_ptr()) == 0     # device-side wait on the done flag
        torch.cuda.synchronize()
        assert flags.tolist() == [epoch, epoch]
        assert layer_flags.tolist() == [[epoch] * nl, [epoch] * nl]
        assert ws.tolist() == [0] * (nl + 4)
    check_against_oracle(src_h, d0_h, d0_p, sid, did0)
    check_against_oracle(src_h, d1_h, d1_p, sid, did1)


def test_full_size_config2_roundtrip_property():
    # BASELINE configs[1] at full size on one device: 256 blocks x 32 layers x K/V x 32 KiB = 512 MiB.
    # Size-independent property: gather->scatter with a permutation, then the inverse permutation,
    # restores the source pool exactly; plus a BLAKE3 spot check of one block against the oracle formula.
    nb_pool, n, nl = 320, 256, 32
    region = 16 * 1024 * 2
    def pool():
        bufs = [torch.empty(2 * nb_pool * region, dtype=torch.uint8, device="cuda") for _ in range(nl)]
        base = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
        return bufs, base, K.PagedLayout(base.data_ptr(), region, region * nb_pool, region, nl, 2, nb_pool)
    a_bufs, a_base, a = pool()
    b_bufs, b_base, b = pool()
    c_bufs, c_base, c = pool()
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in a_bufs:
        t.copy_(torch.randint(0, 256, t.shape, dtype=torch.uint8, device="cuda", generator=g))
    for t in b_bufs + c_bufs:
        t.zero_()
    sid = np.random.default_rng(0).permutation(nb_pool)[:n]
    did = np.random.default_rng(1).permutation(nb_pool)[:n]
    s, d = ids_dev(sid), ids_dev(did)
    sp = stream_ptr()
    assert K.paged_copy(a, [K.PagedDst(b, s.data_ptr(), d.data_ptr(), 0, 0)], n, 0, nl, 0, None, sp) == 0
    assert K.paged_copy(b, [K.PagedDst(c, d.data_ptr(), s.data_ptr(), 0, 0)], n, 0, nl, 0, None, sp) == 0
    torch.cuda.synchronize()
    moved = torch.zeros(nb_pool, dtype=torch.bool, device="cuda")
    moved[torch.from_numpy(sid).cuda()] = True
    for l in range(nl):
        av = a_bufs[l].view(2, nb_pool, region)
        cv = c_bufs[l].view(2, nb_pool, region)
        assert torch.equal(cv[:, moved], av[:, moved])
        assert not cv[:, ~moved].any()
    import blake3
    h1, h2 = blake3.blake3(), blake3.blake3()
    for l in range(nl):
        for o in range(2):
            h1.update(a_bufs[l].view(2, nb_pool, region)[o, int(sid[5])].cpu().numpy().tobytes())
            h2.update(b_bufs[l].view(2, nb_pool, region)[o, int(did[5])].cpu().numpy().tobytes())
    assert h1.hexdigest() == h2.hexdigest()


# ------------------------------------------------------------------ full-size properties of the other BASELINE configs
def _torch_pool(nl, nbp, region, fill=None, seed=0):
    bufs = [torch.empty(2 * nbp * region, dtype=torch.uint8, device="cuda") for _ in range(nl)]
    g = torch.Generator(device="cuda").manual_seed(seed)
    for t in bufs:
        if fill is None:
            t.copy_(torch.randint(0, 256, t.shape, dtype=torch.uint8, device="cuda", generator=g))
        else:
            t.fill_(fill)
    base = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
    return bufs, base, K.PagedLayout(base.data_ptr(), region, region * nbp, region, nl, 2, nbp)


def test_full_size_config3_fp8_roundtrip_is_identity():
    # BASELINE configs[2] at full size: 256 blocks x 32 layers x K/V, fp8 source (16 KiB regions) -> bf16 (32 KiB).
    # Property: fp8 -> bf16 -> fp8 through the two fused-cast kernels is the identity on every non-NaN code, and NaN
    # codes come back as NaN (0x7f | sign is not preserved by torch's up-cast: 0xff -> 0x7fc0 -> 0x7f).
    nl, nbp, n = 32, 288, 256
    a_bufs, a_base, a = _torch_pool(nl, nbp, 16384, seed=3)
    b_bufs, b_base, b = _torch_pool(nl, nbp, 32768, fill=0)
    c_bufs, c_base, c = _torch_pool(nl, nbp, 16384, fill=0)
    sid = ids_dev(np.random.default_rng(0).permutation(nbp)[:n])
    did = ids_dev(np.random.default_rng(1).permutation(nbp)[:n])
    sp = stream_ptr()
    assert K.paged_copy(a, [K.PagedDst(b, sid.data_ptr(), did.data_ptr(), 0, 0)], n, 0, nl, K.CastMode.FP8E4M3_TO_BF16, None, sp) == 0
    assert K.paged_copy(b, [K.PagedDst(c, did.data_ptr(), sid.data_ptr(), 0, 0)], n, 0, nl, K.CastMode.BF16_TO_FP8E4M3, None, sp) == 0
    torch.cuda.synchronize()
    rows = sid.long()
    for l in (0, 7, nl - 1):
        av = a_bufs[l].view(2, nbp, 16384)[:, rows]
        cv = c_bufs[l].view(2, nbp, 16384)[:, rows]
        nan = (av & 0x7F) == 0x7F
        assert torch.equal(cv[~nan], av[~nan])
        assert bool(((cv[nan] & 0x7F) == 0x7F).all())
    # and EVERY moved region (256 blocks x 32 layers x K/V) against the torch-generated golden table, on the device
    table = torch.from_numpy(np.load(os.path.join(GOLD, "fp8_e4m3_to_bf16_torch.npy")).astype(np.uint16).view(np.int16).copy()).cuda()
    drows = did.long()
    for l in range(nl):
        codes = a_bufs[l].view(2, nbp, 16384)[:, rows].long()                 # [2, n, 16384] fp8 codes
        got = b_bufs[l].view(2, nbp, 32768)[:, drows].view(torch.int16)       # [2, n, 16384] bf16 bit patterns
        assert torch.equal(got, table[codes]), f"layer {l} differs from the golden fp8->bf16 table"
    untouched = torch.ones(nbp, dtype=torch.bool, device="cuda")
    untouched[drows] = False
    assert not b_bufs[9].view(2, nbp, 32768)[:, untouched].any()


def test_full_size_config4_layer_stream_equals_one_shot():
    # BASELINE configs[3] per rank: Llama-3-70B TP4 shard, 80 layers, 8 KiB regions, 256 blocks = 320 MiB.
    # Property: 80 single-layer transfers (layer_range = l..l+1, the reference's streaming pattern) compose to exactly
    # the one-shot transfer, which itself equals a gather by index.
    nl, nbp, n, region = 80, 320, 256, 8192
    s_bufs, s_base, s = _torch_pool(nl, nbp, region, seed=4)
    one_bufs, one_base, one = _torch_pool(nl, nbp, region, fill=0)
    lay_bufs, lay_base, lay = _torch_pool(nl, nbp, region, fill=0)
    sid_np = np.random.default_rng(2).permutation(nbp)[:n]
    did_np = np.random.default_rng(3).permutation(nbp)[:n]
    sid, did = ids_dev(sid_np), ids_dev(did_np)
    sp = stream_ptr()
    assert K.paged_copy(s, [K.PagedDst(one, sid.data_ptr(), did.data_ptr(), 0, 0)], n, 0, nl, 0, None, sp) == 0
    for l in range(nl):
        assert K.paged_copy(s, [K.PagedDst(lay, sid.data_ptr(), did.data_ptr(), 0, 0)], n, l, l + 1, 0, None, sp) == 0
    torch.cuda.synchronize()
    for l in range(nl):
        assert torch.equal(one_bufs[l], lay_bufs[l])
    for l in (0, 39, 79):
        assert torch.equal(one_bufs[l].view(2, nbp, region)[:, did.long()], s_bufs[l].view(2, nbp, region)[:, sid.long()])
    untouched = torch.ones(nbp, dtype=torch.bool, device="cuda")
    untouched[did.long()] = False
    assert not one_bufs[11].view(2, nbp, region)[:, untouched].any()


def test_full_size_config5_16k_ctx_fanout_and_checksum_of_checksums():
    # BASELINE configs[4] shape at 16 k ctx: 1024 blocks x 32 layers x K/V x 32 KiB = 2 GiB, replicated to 2 pools in
    # one launch.  Property: the wrapping 64-bit lane sum over every moved region is permutation-invariant -> equal on source and both
    # destinations (a checksum of checksums), plus exact equality on sampled layers.
    nl, nbp, n, region = 32, 1100, 1024, 32768
    s_bufs, s_base, s = _torch_pool(nl, nbp, region, seed=5)      # keep the layer_base tables alive (device memory)
    d0_bufs, d0_base, d0 = _torch_pool(nl, nbp, region, fill=0)
    d1_bufs, d1_base, d1 = _torch_pool(nl, nbp, region, fill=0)
    sid = ids_dev(np.random.default_rng(7).permutation(nbp)[:n])
    da = ids_dev(np.random.default_rng(8).permutation(nbp)[:n])
    db = ids_dev(np.random.default_rng(9).permutation(nbp)[:n])
    dsts = [K.PagedDst(d0, sid.data_ptr(), da.data_ptr(), 0, 0), K.PagedDst(d1, sid.data_ptr(), db.data_ptr(), 0, 0)]
    assert K.paged_copy(s, dsts, n, 0, nl, 0, None, stream_ptr()) == 0
    torch.cuda.synchronize()

    def fold(bufs, rows):
        acc = torch.zeros(region // 8, dtype=torch.int64, device="cuda")
        for t in bufs:
            v = t.view(2, nbp, region)[:, rows].reshape(-1, region // 8, 8).view(torch.int64).reshape(-1, region // 8)
            acc += v.sum(0)   # wrapping int64 sum of every 8-byte lane: independent of block order
        return acc
    ref = fold(s_bufs, sid.long())
    assert torch.equal(fold(d0_bufs, da.long()), ref)
    assert torch.equal(fold(d1_bufs, db.long()), ref)
    for l in (0, 31):
        assert torch.equal(d0_bufs[l].view(2, nbp, region)[:, da.long()], s_bufs[l].view(2, nbp, region)[:, sid.long()])
        assert torch.equal(d1_bufs[l].view(2, nbp, region)[:, db.long()], s_bufs[l].view(2, nbp, region)[:, sid.long()])


def test_edge_geometries_single_block_many_layers_and_tiny_regions():
    # ragged / extreme shapes: one block, 200 layers, 16-byte regions; 1 layer, outer 1; region just over one tile
    for geom, nb, ids in [(dict(nl=200, no=2, page=1, inner=8, dt=2), 3, ([2], [0])),
                          (dict(nl=1, no=1, page=16, inner=1024, dt=2), 5, ([0, 4, 2], [1, 3, 0])),
                          (dict(nl=2, no=2, page=16, inner=520, dt=2), 6, ([5, 1], [0, 2]))]:
        src_h = make_layout(O.LW, nb, block_dim=O.BLOCK_IS_SECOND_DIM, **geom)
        dst_h = make_layout(O.LW, nb, block_dim=O.BLOCK_IS_SECOND_DIM, **geom)
        randomize(src_h, 41)
        randomize(dst_h, 42)
        src_p, dst_p = DevicePool(src_h), DevicePool(dst_h)
        assert run_paged(src_p, [dst_p], [ids[0]], [ids[1]]) == 0
        check_against_oracle(src_h, dst_h, dst_p, ids[0], ids[1])


def test_gate_timeout_aborts_instead_of_hanging():
    # a gated launch whose producer never releases the layer must give up: completion word 0xFFFFFFFF, done flag untouched
    nb, n, nl = 16, 8, 3
    mk = lambda: make_layout(O.LW, nb, nl=nl, no=2, page=16, inner=256, dt=2, block_dim=O.BLOCK_IS_SECOND_DIM)
    src_h, dst_h = mk(), mk()
    randomize(src_h, 1)
    src_p, dst_p = DevicePool(src_h), DevicePool(dst_h)
    ready = torch.zeros(nl, dtype=torch.int32, device="cuda")
    done = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws = torch.zeros(nl + 4, dtype=torch.int32, device="cuda")
    host_word = torch.zeros(16, dtype=torch.int32).pin_memory()
    s, d = ids_dev(range(n)), ids_dev(range(n, 2 * n))
    K.check(K.set_flags(ready.data_ptr(), 0, 1, 9, stream_ptr()))          # only layer 0 is ever released
    opts = K.PagedCopyOpts(epoch=9, layer_ready_flags=ready.data_ptr(), sync_workspace=ws.data_ptr(), max_ctas=4,
                           completion_flag=host_word.data_ptr(), completion_value=123, gate_timeout_ms=300)
    side = torch.cuda.Stream()
    assert K.paged_copy(src_p.desc, [K.PagedDst(dst_p.desc, s.data_ptr(), d.data_ptr(), done.data_ptr(), 0)], n, 0, nl, 0, opts,
                        stream_ptr(side)) == 0
    side.synchronize()                                                      # returns: the kernel did not hang
    assert (int(host_word[0]) & 0xFFFFFFFF) == 0xFFFFFFFF and done.tolist() == [0]
    assert ws.tolist() == [0] * (nl + 4)                                     # workspace left clean for the next launch
    # layer 0 (released) was copied, later layers were not
    dst_p.download()
    assert dst_h.block_checksums(range(n, 2 * n), range(0, 1)) == {n + i: src_h.block_checksum(i, range(0, 1)) for i in range(n)}
    assert not dst_h.region_bytes(n, 2, 0).any()


@pytest.mark.parametrize("static_schedule", [0, 1], ids=["tickets", "static"])
def test_tile_scheduler_modes_and_pool_slots_without_workspace(static_schedule):
    """The dynamic tile scheduler needs zeroed control words.  Launches that bring no workspace borrow a slot of the library's
    per-device pool (64 slots, reused once the launch behind them has finished); 150 back-to-back launches cycle it twice
    and every one must still copy exactly.  static_schedule=1 is the round-robin split (diagnostics)."""
    nb, n = 48, 40
    mk = lambda: make_layout(O.LW, nb, nl=4, no=2, page=16, inner=512, dt=2, block_dim=O.BLOCK_IS_SECOND_DIM)
    src_h, dst_h = mk(), mk()
    randomize(src_h, 77)
    src_p, dst_p = DevicePool(src_h), DevicePool(dst_h)
    rng = np.random.default_rng(3)
    sp = stream_ptr()
    last = None
    for it in range(150 if not static_schedule else 3):
        sid, did = rng.permutation(nb)[:n], rng.permutation(nb)[:n]
        s, d = ids_dev(sid), ids_dev(did)
        opts = K.PagedCopyOpts(static_schedule=static_schedule, max_ctas=5 + it % 7)
        assert K.paged_copy(src_p.desc, [K.PagedDst(dst_p.desc, s.data_ptr(), d.data_ptr(), 0, 0)], n, 0, 4, 0, opts, sp) == 0
        last = (sid, did)
        if it % 50 == 49 or static_schedule:
            torch.cuda.synchronize()
            dst_p.download()
            assert dst_h.block_checksums(did) == {int(b): src_h.block_checksum(int(a)) for a, b in zip(sid, did)}
    torch.cuda.synchronize()
    dst_p.download()
    sid, did = last
    assert dst_h.block_checksums(did) == {int(b): src_h.block_checksum(int(a)) for a, b in zip(sid, did)}


def test_gate_on_the_stream_when_nothing_may_spin():
    """gate_mode = STREAM_WAIT (what AUTO picks under lazy module loading): the ready flags are waited for by the stream
    front-end (cuStreamWaitValue32) and each layer is its own launch -- same bytes, same flags, no resident spinner."""
    nb, n, nl = 32, 16, 5
    mk = lambda: make_layout(O.LW, nb, nl=nl, no=2, page=16, inner=256, dt=2, block_dim=O.BLOCK_IS_SECOND_DIM)
    src_h, dst_h = mk(), mk()
    randomize(src_h, 5)
    sid, did = np.arange(n), np.arange(n)[::-1].copy() + 9
    src_p, dst_p = DevicePool(src_h), DevicePool(dst_h)
    done = torch.zeros(1, dtype=torch.int32, device="cuda")
    layer_done = torch.zeros(nl, dtype=torch.int32, device="cuda")
    ws = torch.zeros(K.sync_workspace_words(nl), dtype=torch.int32, device="cuda")
    ready = torch.zeros(nl, dtype=torch.int32, device="cuda")
    host_word = torch.zeros(16, dtype=torch.int32).pin_memory()
    s, d = ids_dev(sid), ids_dev(did)
    side = torch.cuda.Stream()
    assert K.gate_would_spin() is True          # tests run with CUDA_MODULE_LOADING=EAGER (conftest)
    before = K.launch_count()
    opts = K.PagedCopyOpts(epoch=4, layer_ready_flags=ready.data_ptr(), sync_workspace=ws.data_ptr(), gate_mode=K.GATE_STREAM_WAIT,
                           completion_flag=host_word.data_ptr(), completion_value=77)
    assert K.paged_copy(src_p.desc, [K.PagedDst(dst_p.desc, s.data_ptr(), d.data_ptr(), done.data_ptr(), layer_done.data_ptr())],
                        n, 0, nl, 0, opts, stream_ptr(side)) == 0
    assert K.launch_count() - before == nl       # one single-layer launch per layer
    for l in range(nl):
        assert int(host_word[0]) == 0 and done.tolist() == [0]               # nothing announced before the last layer
        assert K.set_flags(ready.data_ptr(), l, 1, 4, stream_ptr()) == 0
    side.synchronize()
    torch.cuda.synchronize()
    assert done.tolist() == [4] and layer_done.tolist() == [4] * nl and int(host_word[0]) == 77
    assert ws.tolist() == [0] * K.sync_workspace_words(nl)
    check_against_oracle(src_h, dst_h, dst_p, sid, did)


def test_launches_can_be_captured_into_a_cuda_graph_and_replayed():
    """An engine that runs its decode step from a CUDA graph can capture the hand-off with it: every entry point is plain
    stream work (kernel launches only; with a caller workspace the tile scheduler stays dynamic, without one a capturing
    stream gets the static split because the library's pool is event-guarded).  Replays must move the CURRENT bytes."""
    nl, no, nt, nh, hd = 4, 2, 16, 4, 64
    nb, n = 24, 16
    mk = lambda fill=0: make_layout(O.LW, nb, nl=nl, no=no, page=nt, inner=nh * hd, dt=2, block_dim=O.BLOCK_IS_SECOND_DIM, fill=fill)
    src_h, dst_a, dst_b = mk(), mk(), mk()
    uni_h = make_layout(O.FC, nb, nl=nl, no=no, page=nt, inner=nh * hd, dt=2)
    randomize(src_h, 5)
    src_p, a_p, b_p, u_p = DevicePool(src_h), DevicePool(dst_a), DevicePool(dst_b), DevicePool(uni_h)
    rng = np.random.default_rng(8)
    sid, did = rng.permutation(nb)[:n], rng.permutation(nb)[:n]
    s, d = ids_dev(sid), ids_dev(did)
    ws = torch.zeros(K.sync_workspace_words(nl), dtype=torch.int32, device="cuda:0")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda:0")

    def enqueue(sp):
        o = K.PagedCopyOpts(epoch=3, sync_workspace=ws.data_ptr())
        assert K.paged_copy(src_p.desc, [K.PagedDst(a_p.desc, s.data_ptr(), d.data_ptr(), flag.data_ptr(), 0)], n, 0, nl, 0, o, sp) == 0
        assert K.paged_copy(src_p.desc, [K.PagedDst(b_p.desc, s.data_ptr(), d.data_ptr(), 0, 0)], n, 0, nl, 0, None, sp) == 0
        assert K.paged_permute(K.PermuteSide(src_p.desc, s.data_ptr(), int(K.KvBlockLayout.OperationalNHD)),
                               K.PermuteSide(u_p.desc, d.data_ptr(), int(K.KvBlockLayout.UniversalTP)), n, 0, nl, nh, nt, hd * 2, stream=sp) == 0

    enqueue(stream_ptr())            # warm-up outside the capture: kernels loaded, attributes set
    torch.cuda.synchronize()
    for p in (a_p, b_p, u_p):
        for t in p.bufs:
            t.zero_()
    flag.zero_()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.graph(graph, stream=side):
        enqueue(stream_ptr(torch.cuda.current_stream()))
    torch.cuda.synchronize()
    assert int(flag.item()) == 0 and not any(bool(t.any()) for t in a_p.bufs)      # captured, not executed
    for round_ in range(2):
        if round_:                                                              # new bytes in the same source pool
            randomize(src_h, 99)
            for t, b in zip(src_p.bufs, src_h.buffers):
                t.copy_(torch.from_numpy(b))
            for p in (a_p, b_p, u_p):
                for t in p.bufs:
                    t.zero_()
            flag.zero_()
            torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        assert int(flag.item()) == 3 and ws.tolist() == [0] * K.sync_workspace_words(nl)
        want = {int(b): src_h.block_checksum(int(a)) for a, b in zip(sid, did)}
        for p, h in ((a_p, dst_a), (b_p, dst_b)):
            p.download()
            assert h.block_checksums(did) == want
        u_p.download()
        for a, b in zip(sid, did):
            blk = O.kv_layout_permute(O.read_logical_block(src_h, int(a)), O.KV_OPERATIONAL_NHD, O.KV_UNIVERSAL_TP, nl, no, nt, nh, hd * 2)
            assert np.array_equal(O.read_logical_block(uni_h, int(b)), blk)

"""Pins the oracle's layout arithmetic against the reference's own known-answer tests.

KATs transcribed from (relative to /root/reference):
  lib/kvbm-physical/src/layout/fully_contiguous.rs:356-424   (FC creation + memory_region)
  lib/kvbm-physical/src/layout/layer_separate.rs:345-430     (LW creation + memory_region)
  lib/kvbm-physical/src/layout/config.rs:16-44,165-180       (validation ranges)
"""
import pytest

from oracle import oracle as O


def test_fc_creation_required_bytes():
    # fully_contiguous.rs:361-379
    L = O.Layout(O.FC, 10, 4, 2, 16, 128, 2, bases=[0x1000])
    assert L.required_bytes() == 10 * 4 * 2 * 16 * 128 * 2
    assert L.num_blocks == 10 and L.is_fully_contiguous()


def test_fc_memory_region_kat():
    # fully_contiguous.rs:382-424: MockMemory(0x1000), nb=2 nl=2 no=2 page=16 inner=128 dtype=2
    L = O.Layout(O.FC, 2, 2, 2, 16, 128, 2, bases=[0x1000])
    R = 16 * 128 * 2
    assert L.memory_region(0, 0, 0) == (0x1000, R)
    assert L.memory_region(0, 0, 1) == (0x1000 + R, R)
    assert L.memory_region(0, 1, 0) == (0x1000 + 2 * R, R)
    assert L.memory_region(1, 0, 0) == (0x1000 + 2 * 2 * R, R)


def test_lw_block_first_kat():
    # layer_separate.rs:394-430: per-layer MockMemory at 0x1000 + i*per_layer
    per_layer = 2 * 2 * 16 * 128 * 2
    L = O.Layout(O.LW, 2, 2, 2, 16, 128, 2, block_dim=O.BLOCK_IS_FIRST_DIM,
                 bases=[0x1000 + i * per_layer for i in range(2)])
    R = 16 * 128 * 2
    assert L.memory_region(0, 0, 0) == (0x1000, R)
    assert L.memory_region(0, 1, 0) == (0x1000 + per_layer, R)
    assert L.memory_region(0, 0, 1) == (0x1000 + R, R)
    assert not L.is_fully_contiguous()


def test_lw_strides_both_block_dims():
    # layer_separate.rs:173-184
    R = 16 * 128 * 2
    a = O.Layout(O.LW, 10, 4, 2, 16, 128, 2, block_dim=O.BLOCK_IS_SECOND_DIM, bases=[0] * 4)
    assert (a.block_stride, a.outer_stride) == (R, R * 10)
    b = O.Layout(O.LW, 10, 4, 2, 16, 128, 2, block_dim=O.BLOCK_IS_FIRST_DIM, bases=[0] * 4)
    assert (b.outer_stride, b.block_stride) == (R, R * 2)
    assert O.lib().oracle_required_allocation(a.c, 0) == 10 * 2 * R  # layer_separate.rs:292-296


def test_out_of_range_ids_error():
    # fully_contiguous.rs:231-251 / layer_separate.rs:221-241
    L = O.Layout(O.FC, 2, 2, 2, 16, 128, 2, bases=[0x1000])
    for bad in [(2, 0, 0), (0, 2, 0), (0, 0, 2)]:
        with pytest.raises(O.OracleError) as e:
            L.memory_region(*bad)
        assert e.value.code == O.ERR_RANGE


@pytest.mark.parametrize("kw", [
    dict(num_blocks=0), dict(num_layers=0), dict(outer_dim=0), dict(outer_dim=3),
    dict(page_size=0), dict(inner_dim=0), dict(dtype_width_bytes=1), dict(dtype_width_bytes=3),
    dict(dtype_width_bytes=16),
])
def test_config_validation(kw):
    # config.rs:16-44 (ranges) and :173-180 (dtype width must be a power of two in 2..=8)
    base = dict(num_blocks=2, num_layers=2, outer_dim=2, page_size=16, inner_dim=128, dtype_width_bytes=2)
    base.update(kw)
    with pytest.raises(O.OracleError) as e:
        O.Layout(O.FC, bases=[0x1000], **base)
    assert e.value.code == O.ERR_CONFIG


def test_fp8_extension_accepts_width_one():
    L = O.Layout(O.FC, 2, 2, 2, 16, 128, 1, bases=[0x1000], allow_fp8=True)
    assert L.region_size == 16 * 128

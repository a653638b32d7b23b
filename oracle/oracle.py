"""Python face of the CPU oracle (oracle/kvbm_oracle.c) + BLAKE3 block checksums.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from dynamo_b200/.

Restates (paths relative to /root/reference):
  * layouts            lib/kvbm-physical/src/layout/{config,fully_contiguous,layer_separate}.rs
  * memcpy transfer    lib/kvbm-physical/src/transfer/executor/memcpy.rs:30-165
  * validation         lib/kvbm-physical/src/transfer/validation.rs:55-225
  * fill patterns      lib/kvbm-physical/src/transfer/fill.rs:51-213
  * block checksums    lib/kvbm-physical/src/transfer/checksum.rs:91-158  (BLAKE3, hex digest)
  * K1/K2/K3 semantics lib/kvbm-kernels/cuda/tensor_kernels.cu:109-228,494-541
Parity: pinned by the reference's known-answer tests (tests/test_oracle_*.py); the fp8<->bf16
cast has no reference implementation -> parity unpinned, pinned against torch CPU fixtures.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkvbm_oracle.so")
MAX_LAYERS = 256

BLOCK_IS_FIRST_DIM = 0
BLOCK_IS_SECOND_DIM = 1
FC, LW = 0, 1

OK, ERR_CONFIG, ERR_RANGE, ERR_LENGTH_MISMATCH, ERR_DUP_DST, ERR_OVERLAP, ERR_INCOMPATIBLE, ERR_SIZE_MISMATCH = range(8)


def build(force: bool = False) -> str:
    """Compile the C restatement (and oracle/_ref when the reference tree is present)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "kvbm_oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE, "libkvbm_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


class _CLayout(C.Structure):
    _fields_ = [
        ("kind", C.c_int),
        ("block_dim", C.c_int),
        ("num_blocks", C.c_size_t),
        ("num_layers", C.c_size_t),
        ("outer_dim", C.c_size_t),
        ("page_size", C.c_size_t),
        ("inner_dim", C.c_size_t),
        ("dtype_width_bytes", C.c_size_t),
        ("region_size", C.c_size_t),
        ("block_stride", C.c_size_t),
        ("layer_stride", C.c_size_t),
        ("outer_stride", C.c_size_t),
        ("layer_base", C.c_size_t * MAX_LAYERS),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        P = C.POINTER
        sz = C.c_size_t
        L.oracle_layout_init_fc.argtypes = [P(_CLayout), sz, sz, sz, sz, sz, sz, sz]
        L.oracle_layout_init_fc_ext.argtypes = L.oracle_layout_init_fc.argtypes
        L.oracle_layout_init_lw.argtypes = [P(_CLayout), P(sz), C.c_int, sz, sz, sz, sz, sz, sz]
        L.oracle_layout_init_lw_ext.argtypes = L.oracle_layout_init_lw.argtypes
        L.oracle_required_bytes.argtypes = [P(_CLayout)]
        L.oracle_required_bytes.restype = sz
        L.oracle_bytes_per_block.argtypes = [P(_CLayout)]
        L.oracle_bytes_per_block.restype = sz
        L.oracle_required_allocation.argtypes = [P(_CLayout), sz]
        L.oracle_required_allocation.restype = sz
        L.oracle_memory_region.argtypes = [P(_CLayout), sz, sz, sz, P(sz), P(sz)]
        L.oracle_validate_block_transfer.argtypes = [P(sz), sz, P(sz), sz, P(_CLayout), P(_CLayout), C.c_int]
        L.oracle_can_use_whole_block.argtypes = [P(_CLayout), P(_CLayout), C.c_int, sz, sz]
        L.oracle_execute_memcpy_transfer.argtypes = [P(_CLayout), P(_CLayout), P(sz), P(sz), sz, C.c_int, sz, sz]
        L.oracle_execute_memcpy_transfer_mt.argtypes = L.oracle_execute_memcpy_transfer.argtypes + [C.c_int]
        L.oracle_fill_blocks.argtypes = [P(_CLayout), P(sz), sz, C.c_int]
        L.oracle_fill_layers.argtypes = [P(_CLayout), P(sz), sz, sz, sz, C.c_int]
        L.oracle_vectorized_copy.argtypes = [P(C.c_void_p), P(C.c_void_p), sz, sz]
        L.oracle_vectorized_copy.restype = None
        for f in (L.oracle_universal_from_block, L.oracle_block_from_universal):
            f.argtypes = [P(C.c_void_p), P(C.c_void_p), sz, sz, sz, sz, sz, sz, sz, C.c_int]
            f.restype = None
        L.oracle_e4m3_to_bf16.argtypes = [C.c_uint8]
        L.oracle_e4m3_to_bf16.restype = C.c_uint16
        L.oracle_bf16_to_e4m3_satfinite.argtypes = [C.c_uint16]
        L.oracle_bf16_to_e4m3_satfinite.restype = C.c_uint8
        L.oracle_cast_e4m3_to_bf16.argtypes = [C.c_void_p, C.c_void_p, sz]
        L.oracle_cast_e4m3_to_bf16.restype = None
        L.oracle_cast_bf16_to_e4m3.argtypes = [C.c_void_p, C.c_void_p, sz]
        L.oracle_cast_bf16_to_e4m3.restype = None
        L.oracle_execute_cast_transfer.argtypes = L.oracle_execute_memcpy_transfer.argtypes + [C.c_int]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code: int, what: str = ""):
        super().__init__(f"oracle error {code} {what}")
        self.code = code


def _ids(a: Iterable[int]):
    arr = np.ascontiguousarray(np.asarray(list(a) if not isinstance(a, np.ndarray) else a, dtype=np.uint64))
    return arr, arr.ctypes.data_as(C.POINTER(C.c_size_t))


class Layout:
    """A layout over host memory owned by numpy arrays (System storage), or over raw addresses."""

    def __init__(self, kind: int, num_blocks: int, num_layers: int, outer_dim: int, page_size: int,
                 inner_dim: int, dtype_width_bytes: int = 2, block_dim: int = BLOCK_IS_FIRST_DIM,
                 bases: Optional[Sequence[int]] = None, allow_fp8: bool = False, fill: int = 0):
        self.c = _CLayout()
        self.kind = kind
        self.buffers: list[np.ndarray] = []
        L = lib()
        region = page_size * inner_dim * dtype_width_bytes
        if kind == FC:
            if bases is None:
                buf = np.full(max(1, num_blocks * num_layers * outer_dim * region), fill, dtype=np.uint8)
                self.buffers = [buf]
                bases = [buf.ctypes.data]
            fn = L.oracle_layout_init_fc_ext if allow_fp8 else L.oracle_layout_init_fc
            rc = fn(C.byref(self.c), bases[0], num_blocks, num_layers, outer_dim, page_size, inner_dim,
                    dtype_width_bytes)
        else:
            if bases is None:
                self.buffers = [np.full(max(1, num_blocks * outer_dim * region), fill, dtype=np.uint8)
                                for _ in range(num_layers)]
                bases = [b.ctypes.data for b in self.buffers]
            arr = (C.c_size_t * max(1, len(bases)))(*bases)
            fn = L.oracle_layout_init_lw_ext if allow_fp8 else L.oracle_layout_init_lw
            rc = fn(C.byref(self.c), arr, block_dim, num_blocks, num_layers, outer_dim, page_size,
                    inner_dim, dtype_width_bytes)
        if rc:
            raise OracleError(rc, "layout config")

    # -- geometry -------------------------------------------------------------------------------
    @property
    def num_blocks(self): return self.c.num_blocks
    @property
    def num_layers(self): return self.c.num_layers
    @property
    def outer_dim(self): return self.c.outer_dim
    @property
    def region_size(self): return self.c.region_size
    @property
    def block_stride(self): return self.c.block_stride
    @property
    def outer_stride(self): return self.c.outer_stride
    @property
    def layer_stride(self): return self.c.layer_stride
    def required_bytes(self): return lib().oracle_required_bytes(C.byref(self.c))
    def bytes_per_block(self): return lib().oracle_bytes_per_block(C.byref(self.c))
    def is_fully_contiguous(self): return self.kind == FC

    def memory_region(self, block: int, layer: int, outer: int) -> tuple[int, int]:
        a, s = C.c_size_t(), C.c_size_t()
        rc = lib().oracle_memory_region(C.byref(self.c), block, layer, outer, C.byref(a), C.byref(s))
        if rc:
            raise OracleError(rc, "memory_region")
        return a.value, s.value

    def region_bytes(self, block: int, layer: int, outer: int) -> np.ndarray:
        a, s = self.memory_region(block, layer, outer)
        return np.ctypeslib.as_array((C.c_uint8 * s).from_address(a))

    # -- test helpers (fill.rs / checksum.rs) ------------------------------------------------------
    def fill_blocks(self, ids, pattern: int = -1):
        arr, p = _ids(ids)
        rc = lib().oracle_fill_blocks(C.byref(self.c), p, len(arr), pattern)
        if rc:
            raise OracleError(rc, "fill_blocks")

    def fill_layers(self, ids, layer_begin: int, layer_end: int, pattern: int = -1):
        arr, p = _ids(ids)
        rc = lib().oracle_fill_layers(C.byref(self.c), p, len(arr), layer_begin, layer_end, pattern)
        if rc:
            raise OracleError(rc, "fill_layers")

    def block_checksum(self, block: int, layer_range: Optional[range] = None) -> str:
        """compute_single_block_checksum (checksum.rs:91-158): BLAKE3 over regions, layer-major."""
        import blake3
        h = blake3.blake3()
        layers = layer_range if layer_range is not None else range(self.num_layers)
        if len(layers) and layers[-1] >= self.num_layers:
            raise OracleError(ERR_RANGE, "layer range")
        for l in layers:
            for o in range(self.outer_dim):
                h.update(self.region_bytes(block, l, o).tobytes())
        return h.hexdigest()

    def block_checksums(self, ids, layer_range: Optional[range] = None) -> dict[int, str]:
        return {int(b): self.block_checksum(int(b), layer_range) for b in ids}


def validate_block_transfer(src_ids, dst_ids, src: Layout, dst: Layout) -> int:
    sa, sp = _ids(src_ids)
    da, dp = _ids(dst_ids)
    return lib().oracle_validate_block_transfer(sp, len(sa), dp, len(da), C.byref(src.c), C.byref(dst.c),
                                                int(src is dst))


def can_use_whole_block_transfer(src: Layout, dst: Layout, layer_range: Optional[range]) -> bool:
    lb, le = (layer_range.start, layer_range.stop) if layer_range is not None else (0, 0)
    return bool(lib().oracle_can_use_whole_block(C.byref(src.c), C.byref(dst.c), int(layer_range is not None), lb, le))


def execute_memcpy_transfer(src: Layout, dst: Layout, src_ids, dst_ids, layer_range: Optional[range] = None,
                            nthreads: int = 1, cast_mode: int = 0) -> None:
    sa, sp = _ids(src_ids)
    da, dp = _ids(dst_ids)
    if len(sa) != len(da):
        raise OracleError(ERR_LENGTH_MISMATCH, "ids")  # memcpy.rs:38-44
    lb, le = (layer_range.start, layer_range.stop) if layer_range is not None else (0, 0)
    has = int(layer_range is not None)
    if cast_mode:
        rc = lib().oracle_execute_cast_transfer(C.byref(src.c), C.byref(dst.c), sp, dp, len(sa), has, lb, le, cast_mode)
    elif nthreads > 1:
        rc = lib().oracle_execute_memcpy_transfer_mt(C.byref(src.c), C.byref(dst.c), sp, dp, len(sa), has, lb, le, nthreads)
    else:
        rc = lib().oracle_execute_memcpy_transfer(C.byref(src.c), C.byref(dst.c), sp, dp, len(sa), has, lb, le)
    if rc:
        raise OracleError(rc, "execute_memcpy_transfer")


def vectorized_copy(src_bufs: Sequence[np.ndarray], dst_bufs: Sequence[np.ndarray], copy_size: int) -> None:
    n = len(src_bufs)
    s = (C.c_void_p * n)(*[b.ctypes.data for b in src_bufs])
    d = (C.c_void_p * n)(*[b.ctypes.data for b in dst_bufs])
    lib().oracle_vectorized_copy(s, d, copy_size, n)


def universal_from_block(universals: Sequence[np.ndarray], chunks: Sequence[np.ndarray], nh, nl, no, nt, hd,
                         elem: int, layout: int) -> None:
    u = (C.c_void_p * len(universals))(*[b.ctypes.data for b in universals])
    c = (C.c_void_p * len(chunks))(*[b.ctypes.data for b in chunks])
    lib().oracle_universal_from_block(u, c, len(universals), nh, nl, no, nt, hd, elem, layout)


def block_from_universal(universals: Sequence[np.ndarray], chunks: Sequence[np.ndarray], nh, nl, no, nt, hd,
                         elem: int, layout: int) -> None:
    u = (C.c_void_p * len(universals))(*[b.ctypes.data for b in universals])
    c = (C.c_void_p * len(chunks))(*[b.ctypes.data for b in chunks])
    lib().oracle_block_from_universal(u, c, len(universals), nh, nl, no, nt, hd, elem, layout)


def cast_e4m3_to_bf16(src: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(src, dtype=np.uint8)
    out = np.empty(src.shape, dtype=np.uint16)
    lib().oracle_cast_e4m3_to_bf16(src.ctypes.data, out.ctypes.data, src.size)
    return out


def cast_bf16_to_e4m3(src: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(src, dtype=np.uint16)
    out = np.empty(src.shape, dtype=np.uint8)
    lib().oracle_cast_bf16_to_e4m3(src.ctypes.data, out.ctypes.data, src.size)
    return out


# ----------------------------------------------------------------------------------------------------------------------
# KvBlockLayout transforms -- lib/kvbm-physical/src/layout/kv_block_layout.rs:85-95 (dim_order), the definition the
# reference's select_transform_kernel (transfer/executor/mod.rs:46-100) pairs kernels with.  The reference only HAS kernels
# for Operational{NHD,HND} <-> UniversalTP (K2 / K3, restated in kvbm_oracle.c and pinned there); for those pairs
# `kv_layout_permute` is checked against that restatement in tests/test_oracle_kernels.py.  The other pairs (UniversalPP,
# NHD <-> HND) have no reference kernel: their oracle is the dim_order definition itself -- parity unpinned beyond it.
# ----------------------------------------------------------------------------------------------------------------------
KV_UNKNOWN, KV_UNIVERSAL_TP, KV_UNIVERSAL_PP, KV_OPERATIONAL_HND, KV_OPERATIONAL_NHD = 0, 1, 2, 3, 4
KV_DIM_ORDER = {KV_UNIVERSAL_TP: "hlot", KV_UNIVERSAL_PP: "lhot", KV_OPERATIONAL_HND: "loht", KV_OPERATIONAL_NHD: "loth"}


def kv_layout_permute(block: np.ndarray, src_kv: int, dst_kv: int, nl: int, no: int, nt: int, nh: int, row_bytes: int,
                      layers: Optional[range] = None, dst_old: Optional[np.ndarray] = None) -> np.ndarray:
    """One contiguous block (uint8[nl*no*nt*nh*row_bytes]) re-ordered from `src_kv` to `dst_kv`.  With `layers`, only
    elements of those layers move; every other byte keeps `dst_old`'s value (a partial, layer-wise transfer)."""
    size = {"h": nh, "l": nl, "o": no, "t": nt}
    so, do = KV_DIM_ORDER[src_kv], KV_DIM_ORDER[dst_kv]
    a = np.ascontiguousarray(block, dtype=np.uint8).reshape([size[c] for c in so] + [row_bytes])
    perm = [so.index(c) for c in do] + [4]
    out = np.ascontiguousarray(a.transpose(perm))
    if layers is not None:
        keep = np.ascontiguousarray(dst_old, dtype=np.uint8).reshape(out.shape).copy()
        sel = [slice(None)] * 5
        sel[do.index("l")] = slice(layers.start, layers.stop)
        keep[tuple(sel)] = out[tuple(sel)]
        out = keep
    return out.reshape(-1)


def read_logical_block(layout: "Layout", block: int) -> np.ndarray:
    """The (layer, outer) regions of one block concatenated layer-major: for a fully contiguous layout this IS the block's
    memory (so also the right view of a universal-format block); for a layer-separate one it is the operational block
    [nl, no, region] the regions spell."""
    return np.concatenate([layout.region_bytes(block, l, o) for l in range(layout.num_layers) for o in range(layout.outer_dim)])


def write_logical_block(layout: "Layout", block: int, data: np.ndarray) -> None:
    r = layout.region_size
    k = 0
    for l in range(layout.num_layers):
        for o in range(layout.outer_dim):
            layout.region_bytes(block, l, o)[:] = data[k * r:(k + 1) * r]
            k += 1

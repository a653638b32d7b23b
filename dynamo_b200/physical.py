"""ctypes mirror of the reference's host-side transfer API (lib/kvbm-physical), over libkvbm_physical.so.

Names and argument meaning follow the Rust crate so tests read like the reference's own:
  LayoutConfig                    lib/kvbm-physical/src/layout/config.rs:14-56
  StorageKind / BlockDimension    lib/memory/src/lib.rs, layout/config.rs:151-163
  TransferOptions                 transfer/options.rs:27-81
  TransferManager.register_layout / export_metadata / import_metadata / execute_transfer
                                  manager/mod.rs:101,112,130,227
  select_direct_strategy          transfer/strategy.rs:138-210
  validate_block_transfer         transfer/validation.rs:168-225
  TransferCompleteNotification    transfer/context.rs:438-470

All bytes move in native code (CUDA kernels for device storage, the reference's own memcpy strategy for
host<->host); this module only marshals arguments.  Missing native library => exception, never a fallback.
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass
from typing import List, Optional, Sequence

from . import _lib, kernels


class KvbmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{ErrorCode(code).name if code in ErrorCode._value2member_map_ else code}] {msg}")
        self.code = code
        self.msg = msg


class ErrorCode(enum.IntEnum):
    OK = 0
    ERR = 1
    CONFIG = 2
    RANGE = 3
    LENGTH_MISMATCH = 4
    DUPLICATE_DST = 5
    OVERLAP = 6
    INCOMPATIBLE = 7
    UNSUPPORTED = 8
    CUDA = 9
    HANDLE = 10
    TIMEOUT = 11
    VERSION = 12


class StorageKind(enum.IntEnum):
    System = 0
    Pinned = 1
    Device = 2
    Disk = 3


class BlockDimension(enum.IntEnum):
    BlockIsFirstDim = 0
    BlockIsSecondDim = 1


class TransferStrategy(enum.IntEnum):
    Memcpy = 0
    CudaAsyncH2D = 1
    CudaAsyncD2H = 2
    CudaAsyncD2D = 3
    NixlRead = 4
    NixlWrite = 5
    NixlReadFlipped = 6
    Invalid = 7


class _CConfig(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("num_blocks", "num_layers", "outer_dim", "page_size", "inner_dim",
                                          "alignment", "dtype_width_bytes", "num_heads")] + [("allow_fp8", C.c_int)]


class _CPlan(C.Structure):
    _fields_ = [("two_hop", C.c_int), ("first", C.c_int), ("bounce_location", C.c_int), ("second", C.c_int)]


class _CCaps(C.Structure):
    _fields_ = [("allow_gds", C.c_int), ("allow_gpu_rdma", C.c_int)]


class _COptions(C.Structure):
    _fields_ = [
        ("has_layer_range", C.c_int),
        ("layer_begin", C.c_size_t),
        ("layer_end", C.c_size_t),
        ("cuda_stream", C.c_void_p),
        ("use_caller_stream", C.c_int),
        ("cast_mode", C.c_int),
        ("max_ctas", C.c_int),
        ("layer_ready_flags", C.c_void_p),
        ("layer_done_flags", C.c_void_p),
        ("epoch", C.c_uint32),
        ("gate_timeout_ms", C.c_int),
        ("multicast", C.c_int),
        ("done_flag", C.c_void_p),
        ("bounce_layout", C.c_uint64),
        ("bounce_block_ids", C.c_void_p),
        ("num_bounce_blocks", C.c_size_t),
        ("src_kv_layout", C.c_int),
        ("dst_kv_layout", C.c_int),
        ("gate_mode", C.c_int),
        ("per_dst_done_flags", C.c_void_p),
        ("per_dst_layer_done_flags", C.c_void_p),
    ]


@dataclass
class LayoutConfig:
    num_blocks: int
    num_layers: int
    outer_dim: int
    page_size: int
    inner_dim: int
    alignment: int = 1
    dtype_width_bytes: int = 2
    num_heads: Optional[int] = None
    allow_fp8: bool = False

    def _c(self) -> _CConfig:
        return _CConfig(self.num_blocks, self.num_layers, self.outer_dim, self.page_size, self.inner_dim,
                        self.alignment, self.dtype_width_bytes, self.num_heads or 0, int(self.allow_fp8))

    def validate(self) -> None:
        _check(lib().kvbm_layout_config_validate(C.byref(self._c())))

    def required_bytes(self) -> int:
        return lib().kvbm_layout_required_bytes(C.byref(self._c()))

    def bytes_per_block(self) -> int:
        return lib().kvbm_layout_bytes_per_block(C.byref(self._c()))

    def region_size(self) -> int:
        return self.page_size * self.inner_dim * self.dtype_width_bytes


@dataclass
class TransferOptions:
    layer_range: Optional[range] = None
    cuda_stream: Optional[int] = None     # raw cudaStream_t; when set the caller manages synchronisation
    cast_mode: int = 0
    max_ctas: int = 0
    layer_ready_flags: int = 0
    layer_done_flags: int = 0
    epoch: int = 0
    gate_timeout_ms: int = 0
    multicast: int = 0                    # destination layout lives in a MulticastGroup.map() range (NVLS)
    done_flag: int = 0                    # word on the destination GPU that receives `epoch` on completion (cf. nixl_write_notification)
    bounce_buffer: Optional[tuple] = None # (layout handle, block ids): BounceBuffer of options.rs:45-51 for two-hop transfers
    src_kv_layout: int = 0                # KvBlockLayout overrides (options.rs:63-80); a pair needing a transform runs the permuting launch
    dst_kv_layout: int = 0
    gate_mode: int = 0                    # kernels.GATE_AUTO / GATE_SPIN / GATE_STREAM_WAIT
    per_dst_done_flags: Optional[Sequence[int]] = None        # fan-out: one done-flag address per destination (0 = none)
    per_dst_layer_done_flags: Optional[Sequence[int]] = None  # fan-out: one layer-done array address per destination

    @staticmethod
    def from_layer_range(layer_range: Optional[range]) -> "TransferOptions":
        return TransferOptions(layer_range=layer_range)

    def _c(self) -> _COptions:
        lr = self.layer_range
        bl, bids, nb = 0, None, 0
        if self.bounce_buffer is not None:
            bl, ids = self.bounce_buffer
            self._bounce_keepalive = (C.c_size_t * max(1, len(ids)))(*[int(x) for x in ids])
            bids, nb = C.cast(self._bounce_keepalive, C.c_void_p), len(ids)
        return _COptions(int(lr is not None), lr.start if lr is not None else 0, lr.stop if lr is not None else 0,
                         self.cuda_stream or 0, int(self.cuda_stream is not None), int(self.cast_mode), self.max_ctas,
                         self.layer_ready_flags, self.layer_done_flags, self.epoch, self.gate_timeout_ms,
                         int(self.multicast), self.done_flag, int(bl), bids, nb, int(self.src_kv_layout),
                         int(self.dst_kv_layout), int(self.gate_mode), self._ptr_array("_pd", self.per_dst_done_flags),
                         self._ptr_array("_pl", self.per_dst_layer_done_flags))

    def _ptr_array(self, keep: str, ptrs):
        if ptrs is None:
            return None
        arr = (C.c_void_p * max(1, len(ptrs)))(*[int(p) or None for p in ptrs])
        setattr(self, keep, arr)          # keep the array alive as long as the options object
        return C.cast(arr, C.c_void_p)


@dataclass
class TransferPlan:
    two_hop: bool
    first: TransferStrategy
    bounce_location: Optional[StorageKind] = None
    second: Optional[TransferStrategy] = None


_configured = False

EXPORTED_SYMBOLS = [
    "kvbm_last_error", "kvbm_layout_config_validate", "kvbm_layout_required_bytes", "kvbm_layout_bytes_per_block",
    "kvbm_select_direct_strategy", "kvbm_validate_block_transfer", "kvbm_manager_create", "kvbm_manager_destroy",
    "kvbm_manager_register_fully_contiguous", "kvbm_manager_register_layer_separate", "kvbm_manager_unregister",
    "kvbm_layout_memory_region", "kvbm_layout_is_fully_contiguous", "kvbm_manager_enable_peer_access",
    "kvbm_manager_export_metadata", "kvbm_manager_import_metadata", "kvbm_manager_import_metadata_mapped", "kvbm_manager_execute_transfer",
    "kvbm_manager_execute_fanout", "kvbm_notification_is_complete", "kvbm_notification_wait",
    "kvbm_manager_bytes_moved", "kvbm_manager_h2d_bytes", "kvbm_manager_set_capabilities",
    "kvbm_manager_export_serialized_layout", "kvbm_manager_import_serialized_layout", "kvbm_layout_descriptor_json",
    "kvbm_manager_import_descriptor_json", "kvbm_select_transform_kernel", "kvbm_kv_layout_requires_transform",
    "kvbm_manager_set_kv_block_layout", "kvbm_manager_kv_block_layout",
    "kvbm_select_direct_strategy_remote", "kvbm_select_strategy", "kvbm_manager_select_strategy",
    "kvbm_mc_supported", "kvbm_mc_group_create", "kvbm_mc_group_export_fd", "kvbm_mc_group_import_fd",
    "kvbm_mc_group_size", "kvbm_mc_group_add_device", "kvbm_mc_group_bind_local", "kvbm_mc_group_bind_addr", "kvbm_mc_group_map",
    "kvbm_mc_group_destroy",
]


_LIB: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _configured, _LIB
    if _LIB is not None:      # hot path: every transfer goes through here twice
        return _LIB
    kernels.lib()  # dependency (resolved through $ORIGIN rpath as well)
    L = _lib.load(_lib.PHYSICAL_SO)
    if not _configured:
        vp, sz, i, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
        P = C.POINTER
        L.kvbm_last_error.restype = C.c_char_p
        L.kvbm_layout_config_validate.argtypes = [P(_CConfig)]
        L.kvbm_layout_required_bytes.argtypes = [P(_CConfig)]
        L.kvbm_layout_required_bytes.restype = sz
        L.kvbm_layout_bytes_per_block.argtypes = [P(_CConfig)]
        L.kvbm_layout_bytes_per_block.restype = sz
        L.kvbm_select_direct_strategy.argtypes = [i, i, P(_CCaps), P(_CPlan)]
        L.kvbm_select_direct_strategy_remote.argtypes = [i, i, i, P(_CCaps), P(_CPlan)]
        L.kvbm_select_strategy.argtypes = [i, i, i, i, P(_CCaps), P(_CPlan)]
        L.kvbm_manager_select_strategy.argtypes = [vp, u64, u64, P(_CPlan)]
        L.kvbm_validate_block_transfer.argtypes = [P(sz), sz, P(sz), sz, sz, sz, i]
        L.kvbm_manager_create.argtypes = [i, u64, P(vp)]
        L.kvbm_manager_destroy.argtypes = [vp]
        L.kvbm_manager_destroy.restype = None
        L.kvbm_manager_register_fully_contiguous.argtypes = [vp, P(_CConfig), vp, sz, i, i, P(u64)]
        L.kvbm_manager_register_layer_separate.argtypes = [vp, P(_CConfig), P(vp), P(sz), i, i, i, P(u64)]
        L.kvbm_manager_unregister.argtypes = [vp, u64]
        L.kvbm_layout_memory_region.argtypes = [vp, u64, sz, sz, sz, P(sz), P(sz)]
        L.kvbm_layout_is_fully_contiguous.argtypes = [vp, u64]
        L.kvbm_manager_enable_peer_access.argtypes = [vp, i]
        L.kvbm_manager_export_metadata.argtypes = [vp, u64, vp, sz, P(sz)]
        L.kvbm_manager_import_metadata.argtypes = [vp, vp, sz, P(u64)]
        L.kvbm_manager_import_metadata_mapped.argtypes = [vp, vp, sz, P(vp), sz, P(u64)]
        # the id lists are `const size_t*`; declared void* so a raw address (numpy buffer) passes without a ctypes cast
        L.kvbm_manager_execute_transfer.argtypes = [vp, u64, vp, u64, vp, sz, P(_COptions), P(u64)]
        L.kvbm_manager_execute_fanout.argtypes = [vp, u64, i, P(u64), P(P(sz)), P(P(sz)), sz, i, P(_COptions), P(u64)]
        L.kvbm_manager_set_capabilities.argtypes = [vp, P(_CCaps)]
        L.kvbm_manager_set_kv_block_layout.argtypes = [vp, u64, i]
        L.kvbm_manager_kv_block_layout.argtypes = [vp, u64]
        L.kvbm_select_transform_kernel.argtypes = [i, i]
        L.kvbm_kv_layout_requires_transform.argtypes = [i, i]
        L.kvbm_notification_is_complete.argtypes = [vp, u64]
        L.kvbm_notification_wait.argtypes = [vp, u64, C.c_int64]
        L.kvbm_manager_bytes_moved.argtypes = [vp]
        L.kvbm_manager_bytes_moved.restype = u64
        L.kvbm_manager_h2d_bytes.argtypes = [vp]
        L.kvbm_manager_h2d_bytes.restype = u64
        L.kvbm_manager_export_serialized_layout.argtypes = [vp, vp, sz, P(sz)]
        L.kvbm_manager_import_serialized_layout.argtypes = [vp, vp, sz, P(u64), sz, P(sz)]
        L.kvbm_layout_descriptor_json.argtypes = [vp, u64, vp, sz, P(sz)]
        L.kvbm_manager_import_descriptor_json.argtypes = [vp, C.c_char_p, sz, P(u64)]
        L.kvbm_mc_supported.argtypes = [i]
        L.kvbm_mc_group_create.argtypes = [i, sz, i, P(vp)]
        L.kvbm_mc_group_export_fd.argtypes = [vp, P(i)]
        L.kvbm_mc_group_import_fd.argtypes = [i, i, sz, P(vp)]
        L.kvbm_mc_group_size.argtypes = [vp]
        L.kvbm_mc_group_size.restype = sz
        L.kvbm_mc_group_add_device.argtypes = [vp, i]
        L.kvbm_mc_group_bind_local.argtypes = [vp, i, P(vp)]
        L.kvbm_mc_group_bind_addr.argtypes = [vp, i, vp, sz]
        L.kvbm_mc_group_map.argtypes = [vp, i, P(vp)]
        L.kvbm_mc_group_destroy.argtypes = [vp]
        L.kvbm_mc_group_destroy.restype = None
        _configured = True
    _LIB = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise KvbmError(rc, lib().kvbm_last_error().decode(errors="replace"))


try:
    import numpy as _np
    _NP_ID_DTYPES = (_np.dtype(_np.uint64), _np.dtype(_np.int64))
except ImportError:  # pragma: no cover
    _np = None
    _NP_ID_DTYPES = ()


def _size_array(ids: Sequence[int]):
    """Host block-id list -> `const size_t*`.  numpy int64/uint64 arrays are passed without copying."""
    n = len(ids)
    if _np is not None and isinstance(ids, _np.ndarray) and ids.dtype in _NP_ID_DTYPES and ids.flags.c_contiguous:
        return _NpView(ids), n
    return (C.c_size_t * max(1, n))(*[int(x) for x in ids]), n


class _NpView:
    """Keeps the numpy array alive and presents its buffer as a pointer argument."""
    __slots__ = ("arr", "addr")

    def __init__(self, arr):
        self.arr = arr
        self.addr = arr.__array_interface__["data"][0]

    @property
    def _as_parameter_(self):     # only the fan-out / validation paths need a typed pointer
        return C.cast(self.addr, C.POINTER(C.c_size_t))


def _plan(plan: _CPlan) -> TransferPlan:
    if plan.two_hop:
        return TransferPlan(True, TransferStrategy(plan.first), StorageKind(plan.bounce_location),
                            TransferStrategy(plan.second))
    return TransferPlan(False, TransferStrategy(plan.first))


def select_direct_strategy(src: StorageKind, dst: StorageKind, allow_gds: bool = False,
                           allow_gpu_rdma: bool = False, dst_is_remote: bool = False) -> TransferPlan:
    """select_direct_strategy(src, dst, dst_is_remote, capabilities) -- transfer/strategy.rs:138-243."""
    plan = _CPlan()
    caps = _CCaps(int(allow_gds), int(allow_gpu_rdma))
    _check(lib().kvbm_select_direct_strategy_remote(int(src), int(dst), int(dst_is_remote), C.byref(caps), C.byref(plan)))
    return _plan(plan)


def select_strategy(src: StorageKind, src_is_local: bool, dst: StorageKind, dst_is_local: bool, allow_gds: bool = False,
                    allow_gpu_rdma: bool = False) -> TransferPlan:
    """select_strategy + select_remote_strategy_v2 -- transfer/strategy.rs:78-108,245-281."""
    plan = _CPlan()
    caps = _CCaps(int(allow_gds), int(allow_gpu_rdma))
    _check(lib().kvbm_select_strategy(int(src), int(src_is_local), int(dst), int(dst_is_local), C.byref(caps), C.byref(plan)))
    return _plan(plan)


class TransformKernel(enum.IntEnum):
    """TransformKernel (transfer/executor/mod.rs:27-41)"""
    NONE = 0
    BlockToUniversal = 1
    UniversalToBlock = 2
    OperationalTranspose = 3
    Unsupported = 4
    UniversalToUniversal = 5   # extension (a TODO in the reference); executed by transfers, never selected by the function below


def select_transform_kernel(src: "kernels.KvBlockLayout", dst: "kernels.KvBlockLayout") -> TransformKernel:
    """select_transform_kernel (transfer/executor/mod.rs:46-100)."""
    return TransformKernel(lib().kvbm_select_transform_kernel(int(src), int(dst)))


def requires_transform(a: "kernels.KvBlockLayout", b: "kernels.KvBlockLayout") -> bool:
    """KvBlockLayout::requires_transform (layout/kv_block_layout.rs:107-119)."""
    return bool(lib().kvbm_kv_layout_requires_transform(int(a), int(b)))


def validate_block_transfer(src_ids: Sequence[int], dst_ids: Sequence[int], src_num_blocks: int,
                            dst_num_blocks: int, same_layout: bool = False) -> None:
    s, ns = _size_array(src_ids)
    d, nd = _size_array(dst_ids)
    _check(lib().kvbm_validate_block_transfer(s, ns, d, nd, src_num_blocks, dst_num_blocks, int(same_layout)))


class TransferCompleteNotification:
    def __init__(self, mgr: "TransferManager", token: int):
        self._mgr, self.token = mgr, token

    def is_complete(self) -> bool:
        rc = lib().kvbm_notification_is_complete(self._mgr._h, self.token)
        if rc == -2:
            raise KvbmError(ErrorCode.TIMEOUT, "gated transfer aborted: a layer_ready flag was not released within the gate timeout")
        if rc < 0:
            raise KvbmError(ErrorCode.HANDLE, "unknown notification")
        return rc == 1

    def wait(self, timeout_s: Optional[float] = 30.0) -> None:
        us = -1 if timeout_s is None else int(timeout_s * 1e6)
        _check(lib().kvbm_notification_wait(self._mgr._h, self.token, us))


class TransferManager:
    """TransferManager over one CUDA device (device < 0: host-only, Memcpy strategy only)."""

    def __init__(self, device: int = 0, worker_id: int = 0):
        h = C.c_void_p()
        _check(lib().kvbm_manager_create(device, worker_id, C.byref(h)))
        self._h = h
        self.device = device

    def close(self) -> None:
        if self._h:
            lib().kvbm_manager_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- layouts ------------------------------------------------------------------------------------
    def register_fully_contiguous(self, config: LayoutConfig, base: int, size: int, storage: StorageKind,
                                  device_id: int = 0) -> int:
        out = C.c_uint64()
        _check(lib().kvbm_manager_register_fully_contiguous(self._h, C.byref(config._c()), base, size, int(storage),
                                                            device_id, C.byref(out)))
        return out.value

    def register_layer_separate(self, config: LayoutConfig, layer_bases: Sequence[int], layer_sizes: Sequence[int],
                                block_dim: BlockDimension, storage: StorageKind, device_id: int = 0) -> int:
        n = len(layer_bases)
        if n != config.num_layers or len(layer_sizes) != n:
            raise KvbmError(ErrorCode.CONFIG, f"Memory region count ({n}) must match num_layers ({config.num_layers})")
        bases = (C.c_void_p * max(1, n))(*[int(b) for b in layer_bases])
        sizes = (C.c_size_t * max(1, n))(*[int(s) for s in layer_sizes])
        out = C.c_uint64()
        _check(lib().kvbm_manager_register_layer_separate(self._h, C.byref(config._c()), bases, sizes, int(block_dim),
                                                          int(storage), device_id, C.byref(out)))
        return out.value

    def unregister(self, handle: int) -> None:
        _check(lib().kvbm_manager_unregister(self._h, handle))

    def memory_region(self, handle: int, block: int, layer: int, outer: int) -> tuple[int, int]:
        a, s = C.c_size_t(), C.c_size_t()
        _check(lib().kvbm_layout_memory_region(self._h, handle, block, layer, outer, C.byref(a), C.byref(s)))
        return a.value, s.value

    def is_fully_contiguous(self, handle: int) -> bool:
        rc = lib().kvbm_layout_is_fully_contiguous(self._h, handle)
        if rc < 0:
            raise KvbmError(ErrorCode.HANDLE, "invalid handle")
        return bool(rc)

    def select_strategy(self, src: int, dst: int) -> TransferPlan:
        """The plan select_strategy (strategy.rs:78-108) gives for two of this manager's layouts."""
        plan = _CPlan()
        _check(lib().kvbm_manager_select_strategy(self._h, src, dst, C.byref(plan)))
        return _plan(plan)

    def set_kv_block_layout(self, handle: int, kv_layout: "kernels.KvBlockLayout") -> None:
        """The format of one block of this layout (builder `.kv_block_layout()` / `.inner_shape()` in the reference).  A transfer
        between layouts whose formats differ is executed as ONE permuting launch (the reference rejects it)."""
        _check(lib().kvbm_manager_set_kv_block_layout(self._h, handle, int(kv_layout)))

    def kv_block_layout(self, handle: int) -> "kernels.KvBlockLayout":
        rc = lib().kvbm_manager_kv_block_layout(self._h, handle)
        if rc < 0:
            raise KvbmError(ErrorCode.HANDLE, "invalid handle")
        return kernels.KvBlockLayout(rc)

    def enable_peer_access(self, peer_device: int) -> None:
        _check(lib().kvbm_manager_enable_peer_access(self._h, peer_device))

    def export_metadata(self, handle: int) -> bytes:
        n = C.c_size_t()
        _check(lib().kvbm_manager_export_metadata(self._h, handle, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _check(lib().kvbm_manager_export_metadata(self._h, handle, buf, n.value, C.byref(n)))
        return buf.raw[:n.value]

    def import_metadata(self, blob: bytes, local_bases: Optional[Sequence[int]] = None) -> int:
        """`local_bases`: this process's own mappings of another process's HOST pool (shared memory), one address per
        allocation; without it a foreign host pool is imported as a descriptor only."""
        out = C.c_uint64()
        buf = C.create_string_buffer(blob, len(blob))
        if local_bases is None:
            _check(lib().kvbm_manager_import_metadata(self._h, buf, len(blob), C.byref(out)))
        else:
            arr = (C.c_void_p * max(1, len(local_bases)))(*[int(b) for b in local_bases])
            _check(lib().kvbm_manager_import_metadata_mapped(self._h, buf, len(blob), arr, len(local_bases), C.byref(out)))
        return out.value

    # -- the reference's wire formats ----------------------------------------------
    def export_serialized_layout(self) -> bytes:
        """`TransferManager::export_metadata()` (manager/mod.rs:112): every local host/device layout in ONE
        `SerializedLayout` blob (bincode 2 over `RdmaLayoutDescriptors`, manager/metadata.rs:87-134)."""
        n = C.c_size_t()
        _check(lib().kvbm_manager_export_serialized_layout(self._h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(max(1, n.value))
        _check(lib().kvbm_manager_export_serialized_layout(self._h, buf, n.value, C.byref(n)))
        return buf.raw[:n.value]

    def import_serialized_layout(self, blob: bytes) -> List[int]:
        """`TransferManager::import_metadata()` (manager/mod.rs:130): handles of the imported remote layouts."""
        buf = C.create_string_buffer(blob, len(blob))
        n = C.c_size_t()
        _check(lib().kvbm_manager_import_serialized_layout(self._h, buf, len(blob), None, 0, C.byref(n)))
        out = (C.c_uint64 * max(1, n.value))()
        _check(lib().kvbm_manager_import_serialized_layout(self._h, buf, len(blob), out, n.value, C.byref(n)))
        return [int(out[i]) for i in range(n.value)]

    def layout_descriptor_json(self, handle: int) -> str:
        """`LayoutDescriptor::to_json` (layout/serialize.rs:96-103) of a registered layout."""
        n = C.c_size_t()
        _check(lib().kvbm_layout_descriptor_json(self._h, handle, None, 0, C.byref(n)))
        buf = C.create_string_buffer(max(1, n.value))
        _check(lib().kvbm_layout_descriptor_json(self._h, handle, buf, n.value, C.byref(n)))
        return buf.raw[:n.value].decode()

    def import_descriptor_json(self, text: str) -> int:
        """`LayoutDescriptor::from_json` + `PhysicalLayout::from_descriptor` (layout/physical.rs:203-262)."""
        raw = text.encode()
        out = C.c_uint64()
        _check(lib().kvbm_manager_import_descriptor_json(self._h, raw, len(raw), C.byref(out)))
        return out.value

    # -- transfers ----------------------------------------------------------------------------------
    def execute_transfer(self, src: int, src_block_ids: Sequence[int], dst: int, dst_block_ids: Sequence[int],
                         options: Optional[TransferOptions] = None) -> TransferCompleteNotification:
        if len(src_block_ids) != len(dst_block_ids):
            # executor/memcpy.rs:38-44: length mismatch is rejected before anything else
            raise KvbmError(ErrorCode.LENGTH_MISMATCH,
                            f"Block ID lists have mismatched lengths: src={len(src_block_ids)}, dst={len(dst_block_ids)}, bounce=None")
        s, n = _size_array(src_block_ids)
        d, _ = _size_array(dst_block_ids)
        tok = C.c_uint64()
        o = options._c() if options is not None else _COptions()
        rc = lib().kvbm_manager_execute_transfer(self._h, src, s.addr if type(s) is _NpView else s, dst,
                                                 d.addr if type(d) is _NpView else d, n, C.byref(o), C.byref(tok))
        if rc:
            _check(rc)
        return TransferCompleteNotification(self, tok.value)

    def execute_fanout(self, src: int, dsts: Sequence[int], src_block_ids: Sequence[Sequence[int]],
                       dst_block_ids: Sequence[Sequence[int]], replicate: bool = False,
                       options: Optional[TransferOptions] = None) -> TransferCompleteNotification:
        nd = len(dsts)
        n = len(dst_block_ids[0])
        if any(len(x) != n for x in dst_block_ids) or any(len(x) != n for x in src_block_ids):
            raise KvbmError(ErrorCode.LENGTH_MISMATCH, "Block ID lists have mismatched lengths")
        hs = (C.c_uint64 * nd)(*dsts)
        s_arrays = [_size_array(x)[0] for x in (src_block_ids if not replicate else [src_block_ids[0]] * nd)]
        d_arrays = [_size_array(x)[0] for x in dst_block_ids]
        PP = C.POINTER(C.c_size_t)
        as_ptr = lambda a: a._as_parameter_ if isinstance(a, _NpView) else C.cast(a, PP)  # noqa: E731
        sp = (PP * nd)(*[as_ptr(a) for a in s_arrays])
        dp = (PP * nd)(*[as_ptr(a) for a in d_arrays])
        tok = C.c_uint64()
        o = (options or TransferOptions())._c()
        _check(lib().kvbm_manager_execute_fanout(self._h, src, nd, hs, sp, dp, n, int(replicate), C.byref(o), C.byref(tok)))
        return TransferCompleteNotification(self, tok.value)

    def broadcast(self, src: int, dsts: Sequence[int], src_block_ids: Sequence[int],
                  dst_block_ids: Sequence[int], layer_range: Optional[range] = None) -> TransferCompleteNotification:
        """CollectiveOps::broadcast (lib/kvbm-engine/src/collectives/mod.rs:99-106): same blocks, same destination
        ids on every rank -- here one replicate launch over peer mappings instead of grouped ncclBcast."""
        return self.execute_fanout(src, dsts, [src_block_ids] * len(dsts), [dst_block_ids] * len(dsts), replicate=True,
                                   options=TransferOptions(layer_range=layer_range))

    def set_capabilities(self, allow_gds: bool = False, allow_gpu_rdma: bool = True) -> None:
        """TransferCapabilities (strategy.rs:245-278); allow_gpu_rdma=False turns cross-GPU transfers into the reference's
        TwoHop plan through `TransferOptions.bounce_buffer`."""
        caps = _CCaps(int(allow_gds), int(allow_gpu_rdma))
        _check(lib().kvbm_manager_set_capabilities(self._h, C.byref(caps)))

    def bytes_moved(self) -> int:
        return lib().kvbm_manager_bytes_moved(self._h)

    def h2d_bytes(self) -> int:
        return lib().kvbm_manager_h2d_bytes(self._h)


def multicast_supported(device: int = 0) -> bool:
    """CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED (NVSwitch multicast objects usable from this process)."""
    return bool(lib().kvbm_mc_supported(device))


class MulticastGroup:
    """NVLS multicast group (`kvbm_mc_group_*`, include/kvbm_physical.h): the receivers' pools are bound to one
    multicast object; the sender maps it and transfers with `TransferOptions(multicast=1)`, so identical KV blocks reach
    every bound GPU with ONE write per tile -- the replacement of the grouped `ncclBcast` per region
    (lib/kvbm-engine/src/collectives/nccl.rs:421-462).

    Call order across all participants: create()/from_fd() -> add_device() for every device (by its owning process)
    -> barrier -> bind_local() on every receiver -> barrier -> map() on the sender."""

    def __init__(self, handle: int, num_devices: int):
        self._h = C.c_void_p(handle)
        self.num_devices = num_devices

    @staticmethod
    def create(num_devices: int, bytes_per_device: int, shareable: bool = False) -> "MulticastGroup":
        h = C.c_void_p()
        _check(lib().kvbm_mc_group_create(num_devices, bytes_per_device, int(shareable), C.byref(h)))
        return MulticastGroup(h.value, num_devices)

    @staticmethod
    def from_fd(fd: int, num_devices: int, bytes_per_device: int) -> "MulticastGroup":
        h = C.c_void_p()
        _check(lib().kvbm_mc_group_import_fd(fd, num_devices, bytes_per_device, C.byref(h)))
        return MulticastGroup(h.value, num_devices)

    @property
    def size(self) -> int:
        return lib().kvbm_mc_group_size(self._h)

    def export_fd(self) -> int:
        fd = C.c_int(-1)
        _check(lib().kvbm_mc_group_export_fd(self._h, C.byref(fd)))
        return fd.value

    def add_device(self, device: int) -> None:
        _check(lib().kvbm_mc_group_add_device(self._h, device))

    def bind_local(self, device: int) -> int:
        """Allocate + bind this device's pool; returns its ordinary (unicast) device address."""
        p = C.c_void_p()
        _check(lib().kvbm_mc_group_bind_local(self._h, device, C.byref(p)))
        return p.value

    def bind_addr(self, device: int, ptr: int, nbytes: int) -> None:
        """Bind memory the engine already owns (cuMemCreate-backed, granularity-aligned) instead of a pool allocated here."""
        _check(lib().kvbm_mc_group_bind_addr(self._h, device, ptr, nbytes))

    def map(self, device: int) -> int:
        """Map the multicast object for the sending device; returns the multicast address of offset 0."""
        p = C.c_void_p()
        _check(lib().kvbm_mc_group_map(self._h, device, C.byref(p)))
        return p.value

    def close(self) -> None:
        if self._h:
            lib().kvbm_mc_group_destroy(self._h)
            self._h = C.c_void_p()

    def detach(self) -> None:
        """Forget the group without tearing it down (the driver reclaims it at process exit).  Unbinding a multicast
        object takes the driver seconds; a process that is about to exit need not wait for it."""
        self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

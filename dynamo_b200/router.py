"""ctypes mirror of the reference's KV routing index (lib/kv-router), over libkvbm_router.so.

  compute_block_hash_for_seq / compute_seq_hash_for_block   lib/kv-router/src/protocols.rs:74-172
  RadixTree.apply_event / find_matches / remove_worker / clear_all_blocks   indexer/radix_tree.rs:165-500
  OverlapScores {scores, frequencies, tree_sizes}           protocols.rs:733-788
  RouterEvent / KvCacheEvent JSON (serde snake_case tags)   protocols.rs:473-520,619-636,695-735
`prefix_hit_block_table` is the bridge to the transfer path: the blocks a decode worker already holds are skipped.
"""
from __future__ import annotations

import ctypes as C
import json
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from . import _lib

ROUTER_SO = _lib.KERNELS_SO.replace("libkvbm_kernels.so", "libkvbm_router.so")
XXH3_SEED = 1337
KV_EVENT_SUBJECT = "kv-events"


class KvCacheEventError(RuntimeError):
    NAMES = {1: "ParentBlockNotFound", 2: "BlockNotFound", 3: "InvalidBlockSequence", 4: "InvalidArgument"}

    def __init__(self, code: int):
        super().__init__(self.NAMES.get(code, str(code)))
        self.code = code
        self.kind = self.NAMES.get(code, str(code))


_cfg = False


class _DumpEvent(C.Structure):
    """struct kvr_dump_event"""
    _fields_ = [("worker_id", C.c_uint64), ("dp_rank", C.c_uint32), ("has_parent", C.c_uint32), ("event_id", C.c_uint64),
                ("parent_hash", C.c_uint64), ("block_hash", C.c_uint64), ("tokens_hash", C.c_uint64)]


def lib() -> C.CDLL:
    global _cfg
    L = _lib.load(ROUTER_SO)
    if not _cfg:
        u64, u32, sz, vp, i = C.c_uint64, C.c_uint32, C.c_size_t, C.c_void_p, C.c_int
        P = C.POINTER
        L.kvr_compute_hash.argtypes = [vp, sz]
        L.kvr_compute_hash.restype = u64
        L.kvr_compute_block_hash_for_seq.argtypes = [P(u32), sz, u32, C.c_char_p, i, P(u64), sz]
        L.kvr_compute_block_hash_for_seq.restype = sz
        L.kvr_compute_seq_hash_for_block.argtypes = [P(u64), sz, P(u64)]
        L.kvr_compute_seq_hash_for_block.restype = None
        L.kvr_tree_create.argtypes = [C.c_int64]
        L.kvr_tree_create.restype = vp
        L.kvr_tree_destroy.argtypes = [vp]
        L.kvr_tree_destroy.restype = None
        L.kvr_tree_apply_stored.argtypes = [vp, u64, u32, u64, i, u64, sz, P(u64), P(u64)]
        L.kvr_tree_apply_removed.argtypes = [vp, u64, u32, u64, sz, P(u64)]
        L.kvr_tree_apply_cleared.argtypes = [vp, u64, u32]
        L.kvr_tree_remove_worker.argtypes = [vp, u64]
        L.kvr_tree_remove_worker.restype = None
        L.kvr_tree_remove_worker_dp_rank.argtypes = [vp, u64, u32]
        L.kvr_tree_remove_worker_dp_rank.restype = None
        L.kvr_tree_clear_all_blocks.argtypes = [vp, u64]
        L.kvr_tree_clear_all_blocks.restype = None
        L.kvr_tree_get_workers.argtypes = [vp, P(u64), sz]
        L.kvr_tree_get_workers.restype = sz
        L.kvr_tree_lookup_size.argtypes = [vp, u64, u32]
        L.kvr_tree_lookup_size.restype = C.c_int64
        L.kvr_tree_lookup_len.argtypes = [vp]
        L.kvr_tree_lookup_len.restype = sz
        L.kvr_tree_node_info.argtypes = [vp, P(u64), sz, P(sz), P(sz)]
        L.kvr_tree_current_size.argtypes = [vp]
        L.kvr_tree_current_size.restype = sz
        L.kvr_tree_dump_events.argtypes = [vp, P(_DumpEvent), sz]
        L.kvr_tree_dump_events.restype = sz
        L.kvr_tree_find_matches.argtypes = [vp, P(u64), sz, i, P(u64), P(u32), P(u32), P(u64), sz, P(u64), sz, P(sz)]
        L.kvr_tree_find_matches.restype = sz
        _cfg = True
    return L


def _u64s(xs: Sequence[int]):
    return (C.c_uint64 * max(1, len(xs)))(*[int(x) & 0xFFFFFFFFFFFFFFFF for x in xs])


def compute_hash(data: bytes) -> int:
    return lib().kvr_compute_hash(data, len(data))


def compute_block_hash_for_seq(tokens: Sequence[int], kv_block_size: int, lora_name: Optional[str] = None,
                               is_eagle: bool = False) -> List[int]:
    n = len(tokens)
    arr = (C.c_uint32 * max(1, n))(*[int(t) for t in tokens])
    cap = n // max(1, kv_block_size) + 1
    out = (C.c_uint64 * cap)()
    k = lib().kvr_compute_block_hash_for_seq(arr, n, kv_block_size, lora_name.encode() if lora_name else None, int(is_eagle), out, cap)
    return [out[i] for i in range(k)]


def compute_seq_hash_for_block(block_hashes: Sequence[int]) -> List[int]:
    n = len(block_hashes)
    out = (C.c_uint64 * max(1, n))()
    lib().kvr_compute_seq_hash_for_block(_u64s(block_hashes), n, out)
    return [out[i] for i in range(n)]


@dataclass
class OverlapScores:
    scores: Dict[Tuple[int, int], int] = field(default_factory=dict)       # (worker_id, dp_rank) -> matched depth
    frequencies: List[int] = field(default_factory=list)
    tree_sizes: Dict[Tuple[int, int], int] = field(default_factory=dict)


class RadixTree:
    def __init__(self, expiration_ms: Optional[int] = None):
        self._h = lib().kvr_tree_create(-1 if expiration_ms is None else int(expiration_ms))

    def close(self):
        if self._h:
            lib().kvr_tree_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- events ------------------------------------------------------------------------------------------
    def apply_stored(self, worker_id: int, block_hashes: Sequence[int], tokens_hashes: Sequence[int],
                     parent_hash: Optional[int] = None, dp_rank: int = 0, event_id: int = 0) -> None:
        if len(block_hashes) != len(tokens_hashes):
            raise ValueError("block_hashes and tokens_hashes differ in length")
        rc = lib().kvr_tree_apply_stored(self._h, worker_id, dp_rank, event_id, int(parent_hash is not None),
                                         (parent_hash or 0) & 0xFFFFFFFFFFFFFFFF, len(block_hashes), _u64s(block_hashes), _u64s(tokens_hashes))
        if rc:
            raise KvCacheEventError(rc)

    def apply_removed(self, worker_id: int, block_hashes: Sequence[int], dp_rank: int = 0, event_id: int = 0) -> None:
        rc = lib().kvr_tree_apply_removed(self._h, worker_id, dp_rank, event_id, len(block_hashes), _u64s(block_hashes))
        if rc:
            raise KvCacheEventError(rc)

    def apply_cleared(self, worker_id: int, dp_rank: int = 0) -> None:
        rc = lib().kvr_tree_apply_cleared(self._h, worker_id, dp_rank)
        if rc:
            raise KvCacheEventError(rc)

    def apply_event(self, event) -> None:
        """A `RouterEvent` as serde_json writes it (dict / str / bytes):
        {"worker_id": u64, "event": {"event_id": u64, "data": {"stored": {"parent_hash": u64|null, "blocks":
        [{"block_hash": u64, "tokens_hash": u64}]}} | {"removed": {"block_hashes": [u64]}} | "cleared", "dp_rank": u32}}"""
        if isinstance(event, (bytes, str)):
            event = json.loads(event)
        wid = int(event["worker_id"])
        ev = event["event"]
        dp = int(ev.get("dp_rank", 0))
        eid = int(ev.get("event_id", 0))
        data = ev["data"]
        if data == "cleared":
            return self.apply_cleared(wid, dp)
        if "stored" in data:
            st = data["stored"]
            blocks = st["blocks"]
            return self.apply_stored(wid, [b["block_hash"] for b in blocks], [b["tokens_hash"] for b in blocks], st.get("parent_hash"), dp, eid)
        if "removed" in data:
            return self.apply_removed(wid, data["removed"]["block_hashes"], dp, eid)
        raise ValueError(f"unknown KvCacheEventData {data!r}")

    def remove_worker(self, worker_id: int) -> None:
        lib().kvr_tree_remove_worker(self._h, worker_id)

    def remove_worker_dp_rank(self, worker_id: int, dp_rank: int) -> None:
        lib().kvr_tree_remove_worker_dp_rank(self._h, worker_id, dp_rank)

    def clear_all_blocks(self, worker_id: int) -> None:
        lib().kvr_tree_clear_all_blocks(self._h, worker_id)

    def get_workers(self) -> List[int]:
        out = (C.c_uint64 * 4096)()
        n = lib().kvr_tree_get_workers(self._h, out, 4096)
        return [out[i] for i in range(min(n, 4096))]

    def current_size(self) -> int:
        """RadixTree::current_size (radix_tree.rs:567-569)."""
        return lib().kvr_tree_current_size(self._h)

    def dump_tree_as_events(self) -> List[dict]:
        """RadixTree::dump_tree_as_events (radix_tree.rs:505-565) as the RouterEvent dicts `apply_event` takes (serde shape)."""
        n = lib().kvr_tree_dump_events(self._h, None, 0)
        buf = (_DumpEvent * max(1, n))()
        n = min(n, lib().kvr_tree_dump_events(self._h, buf, n))
        return [{"worker_id": e.worker_id, "storage_tier": "device",
                 "event": {"event_id": e.event_id, "dp_rank": e.dp_rank,
                           "data": {"stored": {"parent_hash": e.parent_hash if e.has_parent else None,
                                               "blocks": [{"block_hash": e.block_hash, "tokens_hash": e.tokens_hash}]}}}}
                for e in buf[:n]]

    # -- queries ------------------------------------------------------------------------------------------
    def find_matches(self, sequence: Sequence[int], early_exit: bool = False) -> OverlapScores:
        cap, fcap = 4096, max(32, len(sequence) + 1)
        w, d, s, t = (C.c_uint64 * cap)(), (C.c_uint32 * cap)(), (C.c_uint32 * cap)(), (C.c_uint64 * cap)()
        f = (C.c_uint64 * fcap)()
        nf = C.c_size_t()
        k = lib().kvr_tree_find_matches(self._h, _u64s(sequence), len(sequence), int(early_exit), w, d, s, t, cap, f, fcap, C.byref(nf))
        res = OverlapScores()
        for i in range(min(k, cap)):
            res.scores[(w[i], d[i])] = s[i]
            res.tree_sizes[(w[i], d[i])] = t[i]
        res.frequencies = [f[i] for i in range(min(nf.value, fcap))]
        return res

    # -- introspection (tests) ----------------------------------------------------------------------------
    def lookup_size(self, worker_id: int, dp_rank: int = 0) -> Optional[int]:
        v = lib().kvr_tree_lookup_size(self._h, worker_id, dp_rank)
        return None if v < 0 else int(v)

    def lookup_len(self) -> int:
        return lib().kvr_tree_lookup_len(self._h)

    def node_info(self, path: Sequence[int]) -> Optional[Tuple[int, int]]:
        nw, nc = C.c_size_t(), C.c_size_t()
        rc = lib().kvr_tree_node_info(self._h, _u64s(path), len(path), C.byref(nw), C.byref(nc))
        return None if rc else (nw.value, nc.value)


def prefix_hit_block_table(tree: RadixTree, worker: Tuple[int, int], request_block_hashes: Sequence[int],
                           src_block_ids: Sequence[int], dst_block_ids: Sequence[int]) -> Tuple[List[int], List[int], int]:
    """What the prefill->decode hand-off actually has to move: the decode `worker` already holds the first
    `matched` blocks of this request (its overlap score), so only the suffix is transferred (SURVEY §8d cfg5:
    "only the non-hit suffix ceil((1-h)*blocks) is moved")."""
    scores = tree.find_matches(request_block_hashes, False).scores
    matched = min(scores.get(worker, 0), len(src_block_ids))
    return list(src_block_ids[matched:]), list(dst_block_ids[matched:]), matched


# ------------------------------------------------------------------------------------------------------------------
# KV event publisher: the reference's C ABI (lib/bindings/c/src/lib.rs:112-116,328-391) through ctypes
# ------------------------------------------------------------------------------------------------------------------
_EVENT_CB = C.CFUNCTYPE(None, C.c_char_p, C.c_size_t, C.c_void_p)
_ev_cfg = False


def _ev_lib() -> C.CDLL:
    global _ev_cfg
    L = lib()
    if not _ev_cfg:
        u64, u32, sz, vp = C.c_uint64, C.c_uint32, C.c_size_t, C.c_void_p
        P = C.POINTER
        L.dynamo_llm_init.argtypes = [C.c_char_p, C.c_char_p, u32]
        L.dynamo_llm_init.restype = u32
        L.dynamo_llm_shutdown.restype = u32
        L.dynamo_llm_load_publisher_create.restype = u32
        L.dynamo_kv_event_publish_stored.argtypes = [u64, P(u32), P(sz), P(u64), sz, P(u64), C.c_char_p]
        L.dynamo_kv_event_publish_stored.restype = u32
        L.dynamo_kv_event_publish_removed.argtypes = [u64, P(u64), sz]
        L.dynamo_kv_event_publish_removed.restype = u32
        L.dynamo_kv_event_set_worker_id.argtypes = [u64]
        L.dynamo_kv_event_attach_tree.argtypes = [vp]
        L.dynamo_kv_event_subscribe.argtypes = [_EVENT_CB, vp]
        L.dynamo_kv_event_published_count.restype = u64
        _ev_cfg = True
    return L


class KvEventPublisher:
    """`dynamo_llm_init` + `dynamo_kv_event_publish_stored/removed` as an engine's C++ executor thread calls them.
    One per process (the reference keeps the publisher in a process-wide OnceCell).  Events reach the attached RadixTree
    (in-process indexer) and `on_event(json_bytes)`; in Dynamo they ride the runtime's "kv-events" subject."""

    def __init__(self, namespace: str, component: Optional[str], kv_block_size: int, worker_id: int = 0,
                 tree: Optional[RadixTree] = None, on_event=None):
        L = _ev_lib()
        if L.dynamo_llm_init(namespace.encode(), component.encode() if component else None, kv_block_size) != 0:
            raise RuntimeError("dynamo_llm_init failed")
        L.dynamo_kv_event_set_worker_id(worker_id)
        self._tree = tree
        L.dynamo_kv_event_attach_tree(tree._h if tree is not None else None)
        self._cb = _EVENT_CB(lambda p, n, _u: on_event(C.string_at(p, n))) if on_event is not None else _EVENT_CB(0)
        L.dynamo_kv_event_subscribe(self._cb, None)
        self.kv_block_size = kv_block_size

    def publish_stored(self, event_id: int, token_ids: Sequence[int], num_block_tokens: Sequence[int], block_ids: Sequence[int],
                       parent_hash: Optional[int] = None, lora_name: Optional[str] = None) -> bool:
        toks = (C.c_uint32 * max(1, len(token_ids)))(*[int(t) for t in token_ids])
        nbt = (C.c_size_t * max(1, len(num_block_tokens)))(*[int(x) for x in num_block_tokens])
        ph = C.byref(C.c_uint64(parent_hash & 0xFFFFFFFFFFFFFFFF)) if parent_hash is not None else None
        return _ev_lib().dynamo_kv_event_publish_stored(event_id, toks, nbt, _u64s(block_ids), len(block_ids), ph,
                                                        lora_name.encode() if lora_name else None) == 0

    def publish_removed(self, event_id: int, block_ids: Sequence[int]) -> bool:
        return _ev_lib().dynamo_kv_event_publish_removed(event_id, _u64s(block_ids), len(block_ids)) == 0

    @staticmethod
    def published_count() -> int:
        return int(_ev_lib().dynamo_kv_event_published_count())

    def shutdown(self) -> None:
        _ev_lib().dynamo_llm_shutdown()

"""Pure-Python restatement of the reference's KV routing index (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/lib/kv-router/src: protocols.rs:20-172 (hashing; XXH3 from the `xxhash` wheel = the same
third-party algorithm the reference takes from the xxhash-rust crate), indexer/radix_tree.rs:165-500 (tree),
active_set.rs:9-40.  The reference pins hashes by properties only (no constants), so hash parity is anchored on the
published XXH3 algorithm; tree parity is pinned by the reference's own scenario tests (tests/test_router.py)."""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Sequence, Tuple

import xxhash

SEED = 1337
M64 = (1 << 64) - 1


def compute_hash(data: bytes) -> int:
    return xxhash.xxh3_64_intdigest(data, seed=SEED)


def compute_block_hash_for_seq(tokens: Sequence[int], kv_block_size: int, lora_name: Optional[str] = None, is_eagle: bool = False) -> List[int]:
    if kv_block_size == 0:
        return []
    seed = SEED
    if lora_name:
        seed = (SEED + xxhash.xxh3_64_intdigest(lora_name.encode())) & M64
    stride = kv_block_size
    window = stride + 1 if is_eagle else stride
    out, start = [], 0
    while start + window <= len(tokens):
        out.append(xxhash.xxh3_64_intdigest(struct.pack(f"<{window}I", *tokens[start:start + window]), seed=seed))
        start += stride
    return out


def compute_seq_hash_for_block(block_hashes: Sequence[int]) -> List[int]:
    out: List[int] = []
    for i, h in enumerate(block_hashes):
        out.append(h if i == 0 else compute_hash(struct.pack("<QQ", out[-1], h)))
    return out


class _Block:
    __slots__ = ("children", "workers", "block_hash")

    def __init__(self, block_hash=None):
        self.children: Dict[int, "_Block"] = {}
        self.workers = set()
        self.block_hash = block_hash

    def drop_worker(self, w):
        self.workers.discard(w)
        if not self.workers:
            self.children.clear()


class RadixTree:
    def __init__(self):
        self.root = _Block()
        self.lookup: Dict[Tuple[int, int], Dict[int, _Block]] = {}

    def apply_stored(self, worker_id, block_hashes, tokens_hashes, parent_hash=None, dp_rank=0):
        w = (worker_id, dp_rank)
        wl = self.lookup.setdefault(w, {})
        if parent_hash is not None:
            if parent_hash not in wl:
                return "ParentBlockNotFound"
            cur = wl[parent_hash]
        else:
            cur = self.root
        need = False
        for bh, th in zip(block_hashes, tokens_hashes):
            if need:
                cur.workers.add(w)
            need = True
            child = cur.children.get(th)
            if child is None:
                child = wl.get(bh) or _Block(bh)
                cur.children[th] = child
            if child is cur:
                return "InvalidBlockSequence"
            wl[bh] = child
            cur = child
        if need:
            cur.workers.add(w)
        return None

    def apply_removed(self, worker_id, block_hashes, dp_rank=0):
        w = (worker_id, dp_rank)
        wl = self.lookup.setdefault(w, {})
        err = None
        for bh in block_hashes:
            blk = wl.get(bh)
            if blk is None:
                err = err or "BlockNotFound"
                continue
            blk.drop_worker(w)
            del wl[bh]
        return err

    def _remove_or_clear(self, worker_id, keep):
        for w in [k for k in self.lookup if k[0] == worker_id]:
            for blk in self.lookup.pop(w).values():
                blk.drop_worker(w)
            if keep:
                self.lookup[w] = {}

    def apply_cleared(self, worker_id, dp_rank=0):
        self.lookup.setdefault((worker_id, dp_rank), {})
        self._remove_or_clear(worker_id, True)

    def clear_all_blocks(self, worker_id):
        self._remove_or_clear(worker_id, True)

    def remove_worker(self, worker_id):
        self._remove_or_clear(worker_id, False)

    def find_matches(self, sequence, early_exit=False):
        scores: Dict[Tuple[int, int], int] = {}
        if not sequence:
            return scores
        cur = self.root.children.get(sequence[0])
        if cur is None:
            return scores
        active = set(cur.workers)
        if not active:
            return scores
        if early_exit and len(active) == 1:
            return {w: 1 for w in active}
        depth = 1
        for idx in range(1, len(sequence)):
            blk = cur.children.get(sequence[idx])
            if blk is None:
                break
            if len(blk.workers) != len(active):
                for w in list(active):
                    if w not in blk.workers:
                        scores[w] = depth
                        active.discard(w)
            if not active:
                break
            if early_exit and len(active) == 1:
                depth = idx + 1
                break
            cur = blk
            depth = idx + 1
        for w in active:
            scores[w] = depth
        return scores

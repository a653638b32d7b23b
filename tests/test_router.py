"""KV routing index: C++ RadixTree + XXH3 hashing vs the reference's own scenario tests and a
pure-Python restatement.

Reference tests transcribed (relative to /root/reference/lib/kv-router/src):
  indexer/radix_tree.rs:578-858   test_radix_tree            (store / partial remove / parent-hash continuation)
  indexer/radix_tree.rs:859-901   test_radix_tree_apply_event_errors
  indexer/radix_tree.rs:902-1033  test_clear_all_blocks
  indexer/tests.rs:500-560,786-812,1008-1048   partial match, shared prefix, dp ranks, clear clears all dp ranks
  protocols.rs:929-1187           hash determinism / LoRA / eagle-window properties
test_utils.rs:60-101: make_blocks(i) -> tokens_hash = i, block_hash = i*100.
"""
import json
import struct

import numpy as np
import pytest
import xxhash

from dynamo_b200 import router as R
from oracle import router_oracle as RO

W0, W1 = (0, 0), (1, 0)


def store(t, worker, ids, parent=None, dp=0):
    t.apply_stored(worker, [i * 100 for i in ids], list(ids), None if parent is None else parent, dp)


def remove(t, worker, ids, dp=0):
    t.apply_removed(worker, [i * 100 for i in ids], dp)


def test_radix_tree_reference_scenario():
    t = R.RadixTree()
    store(t, 0, [1, 2, 3])
    assert t.find_matches([1, 2, 3]).scores[W0] == 3
    assert t.lookup_len() == 1 and t.lookup_size(0) == 3
    assert t.node_info([]) == (0, 1) and t.node_info([1]) == (1, 1)
    store(t, 1, [1, 4, 5])
    s = t.find_matches([1, 2, 3]).scores
    assert s[W0] == 3 and s[W1] == 1
    assert t.lookup_len() == 2 and t.lookup_size(0) == 3 and t.lookup_size(1) == 3
    assert t.node_info([]) == (0, 1) and t.node_info([1]) == (2, 2)
    remove(t, 1, [5])
    assert t.lookup_size(0) == 3 and t.lookup_size(1) == 2 and t.node_info([1]) == (2, 2)
    remove(t, 1, [4])
    assert t.lookup_size(1) == 1 and t.node_info([1]) == (2, 2)
    store(t, 1, [2, 6, 7], parent=100)                        # continue under block_hash 100
    s = t.find_matches([1, 2, 3]).scores
    assert s[W0] == 3 and s[W1] == 2
    assert t.lookup_size(0) == 3 and t.lookup_size(1) == 4
    assert t.node_info([1]) == (2, 2) and t.node_info([1, 2])[0] == 2   # block 200 now has both workers
    ts = t.find_matches([1, 2, 3]).tree_sizes
    assert ts[W0] == 3 and ts[W1] == 4


def test_apply_event_errors():
    t = R.RadixTree()
    with pytest.raises(R.KvCacheEventError) as e:
        store(t, 0, [1, 2, 3], parent=12345)
    assert e.value.kind == "ParentBlockNotFound"
    with pytest.raises(R.KvCacheEventError) as e:
        remove(t, 0, [1, 2, 3])
    assert e.value.kind == "BlockNotFound"
    store(t, 0, [1])
    with pytest.raises(R.KvCacheEventError) as e:             # parent appears in its own continuation
        store(t, 0, [1, 2, 3], parent=100)
    assert e.value.kind == "InvalidBlockSequence"


def test_clear_all_blocks_reference_scenario():
    t = R.RadixTree()
    assert t.find_matches([0]).scores == {}
    t.clear_all_blocks(0)
    assert t.lookup_size(0) is None
    store(t, 0, [0, 1, 3])
    store(t, 1, [0, 2, 3])
    assert t.find_matches([0]).scores == {W0: 1, W1: 1}
    t.clear_all_blocks(0)
    assert t.lookup_size(0) == 0
    assert t.find_matches([0, 2]).scores == {W1: 2}
    assert t.find_matches([0, 1, 3]).scores == {W1: 1}
    store(t, 0, [4, 5])
    assert t.find_matches([4, 5]).scores == {W0: 2}
    t.clear_all_blocks(0)
    t.clear_all_blocks(0)
    assert t.lookup_size(0) == 0
    t.clear_all_blocks(1)
    assert t.lookup_len() == 2 and t.lookup_size(0) == 0 and t.lookup_size(1) == 0
    store(t, 0, [6])
    store(t, 1, [6])
    t.remove_worker(0)
    t.clear_all_blocks(0)
    assert t.lookup_size(0) is None
    assert t.find_matches([6]).scores == {W1: 1}
    t.clear_all_blocks(2)
    assert t.lookup_size(2) is None and t.lookup_size(1) is not None
    assert t.find_matches([6]).scores == {W1: 1}


def test_partial_match_shared_prefix_dp_ranks_and_early_exit():
    t = R.RadixTree()
    store(t, 0, [1, 2, 3])
    assert t.find_matches([1, 2, 9]).scores == {W0: 2}         # tests.rs:500-511
    assert t.find_matches([9, 1, 2]).scores == {}              # miss query
    assert t.find_matches([]).scores == {}                     # empty query
    store(t, 1, [1, 2, 4])
    s = t.find_matches([1, 2, 3]).scores
    assert s == {W0: 3, W1: 2}                                 # tests.rs:529-560
    store(t, 0, [1, 2, 3], dp=1)                               # tests.rs:786-812: dp ranks are separate workers
    s = t.find_matches([1, 2, 3]).scores
    assert s[(0, 0)] == 3 and s[(0, 1)] == 3 and s[W1] == 2
    t.apply_cleared(0)                                         # tests.rs:1008-1048: cleared drops every dp rank of the worker
    assert t.find_matches([1, 2, 3]).scores == {W1: 2}
    assert t.lookup_size(0, 0) == 0 and t.lookup_size(0, 1) == 0
    assert t.get_workers() == [0, 1]
    t2 = R.RadixTree()
    store(t2, 7, [1, 2, 3, 4])
    assert t2.find_matches([1, 2, 3, 4], early_exit=True).scores == {(7, 0): 1}   # single active worker exits at depth 1


def test_remove_mid_chain_block_is_not_cascaded():
    # radix_tree.rs:243-256 + tests.rs:851-903: a Removed event does not cascade; descendants keep stale workers
    t = R.RadixTree()
    store(t, 0, [1, 2, 3])
    remove(t, 0, [2])
    assert t.find_matches([1, 2, 3]).scores == {W0: 1}
    assert t.lookup_size(0) == 2


def test_router_event_json_shapes():
    t = R.RadixTree()
    ev = {"worker_id": 3, "event": {"event_id": 1, "data": {"stored": {"parent_hash": None, "blocks": [
        {"block_hash": 100, "tokens_hash": 1}, {"block_hash": 200, "tokens_hash": 2}]}}, "dp_rank": 0}}
    t.apply_event(json.dumps(ev))
    assert t.find_matches([1, 2]).scores == {(3, 0): 2}
    t.apply_event({"worker_id": 3, "event": {"event_id": 2, "data": {"removed": {"block_hashes": [200]}}, "dp_rank": 0}})
    assert t.find_matches([1, 2]).scores == {(3, 0): 1}
    t.apply_event({"worker_id": 3, "event": {"event_id": 3, "data": "cleared", "dp_rank": 0}})
    assert t.find_matches([1, 2]).scores == {}


def test_frequency_tracking():
    t = R.RadixTree(expiration_ms=60_000)                      # tests.rs:1955-2040 in miniature
    store(t, 0, [1, 2, 3])
    assert t.find_matches([1, 2, 3]).frequencies == []         # first visit: zeros are omitted
    assert t.find_matches([1, 2, 3]).frequencies == [1, 1, 1]
    assert t.find_matches([1, 2]).frequencies == [2, 2]
    assert t.find_matches([1, 2, 3]).frequencies == [3, 3, 2]
    assert R.RadixTree().find_matches([1]).frequencies == []


# ---------------------------------------------------------------- hashing (protocols.rs)
def test_hash_definitions_against_xxh3():
    assert R.compute_hash(b"hello") == xxhash.xxh3_64_intdigest(b"hello", seed=1337)          # protocols.rs:20-25
    toks = list(range(1000, 1100))
    h = R.compute_block_hash_for_seq(toks, 16)
    assert len(h) == 6                                                                         # only full blocks
    assert h[2] == xxhash.xxh3_64_intdigest(struct.pack("<16I", *toks[32:48]), seed=1337)      # le_bytes(u32)
    assert h == RO.compute_block_hash_for_seq(toks, 16)
    s = R.compute_seq_hash_for_block(h)
    assert s[0] == h[0] and s[1] == xxhash.xxh3_64_intdigest(struct.pack("<QQ", s[0], h[1]), seed=1337)
    assert s == RO.compute_seq_hash_for_block(h)
    assert R.compute_block_hash_for_seq(toks, 0) == [] and R.compute_block_hash_for_seq([], 16) == []
    assert R.compute_seq_hash_for_block([]) == []


def test_hash_properties_lora_and_eagle():
    toks = list(np.random.default_rng(0).integers(0, 50000, 70))
    base = R.compute_block_hash_for_seq(toks, 16)
    assert base == R.compute_block_hash_for_seq(toks, 16)                                      # deterministic
    assert base == R.compute_block_hash_for_seq(toks, 16, lora_name="")                        # empty name == base model
    a, b = R.compute_block_hash_for_seq(toks, 16, "adapter-a"), R.compute_block_hash_for_seq(toks, 16, "adapter-b")
    assert len(a) == len(base) and all(x != y for x, y in zip(a, base)) and all(x != y for x, y in zip(a, b))
    assert a == RO.compute_block_hash_for_seq(toks, 16, "adapter-a")
    seed = (1337 + xxhash.xxh3_64_intdigest(b"adapter-a")) & ((1 << 64) - 1)                   # seed mixing, protocols.rs:83-86
    assert a[0] == xxhash.xxh3_64_intdigest(struct.pack("<16I", *toks[:16]), seed=seed)
    eagle = R.compute_block_hash_for_seq(toks, 16, is_eagle=True)                              # window 17, stride 16
    assert len(eagle) == (len(toks) - 1) // 16
    assert eagle[1] == xxhash.xxh3_64_intdigest(struct.pack("<17I", *toks[16:33]), seed=1337)
    assert eagle == RO.compute_block_hash_for_seq(toks, 16, is_eagle=True)


# ---------------------------------------------------------------- randomized differential vs the Python restatement
def test_random_event_streams_match_oracle():
    rng = np.random.default_rng(42)
    for trial in range(30):
        t, o = R.RadixTree(), RO.RadixTree()
        for step in range(120):
            w, dp = int(rng.integers(0, 4)), int(rng.integers(0, 2))
            kind = rng.integers(0, 10)
            if kind < 6:
                n = int(rng.integers(1, 6))
                ids = [int(x) for x in rng.integers(1, 12, n)]
                parent = None
                known = list(o.lookup.get((w, dp), {}).keys())
                if known and rng.integers(0, 2):
                    parent = int(known[int(rng.integers(0, len(known)))])
                want = o.apply_stored(w, [i * 100 for i in ids], ids, parent, dp)
                try:
                    t.apply_stored(w, [i * 100 for i in ids], ids, parent, dp)
                    got = None
                except R.KvCacheEventError as e:
                    got = e.kind
                assert got == want, (trial, step)
            elif kind < 8:
                ids = [int(x) for x in rng.integers(1, 12, int(rng.integers(1, 4)))]
                want = o.apply_removed(w, [i * 100 for i in ids], dp)
                try:
                    t.apply_removed(w, [i * 100 for i in ids], dp)
                    got = None
                except R.KvCacheEventError as e:
                    got = e.kind
                assert got == want
            elif kind == 8:
                o.apply_cleared(w, dp)
                t.apply_cleared(w, dp)
            else:
                o.remove_worker(w)
                t.remove_worker(w)
            q = [int(x) for x in rng.integers(1, 12, int(rng.integers(1, 7)))]
            for ee in (False, True):
                assert t.find_matches(q, ee).scores == o.find_matches(q, ee), (trial, step, q, ee)
            assert t.lookup_len() == len(o.lookup)


def test_prefix_hit_block_table_feeds_the_transfer_path():
    # configs[4]: a decode worker that already caches the first blocks of the prompt only needs the suffix moved
    block_size, n_blocks = 16, 20
    tokens = list(np.random.default_rng(1).integers(0, 32000, block_size * n_blocks))
    local = R.compute_block_hash_for_seq(tokens, block_size)
    seq = R.compute_seq_hash_for_block(local)
    t = R.RadixTree()
    t.apply_stored(5, seq[:12], local[:12])                   # decode worker 5 holds the first 12 blocks (h = 0.6)
    src_ids, dst_ids = list(range(100, 120)), list(range(40, 60))
    s, d, matched = R.prefix_hit_block_table(t, (5, 0), local, src_ids, dst_ids)
    assert matched == 12 and s == src_ids[12:] and d == dst_ids[12:]
    s, d, matched = R.prefix_hit_block_table(t, (6, 0), local, src_ids, dst_ids)
    assert matched == 0 and len(s) == 20


# ------------------------------------------------------------------------------------------------------------------
# KV event publisher C ABI (lib/bindings/c/src/lib.rs:112-116,228-391)
# ------------------------------------------------------------------------------------------------------------------
def test_compute_block_hash_golden_constant():
    """protocols.rs:937-951 (test_router_event_new): compute_block_hash(b"test data") == 13226331709069118873."""
    assert R.compute_hash(b"test data") == 13226331709069118873
    assert xxhash.xxh3_64_intdigest(b"test data", seed=1337) == 13226331709069118873


def test_event_publish_before_init_is_an_error_and_shutdown_without_init_too():
    L = R._ev_lib()
    L.dynamo_llm_shutdown()                                         # whatever an earlier test left behind
    assert L.dynamo_llm_shutdown() == 1                             # "Runtime not initialized" (lib.rs:177-190)
    ids = (R.C.c_uint64 * 1)(7)
    assert L.dynamo_kv_event_publish_removed(1, ids, 1) == 1        # the reference unwraps KV_PUB: no publisher, no publish
    assert L.dynamo_llm_load_publisher_create() == 0                # lib.rs:192-194
    assert L.dynamo_llm_init(None, b"backend", 4) == 1              # namespace must be a C string (lib.rs:149-155)


def test_event_publish_stored_and_removed_feed_the_radix_tree_and_emit_router_event_json():
    KB = 4
    events = []
    tree = R.RadixTree()
    R._ev_lib().dynamo_llm_shutdown()
    pub = R.KvEventPublisher("ns", None, KB, worker_id=42, tree=tree, on_event=events.append)
    try:
        tokens = list(range(100, 100 + 3 * KB))
        want_hashes = R.compute_block_hash_for_seq(tokens, KB)
        assert want_hashes == [xxhash.xxh3_64_intdigest(struct.pack("<4I", *tokens[i:i + KB]), seed=1337) for i in range(0, 12, KB)]
        assert pub.publish_stored(1, tokens, [KB, KB, KB], [1000, 1001, 1002])
        ev = json.loads(events[-1])
        assert ev == {"worker_id": 42, "storage_tier": "device",
                      "event": {"event_id": 1, "dp_rank": 0,
                                "data": {"stored": {"parent_hash": None,
                                                    "blocks": [{"block_hash": 1000 + i, "tokens_hash": want_hashes[i], "mm_extra_info": None}
                                                               for i in range(3)]}}}}
        assert tree.find_matches(want_hashes).scores == {(42, 0): 3} and tree.lookup_size(42) == 3
        # the JSON is what RadixTree.apply_event (the serde shape of RouterEvent) accepts: a second indexer converges
        twin = R.RadixTree()
        twin.apply_event(events[-1])
        assert twin.find_matches(want_hashes).scores == {(42, 0): 3}
        # continuation under a parent + a partial block: the first block whose token count != kv_block_size ends the event
        more = list(range(500, 500 + 2 * KB + 2))
        assert pub.publish_stored(2, more, [KB, KB, 2], [1003, 1004, 1005], parent_hash=1002)
        ev2 = json.loads(events[-1])["event"]["data"]["stored"]
        assert ev2["parent_hash"] == 1002 and [b["block_hash"] for b in ev2["blocks"]] == [1003, 1004]
        full = want_hashes + R.compute_block_hash_for_seq(more, KB)
        assert tree.find_matches(full).scores == {(42, 0): 5}
        # a partial block FIRST publishes an empty Stored event (kv_event_create_stored_from_parts breaks before pushing)
        assert pub.publish_stored(3, [1, 2], [2], [2000])
        assert json.loads(events[-1])["event"]["data"]["stored"]["blocks"] == []
        # LoRA adapters hash to different blocks (protocols.rs:74-110)
        assert pub.publish_stored(4, tokens[:KB], [KB], [3000], lora_name="adapter-a")
        th = json.loads(events[-1])["event"]["data"]["stored"]["blocks"][0]["tokens_hash"]
        assert th == R.compute_block_hash_for_seq(tokens[:KB], KB, lora_name="adapter-a")[0] != want_hashes[0]
        # removed
        assert pub.publish_removed(5, [1004])
        assert json.loads(events[-1])["event"]["data"] == {"removed": {"block_hashes": [1004]}}
        assert tree.find_matches(full).scores == {(42, 0): 4}
        # an event the indexer rejects surfaces as ERR (parent unknown -> ParentBlockNotFound)
        assert not pub.publish_stored(6, tokens[:KB], [KB], [4000], parent_hash=999999)
        assert not pub.publish_removed(7, [555])
        # a second init keeps the first publisher (OnceCell semantics): block size stays 4
        assert R._ev_lib().dynamo_llm_init(b"other", b"x", 64) == 0
        assert pub.publish_stored(8, list(range(4)), [4], [5000])
        assert len(json.loads(events[-1])["event"]["data"]["stored"]["blocks"]) == 1
        assert R.KvEventPublisher.published_count() >= 8
    finally:
        pub.shutdown()
        tree.close()


def test_dump_tree_as_events_rebuilds_the_tree_and_current_size_counts_blocks():
    """radix_tree.rs:505-569: a breadth-first dump of single-block Stored events (ids 0..n-1, parents before children)
    replayed into an empty tree gives the same scores, sizes and structure (router replica sync, kv_indexer.rs:243)."""
    t = R.RadixTree()
    assert t.current_size() == 0 and t.dump_tree_as_events() == []
    t.apply_stored(0, [100, 200, 300], [1, 2, 3])
    t.apply_stored(1, [100, 200, 400], [1, 2, 4], dp_rank=1)
    t.apply_stored(0, [500], [5], parent_hash=200)
    assert t.current_size() == 3 + 3 + 1
    ev = t.dump_tree_as_events()
    assert [e["event"]["event_id"] for e in ev] == list(range(len(ev))) and len(ev) == t.current_size()
    assert all(e["storage_tier"] == "device" and len(e["event"]["data"]["stored"]["blocks"]) == 1 for e in ev)
    seen = set()
    for e in ev:                                       # breadth-first: a block's parent was announced by an earlier event
        st = e["event"]["data"]["stored"]
        assert st["parent_hash"] is None or st["parent_hash"] in seen
        seen.add(st["blocks"][0]["block_hash"])
    roots = [e for e in ev if e["event"]["data"]["stored"]["parent_hash"] is None]
    assert {(e["worker_id"], e["event"]["dp_rank"]) for e in roots} == {(0, 0), (1, 1)}
    assert all(e["event"]["data"]["stored"]["blocks"][0] == {"block_hash": 100, "tokens_hash": 1} for e in roots)
    # random trees over real sequence hashes (block hash = hash of the whole prefix, so the structure is a tree)
    rng = np.random.default_rng(7)
    for trial in range(25):
        a = R.RadixTree()
        held = {}                                      # (w, dp) -> {block hash: parent hash or None}
        for step in range(80):
            w, dp = int(rng.integers(0, 4)), int(rng.integers(0, 2))
            mine = held.setdefault((w, dp), {})
            if not mine or rng.integers(0, 4):
                toks = [int(x) for x in rng.integers(1, 6, int(rng.integers(1, 7)))]
                seq = R.compute_seq_hash_for_block(toks)
                a.apply_stored(w, seq, toks, None, dp)
                for i, h in enumerate(seq):
                    mine.setdefault(h, seq[i - 1] if i else None)
            else:                                       # engines evict leaves: a held block none of whose children is held
                leaves = [h for h in mine if h not in set(mine.values())]
                h = leaves[int(rng.integers(0, len(leaves)))]
                a.apply_removed(w, [h], dp)
                del mine[h]
        b = R.RadixTree()
        for e in a.dump_tree_as_events():
            b.apply_event(json.dumps(e))               # through the serde shape, as it would cross the wire
        assert b.current_size() == a.current_size()
        for w in range(4):
            for dp in range(2):
                assert (b.lookup_size(w, dp) or 0) == (a.lookup_size(w, dp) or 0)   # a worker whose blocks are all gone is not re-created
        for _ in range(40):
            q = [int(x) for x in rng.integers(1, 6, int(rng.integers(1, 7)))]
            assert b.find_matches(q).scores == a.find_matches(q).scores
            ia, ib = a.node_info(q), b.node_info(q)
            # emptied blocks (no worker left) stay in the original as dead children and are not re-created by the replay
            assert (ia is None or ia[0] == 0) if ib is None else (ia is not None and ia[0] == ib[0] and ia[1] >= ib[1])
        a.close()
        b.close()
    # hand-made hashes that close a cycle (100 -> 200 -> 100): the dump still ends
    c = R.RadixTree()
    c.apply_stored(0, [100, 200], [1, 2])
    c.apply_stored(0, [200, 100], [2, 1])
    assert 0 < len(c.dump_tree_as_events()) < 64


def test_event_publish_is_safe_from_many_threads():
    """TRT-LLM's executor calls dynamo_kv_event_publish_* from its own threads (lib/bindings/c/src/lib.rs:328-391 goes
    through a static publisher): concurrent publishers must neither lose nor tear events."""
    import threading
    KB, THREADS, PER = 4, 8, 150
    events = []
    tree = R.RadixTree()
    R._ev_lib().dynamo_llm_shutdown()
    pub = R.KvEventPublisher("ns", "backend", KB, worker_id=7, tree=tree, on_event=events.append)
    before = R.KvEventPublisher.published_count()
    errors = []

    def worker(t):
        try:
            for i in range(PER):
                base = (t * PER + i) * 16
                toks = [base + k for k in range(2 * KB)]
                if not pub.publish_stored(t * 10000 + i, toks, [KB, KB], [base + 1, base + 2]):
                    errors.append(("stored", t, i))
                if i % 3 == 0 and not pub.publish_removed(t * 10000 + 5000 + i, [base + 2]):
                    errors.append(("removed", t, i))
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    try:
        ts = [threading.Thread(target=worker, args=(t,)) for t in range(THREADS)]
        for th in ts:
            th.start()
        for th in ts:
            th.join()
        assert errors == []
        n_removed = THREADS * len(range(0, PER, 3))
        assert R.KvEventPublisher.published_count() - before == THREADS * PER + n_removed == len(events)
        assert tree.current_size() == 2 * THREADS * PER - n_removed
        ids = set()
        for raw in events:                                   # every callback payload is one complete RouterEvent
            ev = json.loads(raw)
            assert ev["worker_id"] == 7
            ids.add(ev["event"]["event_id"])
        assert len(ids) == len(events)
        # a replica fed from the emitted JSON (in arrival order) converges to the same index
        twin = R.RadixTree()
        for raw in events:
            twin.apply_event(raw)
        assert twin.current_size() == tree.current_size()
        q = R.compute_block_hash_for_seq([16 * 37 + k for k in range(2 * KB)], KB)
        assert twin.find_matches(q).scores == tree.find_matches(q).scores != {}
        twin.close()
    finally:
        pub.shutdown()
        tree.close()


def test_remove_worker_verifies_hash_removal_and_default_tree():
    """radix_tree.rs:1035-1141 (test_radix_tree_default, test_remove_worker_verifies_hash_removal): create_store_event(worker,
    id, [h..]) stores block hashes h*100 under tokens hashes h."""
    t = R.RadixTree()
    assert t.lookup_len() == 0 and t.node_info([]) == (0, 0) and t.current_size() == 0      # empty root, empty lookup
    t.apply_stored(0, [100, 200, 300], [1, 2, 3])
    t.apply_stored(1, [100, 200, 300], [1, 2, 3])
    t.apply_stored(2, [100, 400, 500], [1, 4, 5])
    assert t.lookup_size(0) == 3
    assert t.node_info([1]) == (3, 2)                      # block 100: all three workers; children: tokens 2 and 4
    t.remove_worker(0)
    assert t.lookup_size(0) is None and t.lookup_len() == 2
    assert t.node_info([1])[0] == 2                        # workers 1 and 2 remain on block 100
    assert t.node_info([1, 2])[0] == 1                     # block 200: only worker 1
    assert t.find_matches([1, 2, 3]).scores == {(1, 0): 3, (2, 0): 1}
    assert t.get_workers() == [1, 2]
    t.close()


def test_radix_tree_memory_safety_fuzz_under_asan_ubsan(tmp_path):
    """tests/c/fuzz_router.cpp built from radix_tree.cpp with -fsanitize=address,undefined: random Stored / Removed /
    Cleared events, worker removal, queries and dumps -- once with real sequence hashes (a tree) and once with hand-made
    hashes that share blocks across depths and close cycles.  No use-after-free, no overflow, and no leak in either mode
    (the tree cuts every edge of every block it ever created when it is destroyed)."""
    import os
    import subprocess
    import pyarrow
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    probe = tmp_path / "p.cpp"
    probe.write_text("int main(){return 0;}\n")
    if subprocess.run(["g++", "-fsanitize=address,undefined", str(probe), "-o", str(tmp_path / "p")], capture_output=True).returncode != 0:
        pytest.skip("g++ has no AddressSanitizer runtime here")
    xxh = os.path.join(pyarrow.get_include(), "arrow", "vendored", "xxhash")
    exe = tmp_path / "fuzz_router"
    b = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
                        "-I", os.path.join(root, "include"), "-I", xxh, os.path.join(root, "tests/c/fuzz_router.cpp"),
                        os.path.join(root, "dynamo_b200/csrc/router/radix_tree.cpp"), "-o", str(exe)], capture_output=True, text=True)
    assert b.returncode == 0, b.stderr[-2000:]
    for mode in ("tree", "graph"):
        r = subprocess.run([str(exe), "300", mode], capture_output=True, text=True, timeout=600, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1"))
        assert r.returncode == 0 and "router fuzz ok" in r.stdout, (mode, r.stdout[-300:], r.stderr[-3000:])

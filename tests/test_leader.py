"""Leader side of the KVBM connector driving the worker end to end: the leader's slot state machine
produces ConnectorMetadata + BlockTransferRequests, the worker moves the bytes, prefix hits onboard from the host tier.

Reference behaviour mirrored (relative to /root/reference/lib/bindings/kvbm/src/block_manager/vllm/connector):
  leader.rs:216-283,287-327,329-538,540-598      get_num_new_matched_tokens / update_state_after_alloc /
                                                 build_connector_metadata / request_finished
  leader/slot.rs:593-896                         apply_scheduler_output (candidate blocks, priority filter, dedup contract)
  leader/slot.rs:977-1126                        acquire_local_matches (block-boundary rule)
  leader/slot.rs:1934-2270                       the slot's own unit tests (dedup / priority patterns), transcribed below
"""
import json

import pytest
import torch

from dynamo_b200.connector import (DEVICE, HOST, IMMEDIATE, LOAD, SCHEDULED, STORE, ConnectorMetadata, KvConnectorWorker)
from dynamo_b200.leader import (DECODING, FINISHING, ONBOARDING, PREFILLING, SKIPPED_PREFILL, KvbmRequest, KvConnectorLeader,
                                SchedulerOutput, SlotError)
from oracle import oracle as O

NB, NL, PAGE, HEADS, HD, HOSTB = 24, 2, 16, 2, 8, 12


def make_pair(offload_min_priority=0):
    caches = [(f"model.layers.{l}.attn", torch.zeros(2, NB, PAGE, HEADS, HD, dtype=torch.bfloat16)) for l in range(NL)]
    w = KvConnectorWorker(None, "worker-0", host_blocks=HOSTB)
    w.register_kv_caches(NB, PAGE, 0, 2, caches, [0] * NL)
    sent = []

    def send(req):
        sent.append(req)
        w.handle_block_transfer(json.dumps(req.to_json()))     # over the wire as JSON, like ZMQ would carry it
    leader = KvConnectorLeader("worker-0", PAGE, HOSTB, send, offload_min_priority=offload_min_priority)
    dev = O.Layout(O.LW, NB, NL, 2, PAGE, HEADS * HD, 2, block_dim=O.BLOCK_IS_SECOND_DIM, bases=[t.data_ptr() for _, t in caches])
    dev.fill_blocks(range(NB), -1)
    host = O.Layout(O.FC, HOSTB, NL, 2, PAGE, HEADS * HD, 2, bases=[w._host_mem.data_ptr()])
    return leader, w, caches, dev, host, sent


def forward_pass(w, caches, md_bytes):
    w.bind_connector_metadata(md_bytes)
    for name, _ in caches:
        w.save_kv_layer(name)
    w.clear_connector_metadata()


def settle(leader, w, sent):
    """What the leader's transfer tasks await in the reference: the worker-side completion of every request it sent."""
    w._poll()
    for req in sent:
        cr = req.connector_req
        slot = w.slots.get(cr.request_id)
        if slot is None or cr.uuid in slot.completed:
            leader.transfer_complete(cr.uuid)


def test_offload_then_prefix_hit_onboards_end_to_end():
    leader, w, caches, dev, host, sent = make_pair()
    prompt = list(range(1000, 1000 + 3 * PAGE + 8))                       # 3 full blocks + 8 tokens
    # ---- request A: nothing cached anywhere -> prefill, the 3 full blocks are offloaded Device -> Host
    leader.create_slot(KvbmRequest("A"), prompt)
    assert leader.get_num_new_matched_tokens("A", len(prompt), 0) == (0, False)
    blocks_a = [5, 9, 2, 17]
    leader.update_state_after_alloc("A", blocks_a, 0)
    so = SchedulerOutput()
    so.add_new_request("A", prompt, blocks_a, 0)
    so.add_num_scheduled_tokens({"A": len(prompt)})
    md = ConnectorMetadata.from_bytes(leader.build_connector_metadata(so))
    assert md.iteration == 1 and md.new_slots == [{"request_id": "A", "expected_immediate_ops": 0}]
    assert [(o.transfer_type, o.request_type) for o in md.operations] == [(STORE, SCHEDULED)]
    assert len(sent) == 1 and sent[0].from_pool == DEVICE and sent[0].to_pool == HOST
    assert [s for s, _ in sent[0].blocks] == blocks_a[:3] and sent[0].connector_req.uuid == md.operations[0].uuid
    assert leader.slots["A"].state == PREFILLING and leader.slots["A"].evaluated_blocks == 3
    forward_pass(w, caches, md.to_bytes())
    settle(leader, w, sent)
    for (s, hb) in sent[0].blocks:
        assert host.block_checksum(hb) == dev.block_checksum(s)
    assert len(leader.host.registered) == 3
    # decode step: one new token, no new full block -> no operation
    so = SchedulerOutput()
    so.add_cached_request("A", False, [7], [], len(prompt))
    so.add_num_scheduled_tokens({"A": 1})
    md = ConnectorMetadata.from_bytes(leader.build_connector_metadata(so))
    assert md.iteration == 2 and md.operations == [] and leader.slots["A"].state == DECODING
    forward_pass(w, caches, md.to_bytes())
    assert leader.request_finished("A", blocks_a) is True and not leader.has_slot("A")      # no pending ops -> Finished, removed
    assert w.get_finished(["A"])[0] == {"A"}
    assert leader.request_finished("never-seen", []) is False
    # ---- request B: same first 3 blocks -> matched in the host tier, onboarded into B's fresh device blocks
    prompt_b = prompt[:3 * PAGE] + list(range(5000, 5000 + 20))
    leader.create_slot(KvbmRequest("B"), prompt_b)
    n, is_async = leader.get_num_new_matched_tokens("B", len(prompt_b), 0)
    assert (n, is_async) == (3 * PAGE, True) and leader.matched_tokens == 3 * PAGE
    blocks_b = [11, 12, 13, 14, 15]
    sent.clear()
    leader.update_state_after_alloc("B", blocks_b[:3], n)                # first call: the blocks the external tokens land in
    assert leader.slots["B"].state == ONBOARDING and len(sent) == 1
    assert sent[0].from_pool == HOST and sent[0].to_pool == DEVICE and [d for _, d in sent[0].blocks] == blocks_b[:3]
    assert sent[0].connector_req.request_type == IMMEDIATE
    leader.update_state_after_alloc("B", blocks_b[3:], 0)                # second call: the rest of the prefill's blocks
    so = SchedulerOutput()                                               # vLLM lists an onboarding request as "new" only later
    md = ConnectorMetadata.from_bytes(leader.build_connector_metadata(so))
    assert md.new_slots == [{"request_id": "B", "expected_immediate_ops": 1}]
    assert [(o.transfer_type, o.request_type) for o in md.operations] == [(LOAD, IMMEDIATE)]
    forward_pass(w, caches, md.to_bytes())
    settle(leader, w, sent)
    for a_blk, b_blk in zip(blocks_a[:3], blocks_b[:3]):
        assert dev.block_checksum(b_blk) == dev.block_checksum(a_blk)   # B's blocks now hold A's KV
    assert w.get_finished([])[1] == {"B"}                                # finished onboarding (worker.rs:421-470)
    # B continues as a new request for the remaining tokens; the 4th block (20 new tokens -> 1 full block) is offloaded
    so = SchedulerOutput()
    so.add_new_request("B", prompt_b, blocks_b, 3 * PAGE)
    so.add_num_scheduled_tokens({"B": len(prompt_b) - 3 * PAGE})
    sent.clear()
    md = ConnectorMetadata.from_bytes(leader.build_connector_metadata(so))
    assert md.new_slots == [{"request_id": "B", "expected_immediate_ops": 0}]
    assert len(sent) == 1 and [s for s, _ in sent[0].blocks] == [blocks_b[3]]
    forward_pass(w, caches, md.to_bytes())
    settle(leader, w, sent)
    assert len(leader.host.registered) == 4
    # finishing while an offload is still outstanding -> Finishing until the engine reports it
    so = SchedulerOutput()
    so.add_cached_request("B", False, list(range(12)), [], len(prompt_b))   # 12 more tokens: 68 -> 80 = 5 full blocks
    so.add_num_scheduled_tokens({"B": 12})
    sent.clear()
    md = ConnectorMetadata.from_bytes(leader.build_connector_metadata(so))
    assert len(sent) == 1 and [s for s, _ in sent[0].blocks] == [blocks_b[4]]
    assert leader.request_finished("B", blocks_b) is True and leader.slots["B"].state == FINISHING
    forward_pass(w, caches, md.to_bytes())
    settle(leader, w, sent)
    assert not leader.has_slot("B")
    w.close()


def test_block_boundary_match_keeps_one_block_to_compute():
    """slot.rs:1079-1096: a match covering the WHOLE prompt drops its last block."""
    leader, w, caches, dev, host, sent = make_pair()
    prompt = list(range(2 * PAGE))
    leader.create_slot(KvbmRequest("A"), prompt)
    leader.update_state_after_alloc("A", [1, 2], 0)
    so = SchedulerOutput()
    so.add_new_request("A", prompt, [1, 2], 0)
    so.add_num_scheduled_tokens({"A": len(prompt)})
    forward_pass(w, caches, leader.build_connector_metadata(so))
    settle(leader, w, sent)
    leader.create_slot(KvbmRequest("B"), prompt)
    assert leader.get_num_new_matched_tokens("B", len(prompt), 0) == (PAGE, True)       # 2 matched, 1 kept for compute
    leader.create_slot(KvbmRequest("C"), prompt[:PAGE + 3])
    assert leader.get_num_new_matched_tokens("C", PAGE + 3, 0) == (PAGE, True)
    leader.create_slot(KvbmRequest("D"), prompt[:PAGE])
    assert leader.get_num_new_matched_tokens("D", PAGE, 0) == (0, False)                 # the only block is the boundary block
    leader.create_slot(KvbmRequest("E"), prompt[:PAGE - 1])
    assert leader.get_num_new_matched_tokens("E", PAGE - 1, 0) == (0, False)             # not even one full block
    leader.create_slot(KvbmRequest("S", salt_hash=77), prompt)
    assert leader.get_num_new_matched_tokens("S", len(prompt), 0) == (0, False)          # another salt: no sharing
    w.close()


def drain_offload_ids(sent):
    out = [[s for s, _ in r.blocks] for r in sent if r.from_pool == DEVICE]
    sent.clear()
    return out


def new_slot(leader, rid, n_tokens):
    leader.create_slot(KvbmRequest(rid), list(range(n_tokens)))
    return leader.slots[rid]


def test_slot_patterns_from_the_reference_unit_tests():
    """slot.rs:1934-2270."""
    leader, w, *_ , sent = make_pair()
    w.handle_block_transfer = lambda *a, **k: None                       # the engine side is not under test here
    leader.send = sent.append
    # vllm pattern (test_vllm_pattern_no_double_add): blocks arrive via update_state_after_alloc, new request applies []
    s = new_slot(leader, "v", 4 * PAGE)
    s.device_blocks.extend([10, 11, 12, 13])
    s.apply_scheduler_output([], [], 0, 4 * PAGE, None)
    assert s.device_blocks == [10, 11, 12, 13] and drain_offload_ids(sent) == [[10, 11, 12, 13]]
    # trtllm pattern: blocks arrive with the scheduler output itself
    s = new_slot(leader, "t", 4 * PAGE)
    s.apply_scheduler_output([], [20, 21, 22, 23], 0, 4 * PAGE, None)
    assert s.device_blocks == [20, 21, 22, 23] and drain_offload_ids(sent) == [[20, 21, 22, 23]]
    # partial overlap dedup: [3,4,5,6,7] + [6,7,8,9] -> [3..9]
    s = new_slot(leader, "d", 7 * PAGE)
    s.apply_scheduler_output([], [3, 4, 5, 6, 7], 0, 2 * PAGE, None)
    s.apply_scheduler_output([], [6, 7, 8, 9], 2 * PAGE, 2 * PAGE, None)
    assert s.device_blocks == [3, 4, 5, 6, 7, 8, 9] and drain_offload_ids(sent) == [[3, 4], [5, 6]]
    s.apply_scheduler_output([], [3, 4, 5, 6, 7, 8, 9], 4 * PAGE, PAGE, None)            # full overlap re-provision
    assert s.device_blocks == [3, 4, 5, 6, 7, 8, 9]
    with pytest.raises(AssertionError):                                   # test_invalid_overlap_panics: [..6,7] + [6,8,9]
        s.apply_scheduler_output([], [8, 11], 5 * PAGE, 0, None)
    with pytest.raises(AssertionError):                                   # non-contiguous duplicate
        s.apply_scheduler_output([], [30, 4], 5 * PAGE, 0, None)
    w.close()


def test_priority_filter_is_contiguous_and_terminates_offloading():
    """slot.rs:2039-2094 (test_priority_filtering_*): blocks are offloaded while priority >= threshold; the first block
    below it ends offloading for the request for good."""
    leader, w, *_, sent = make_pair(offload_min_priority=50)
    leader.send = sent.append
    s = new_slot(leader, "p", 6 * PAGE)
    s.apply_scheduler_output([], [1, 2, 3, 4], 0, 4 * PAGE, [80, 60, 10, 90])
    assert drain_offload_ids(sent) == [[1, 2]] and s.offload_terminated_at_block == 2 and s.evaluated_blocks == 4
    s.apply_scheduler_output([], [5, 6], 4 * PAGE, 2 * PAGE, [99, 99])                  # high priority, but offload is over
    assert drain_offload_ids(sent) == [] and s.current_position == 6 * PAGE
    s2 = new_slot(leader, "q", 2 * PAGE)
    s2.apply_scheduler_output([], [7, 8], 0, 2 * PAGE, [10, 99])                        # first block already below: nothing
    assert drain_offload_ids(sent) == [] and s2.offload_terminated_at_block == 0
    w.close()


def test_skipped_requests_resume_and_state_errors():
    leader, w, caches, dev, host, sent = make_pair()
    prompt = list(range(2 * PAGE + 4))
    leader.create_slot(KvbmRequest("A"), prompt)
    leader.update_state_after_alloc("A", [1, 2, 3], 0)
    so = SchedulerOutput()
    so.add_new_request("A", prompt, [1, 2, 3], 0)
    so.add_num_scheduled_tokens({"A": PAGE})                             # chunked prefill: first chunk only
    leader.build_connector_metadata(so)
    assert leader.slots["A"].state == PREFILLING
    leader.build_connector_metadata(SchedulerOutput())                   # in flight but not scheduled -> skipped
    assert leader.slots["A"].state == SKIPPED_PREFILL
    assert leader.get_num_new_matched_tokens("A", len(prompt), PAGE) == (0, False)      # resumes, returns early
    assert leader.slots["A"].state == PREFILLING
    with pytest.raises(SlotError):
        leader.slots["A"].acquire_local_matches(0)                        # not Initialized / Preempted
    with pytest.raises(SlotError):
        leader.slots["A"].trigger_onboarding(PAGE)                        # not OnboardStaged
    with pytest.raises(SlotError):
        leader.create_slot(KvbmRequest("A"), prompt)
    with pytest.raises(SlotError):
        leader.get_num_new_matched_tokens("missing", 1, 0)
    w.close()

"""Worker side of the KVBM connector -- wire formats and the per-iteration state machine.

Reference behaviour mirrored (relative to /root/reference):
  lib/bindings/kvbm/src/block_manager/vllm/connector.rs:158-185      ConnectorMetadata JSON
  lib/llm/src/block_manager/connector/protocol.rs:60-140             request / requirement enums (serde names)
  lib/llm/src/block_manager/distributed/utils.rs:47-84               BlockTransferRequest (connector_req omitted when None)
  lib/bindings/kvbm/src/block_manager/vllm/connector/worker.rs:237-470  bind / clear / save_kv_layer / get_finished
  lib/llm/src/block_manager/connector/scheduler.rs:148-268           slot completion = completed == len(operations)
  lib/llm/src/block_manager/layout.rs:163-176                        layer_separate_auto
"""
import json
import uuid

import numpy as np
import pytest
import torch

from dynamo_b200.connector import (DEVICE, HOST, IMMEDIATE, LOAD, SCHEDULED, STORE, BlockTransferRequest, ConnectorMetadata,
                                   KvConnectorWorker, LeaderTransferRequest, SchedulerRequirement, WorkerTransferRequest,
                                   layer_separate_auto)
from dynamo_b200.physical import BlockDimension
from oracle import oracle as O

NB, NL, PAGE, HEADS, HD = 12, 3, 16, 2, 8


def test_wire_formats_match_serde_json():
    u = str(uuid.UUID(int=7))
    op = WorkerTransferRequest("req-1", u, STORE, SCHEDULED)
    assert op.to_json() == {"request_id": "req-1", "uuid": u, "transfer_type": "Store", "request_type": "Scheduled"}
    md = ConnectorMetadata(3)
    md.create_slot("req-1", 2)
    md.add_operations([op])
    wire = json.loads(md.to_bytes())
    assert wire == {"iteration": 3, "new_slots": [{"request_id": "req-1", "expected_immediate_ops": 2}],
                    "operations": [op.to_json()]}
    back = ConnectorMetadata.from_bytes(md.to_bytes())
    assert back.iteration == 3 and back.operations[0] == op
    # externally tagged enum variants of SchedulerRequirement
    assert SchedulerRequirement("IterationComplete", 4).to_json() == {"IterationComplete": 4}
    assert SchedulerRequirement("LayerComplete", 4, 2).to_json() == {"LayerComplete": [2, 4]}
    assert SchedulerRequirement("LayerNameComplete", 4, "l0").to_json() == {"LayerNameComplete": ["l0", 4]}
    for r in (SchedulerRequirement("IterationComplete", 4), SchedulerRequirement("LayerComplete", 4, 2)):
        assert SchedulerRequirement.from_json(r.to_json()) == r
    # BlockTransferRequest: tuple list as arrays, connector_req skipped when None (skip_serializing_if)
    b = BlockTransferRequest(DEVICE, HOST, [(0, 3), (5, 1)])
    assert b.to_json() == {"from_pool": "Device", "to_pool": "Host", "blocks": [[0, 3], [5, 1]]}
    b2 = BlockTransferRequest(HOST, DEVICE, [(1, 2)], LeaderTransferRequest("r", u, SchedulerRequirement("IterationComplete", 1), SCHEDULED))
    assert BlockTransferRequest.from_json(json.loads(json.dumps(b2.to_json()))) == b2
    with pytest.raises(ValueError):
        WorkerTransferRequest.from_json({"request_id": "x", "uuid": u, "transfer_type": "Copy", "request_type": "Scheduled"})
    with pytest.raises(ValueError):
        BlockTransferRequest.from_json({"from_pool": "Tape", "to_pool": "Host", "blocks": []})


def test_layer_separate_auto_detection():
    assert layer_separate_auto([2, 1024, 16, 8, 128], 1024) == BlockDimension.BlockIsSecondDim   # vLLM flash-attn layout
    assert layer_separate_auto([1024, 2, 16, 8, 128], 1024) == BlockDimension.BlockIsFirstDim
    assert layer_separate_auto([4096, 2, 16, 8, 128], 1024) == BlockDimension.BlockIsFirstDim    # shape[0] >= num_blocks
    with pytest.raises(ValueError):
        layer_separate_auto([7], 4)


def make_worker(host_blocks=8):
    caches = [(f"model.layers.{l}.attn", torch.zeros(2, NB, PAGE, HEADS, HD, dtype=torch.bfloat16)) for l in range(NL)]
    w = KvConnectorWorker(None, "worker-0", host_blocks=host_blocks)
    w.register_kv_caches(NB, PAGE, 0, 2, caches, [0] * NL)
    return w, caches


def fill_sequential(caches):
    twin = O.Layout(O.LW, NB, NL, 2, PAGE, HEADS * HD, 2, block_dim=O.BLOCK_IS_SECOND_DIM,
                    bases=[t.data_ptr() for _, t in caches])
    twin.fill_blocks(range(NB), -1)
    return twin


def test_registration_rules():
    w, caches = make_worker()
    assert w.device_config.outer_dim == 2 and w.device_config.inner_dim == HEADS * HD and w.device_config.num_layers == NL
    with pytest.raises(RuntimeError):                      # worker.rs:139-142
        w.register_kv_caches(NB, PAGE, 0, 2, caches, [0] * NL)
    w2 = KvConnectorWorker()
    with pytest.raises(AssertionError):                    # worker.rs:144-148
        w2.register_kv_caches(NB, PAGE, 0, 2, caches, [0])
    w.close()


def test_store_flow_offload_after_last_layer_and_get_finished():
    """One decode iteration: a Store (offload) operation is announced in the metadata, becomes eligible only when the
    last layer was saved and its iteration completed, moves Device -> Host, and the request finishes afterwards."""
    w, caches = make_worker()
    twin = fill_sequential(caches)
    u = str(uuid.uuid4())
    md = ConnectorMetadata(1)
    md.create_slot("req-A", 0)
    md.add_operations([WorkerTransferRequest("req-A", u, STORE, SCHEDULED)])
    w.bind_connector_metadata(md.to_bytes())
    assert w.has_slot("req-A") and w.slots["req-A"].operations == []          # stores are deferred (worker.rs:300)
    # the leader's transfer request arrives early; it must wait for IterationComplete(1)
    w.handle_block_transfer(BlockTransferRequest(DEVICE, HOST, [(2, 0), (7, 1)],
                                                 LeaderTransferRequest("req-A", u, SchedulerRequirement("IterationComplete", 1), SCHEDULED)).to_json())
    for l in range(NL):
        w.save_kv_layer(caches[l][0])
        assert (len(w.slots["req-A"].operations) == 1) == (l == NL - 1)       # enqueued on the LAST layer (worker.rs:329-350)
    assert not w.is_complete("req-A")
    assert w.get_finished(["req-A"]) == (set(), set())                        # not finished: transfer has not run
    w.clear_connector_metadata()                                              # iteration 1 complete -> requirement met
    off, on = w.get_finished([])
    assert off == {"req-A"} and on == set() and not w.has_slot("req-A")
    host = O.Layout(O.FC, 8, NL, 2, PAGE, HEADS * HD, 2, bases=[w._host_mem.data_ptr()])
    assert host.block_checksum(0) == twin.block_checksum(2) and host.block_checksum(1) == twin.block_checksum(7)
    assert w.get_finished(["never-started"]) == (set(), set())                # unknown ids are ignored (worker.rs:383-389)
    w.close()


def test_load_flow_onboard_is_immediate_and_tracked():
    w, caches = make_worker()
    host = O.Layout(O.FC, 8, NL, 2, PAGE, HEADS * HD, 2, bases=[w._host_mem.data_ptr()])
    host.fill_blocks(range(8), -1)
    u1, u2 = str(uuid.uuid4()), str(uuid.uuid4())
    md = ConnectorMetadata(1)
    md.create_slot("req-B", 1)
    md.add_operations([WorkerTransferRequest("req-B", u1, LOAD, SCHEDULED)])
    w.bind_connector_metadata(md.to_bytes())
    assert w.slots["req-B"].operations == [u1] and "req-B" in w.maybe_finished_onboarding   # loads enqueue at bind
    assert w.get_finished([]) == (set(), set())
    w.handle_block_transfer(BlockTransferRequest(HOST, DEVICE, [(3, 5)], LeaderTransferRequest("req-B", u1, None, SCHEDULED)))
    # an Immediate operation was never announced: it is recorded on arrival (record_operation, scheduler.rs:258-263)
    w.handle_block_transfer(BlockTransferRequest(HOST, DEVICE, [(4, 6)], LeaderTransferRequest("req-B", u2, None, IMMEDIATE)))
    off, on = w.get_finished([])
    assert on == {"req-B"} and off == set()
    dev = O.Layout(O.LW, NB, NL, 2, PAGE, HEADS * HD, 2, block_dim=O.BLOCK_IS_SECOND_DIM, bases=[t.data_ptr() for _, t in caches])
    assert dev.block_checksum(5) == host.block_checksum(3) and dev.block_checksum(6) == host.block_checksum(4)
    w.clear_connector_metadata()
    w.close()


def test_iteration_mismatch_and_duplicate_slot_are_rejected():
    w, _ = make_worker(host_blocks=0)
    md = ConnectorMetadata(5)                                                  # worker starts at iteration 1 (worker.rs:251-255)
    with pytest.raises(AssertionError):
        w.bind_connector_metadata(md.to_bytes())
    w2, _ = make_worker(host_blocks=0)
    md = ConnectorMetadata(1)
    md.create_slot("x", 0)
    w2.bind_connector_metadata(md.to_bytes())
    w2.clear_connector_metadata()
    md2 = ConnectorMetadata(2)
    md2.create_slot("x", 0)
    with pytest.raises(AssertionError):                                        # "slot already exists" (worker.rs:267-270)
        w2.bind_connector_metadata(md2.to_bytes())
    w.close()
    w2.close()


def test_mla_shaped_caches_fall_back_to_outer_dim_one():
    """distributed/worker.rs:554-578: when the candidate K/V axis is > 2 the tensor has no K/V split (MLA:
    [n_blocks, page, latent]) -> outer_dim = 1, inner_dim from every dim after n_blocks."""
    latent = 576
    caches = [(f"l{l}", torch.zeros(NB, PAGE, latent, dtype=torch.bfloat16)) for l in range(NL)]
    w = KvConnectorWorker(None, "mla")
    w.register_kv_caches(NB, PAGE, 0, 2, caches, [0] * NL)
    c = w.device_config
    assert (c.outer_dim, c.inner_dim, c.num_layers, c.num_blocks) == (1, latent, NL, NB)
    w.close()
    # explicit dims (what Python passes via KvTensorLayout) win over inference
    w = KvConnectorWorker(None, "explicit")
    w.register_kv_caches(NB, PAGE, 0, 2, caches, [0] * NL, outer_dim=1, inner_dim=latent)
    assert (w.device_config.outer_dim, w.device_config.inner_dim) == (1, latent)
    w.close()


def test_trtllm_worker_one_fully_contiguous_tensor_and_its_protocol():
    """trtllm_worker.rs:222-380,526-553: one FullyContiguous tensor, loads enqueued by start_load_kv, stores by the last
    save_kv_layer (or one submit_offload_on_event), bytes checked against the oracle."""
    from dynamo_b200.connector import TrtllmKvConnectorWorker
    kv = torch.zeros(NB, NL, 2, PAGE, HEADS * HD, dtype=torch.bfloat16)
    w = TrtllmKvConnectorWorker(None, "rank0", host_blocks=8)
    w.register_kv_caches(NB, PAGE, 0, 2, kv, [0] * NL)
    assert (w.device_config.num_layers, w.device_config.outer_dim, w.device_config.inner_dim) == (NL, 2, HEADS * HD)
    with pytest.raises(RuntimeError):
        w.register_kv_caches(NB, PAGE, 0, 2, kv, [0] * NL)
    dev = O.Layout(O.FC, NB, NL, 2, PAGE, HEADS * HD, 2, bases=[kv.data_ptr()])
    dev.fill_blocks(range(NB), -1)
    host = O.Layout(O.FC, 8, NL, 2, PAGE, HEADS * HD, 2, bases=[w._host_mem.data_ptr()])
    u_store, u_load = str(uuid.uuid4()), str(uuid.uuid4())
    md = ConnectorMetadata(1)
    md.create_slot("r-store", 0)
    md.create_slot("r-load", 0)
    md.add_operations([WorkerTransferRequest("r-store", u_store, STORE, SCHEDULED), WorkerTransferRequest("r-load", u_load, LOAD, SCHEDULED)])
    w.bind_connector_meta(md.to_bytes())
    assert w.slots["r-load"].operations == [] and w.slots["r-store"].operations == []     # nothing enqueued at bind (:320-345)
    w.handle_block_transfer(BlockTransferRequest(DEVICE, HOST, [(3, 0)], LeaderTransferRequest("r-store", u_store, SchedulerRequirement("IterationComplete", 1), SCHEDULED)))
    w.start_load_kv()
    assert w.slots["r-load"].operations == [u_load] and "r-load" in w.maybe_finished_onboarding
    for l in range(NL):
        w.save_kv_layer(l)
    assert w.slots["r-store"].operations == [u_store]                                     # enqueued with the last layer (:357-369)
    w.clear_connector_metadata()
    assert host.block_checksum(0) == dev.block_checksum(3)
    # the load: host block 0 -> device block 9
    w.handle_block_transfer(BlockTransferRequest(HOST, DEVICE, [(0, 9)], LeaderTransferRequest("r-load", u_load, None, SCHEDULED)))
    off, on = w.get_finished(["r-store"])
    assert off == {"r-store"} and on == {"r-load"}
    assert dev.block_checksum(9) == dev.block_checksum(3)
    # second iteration driven by ONE event instead of per-layer calls
    u2 = str(uuid.uuid4())
    md = ConnectorMetadata(2)
    md.create_slot("r2", 0)
    md.add_operations([WorkerTransferRequest("r2", u2, STORE, SCHEDULED)])
    w.bind_connector_meta(md.to_bytes())
    w.handle_block_transfer(BlockTransferRequest(DEVICE, HOST, [(5, 1)], LeaderTransferRequest("r2", u2, SchedulerRequirement("LayerComplete", 2, NL - 1), SCHEDULED)))
    w.submit_offload_on_event(0)
    assert w.slots["r2"].operations == [u2]
    assert w.get_finished(["r2"])[0] == {"r2"}
    assert host.block_checksum(1) == dev.block_checksum(5)
    w.clear_connector_metadata()
    w.close()


def test_scheduled_request_waits_for_the_worker_side_enqueue():
    """scheduler.rs:539-570 (try_prepare_controller): a Scheduled transfer needs BOTH halves -- the leader's request and the
    worker's enqueue of the same uuid -- even when the leader attached no SchedulerRequirement.  Offload enqueues happen on
    the last save_kv_layer (worker.rs:329-350), i.e. after the forward pass produced the blocks."""
    w, caches = make_worker()
    twin = fill_sequential(caches)
    u = str(uuid.uuid4())
    md = ConnectorMetadata(1)
    md.create_slot("req-S", 0)
    md.add_operations([WorkerTransferRequest("req-S", u, STORE, SCHEDULED)])
    w.bind_connector_metadata(md.to_bytes())
    w.handle_block_transfer(BlockTransferRequest(DEVICE, HOST, [(4, 2)], LeaderTransferRequest("req-S", u, None, SCHEDULED)))
    host = O.Layout(O.FC, 8, NL, 2, PAGE, HEADS * HD, 2, bases=[w._host_mem.data_ptr()])
    before = host.block_checksum(2)
    assert len(w._pending) == 1 and not w._inflight                      # requirement=None must NOT fire immediately
    for l in range(NL - 1):
        w.save_kv_layer(caches[l][0])
        assert len(w._pending) == 1 and host.block_checksum(2) == before
    w.save_kv_layer(caches[NL - 1][0])                                     # last layer: the worker half arrives -> runs
    assert not w._pending
    w.clear_connector_metadata()
    assert w.get_finished(["req-S"])[0] == {"req-S"}
    assert host.block_checksum(2) == twin.block_checksum(4)
    w.close()


def test_poll_survives_a_failed_transfer_and_never_double_counts():
    from dynamo_b200.physical import KvbmError
    w, caches = make_worker()

    class _Note:
        def __init__(self, outcome):
            self.outcome, self.polls = outcome, 0

        def is_complete(self):
            self.polls += 1
            if self.outcome == "raise":
                raise KvbmError(1, "transfer aborted (gate timeout)")
            return self.outcome

    md = ConnectorMetadata(1)
    md.create_slot("r", 0)
    w.bind_connector_metadata(md.to_bytes())
    ua, ub, uc = (str(uuid.UUID(int=i)) for i in (1, 2, 3))
    for u_ in (ua, ub, uc):
        w._enqueue(WorkerTransferRequest("r", u_, STORE, SCHEDULED))
    ok, bad, slow = _Note(True), _Note("raise"), _Note(False)
    w._inflight = [(ok, LeaderTransferRequest("r", ua, None, SCHEDULED)), (bad, LeaderTransferRequest("r", ub, None, SCHEDULED)),
                   (slow, LeaderTransferRequest("r", uc, None, SCHEDULED))]
    for _ in range(3):                                                     # repeated polls: no exception escapes, no over-count
        assert not w.is_complete("r")
    assert w.slots["r"].completed == {ua} and list(w.slots["r"].failed) == [ub]
    assert [n for n, _ in w._inflight] == [slow] and bad.polls == 1 and ok.polls == 1
    assert len(w.failures) == 1 and w.failures[0][:2] == ("r", ub)
    slow.outcome = True
    assert w.is_complete("r")                                              # a failed op no longer blocks the request from finishing
    w.clear_connector_metadata()
    w.close()


def test_worker_id_is_stable_across_processes():
    """LayoutHandle carries worker_id (manager/handle.rs:16-50): it has to be the same number in every process."""
    import subprocess
    import sys
    from dynamo_b200.connector import stable_worker_id
    assert stable_worker_id("7") == 7 and stable_worker_id(9) == 9
    a = stable_worker_id("decode-worker-3")
    assert a == stable_worker_id("decode-worker-3") and a != stable_worker_id("decode-worker-4") and a < 2 ** 48
    out = subprocess.run([sys.executable, "-c", "from dynamo_b200.connector import stable_worker_id as f; print(f('decode-worker-3'))"],
                         capture_output=True, text=True, env={"PYTHONHASHSEED": "12345", "PATH": "/usr/bin:/bin",
                                                              "PYTHONPATH": __import__("os").path.dirname(__import__("os").path.dirname(__file__))})
    assert int(out.stdout.strip()) == a, out.stderr

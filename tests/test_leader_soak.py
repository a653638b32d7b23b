"""Randomised end-to-end soak of the connector: a toy continuous-batching scheduler drives KvConnectorLeader, the leader's
ConnectorMetadata and BlockTransferRequests drive KvConnectorWorker, the worker moves real bytes (host-only TransferManager,
Memcpy strategy).  The toy engine writes into every device block it "computes" a pattern derived from that block's SEQUENCE
HASH, so wrong-block, stale-copy and ordering bugs show up as content that does not match its hash:

  * every block registered in the host tier holds the pattern of the hash it is registered under   (offload correctness)
  * every onboarded device block holds the pattern of the hash it was matched by                   (onboard correctness)
  * nothing leaks: at the end no slot, no pending operation, no in-flight host block is left

Reference flow: lib/bindings/kvbm/src/block_manager/vllm/connector/leader.rs:213-598, worker.rs:236-470,
python side lib/bindings/kvbm/python/kvbm/vllm_integration/connector_leader.py (the calls a vLLM scheduler makes)."""
import json

import numpy as np
import pytest
import torch

from dynamo_b200 import router as R
from dynamo_b200.connector import ConnectorMetadata, KvConnectorWorker
from dynamo_b200.leader import KvbmRequest, KvConnectorLeader, SchedulerOutput
from oracle import oracle as O

NB, NL, PAGE, HEADS, HD, HOSTB = 48, 2, 16, 2, 8, 20
INNER = HEADS * HD


def _hashes(tokens):
    n = len(tokens) // PAGE
    return R.compute_seq_hash_for_block(R.compute_block_hash_for_seq(tokens[:n * PAGE], PAGE)) if n else []


def _pattern(h, layer, outer):
    """int16 words of one (layer, outer) region of a block whose sequence hash is h"""
    seed = (h ^ ((layer * 2 + outer + 1) * 0x9E3779B97F4A7C15)) & 0xFFFFFFFFFFFFFFFF
    return np.random.default_rng(seed).integers(-30000, 30000, PAGE * INNER, dtype=np.int16)


class Harness:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.caches = [(f"model.layers.{l}.attn", torch.zeros(2, NB, PAGE, HEADS, HD, dtype=torch.bfloat16)) for l in range(NL)]
        self.w = KvConnectorWorker(None, "worker-0", host_blocks=HOSTB)
        self.w.register_kv_caches(NB, PAGE, 0, 2, self.caches, [0] * NL)
        self.sent = []
        self.leader = KvConnectorLeader("worker-0", PAGE, HOSTB, self._send)
        self.host = O.Layout(O.FC, HOSTB, NL, 2, PAGE, INNER, 2, bases=[self.w._host_mem.data_ptr()])
        self.free = list(range(NB))
        self.reqs = {}            # rid -> dict(tokens, blocks, computed, phase, budget)
        self.waiting_free = {}    # rid -> blocks kept until the worker reports the request finished sending
        self.next_id = 0
        self.families = [[int(x) for x in self.rng.integers(1, 30000, 6 * PAGE)] for _ in range(3)]
        self.checked_onboards = self.checked_host_blocks = 0

    def _send(self, req):
        self.sent.append(req)
        self.w.handle_block_transfer(json.dumps(req.to_json()))

    # ---- the toy engine -------------------------------------------------------------------------------------------------
    def write_kv(self, tokens, blocks, first_block, last_block):
        hs = _hashes(tokens)
        for i in range(first_block, min(last_block, len(hs))):
            for l, (_, t) in enumerate(self.caches):
                v = t.view(torch.int16)
                for o in range(2):
                    v[o, blocks[i]] = torch.from_numpy(_pattern(hs[i], l, o)).view(PAGE, HEADS, HD)

    def device_block_matches(self, block, h):
        return all(np.array_equal(t.view(torch.int16)[o, block].reshape(-1).numpy(), _pattern(h, l, o))
                   for l, (_, t) in enumerate(self.caches) for o in range(2))

    def forward(self, md_bytes):
        self.w.bind_connector_metadata(md_bytes)
        for name, _ in self.caches:
            self.w.save_kv_layer(name)
        self.w.clear_connector_metadata()

    def settle(self):
        self.w._poll()
        for req in self.sent:
            cr = req.connector_req
            slot = self.w.slots.get(cr.request_id)
            if slot is None or cr.uuid in slot.completed:
                self.leader.transfer_complete(cr.uuid)
        self.sent = [r for r in self.sent if r.connector_req.uuid in self.leader._ops]

    # ---- the toy scheduler ----------------------------------------------------------------------------------------------
    def admit(self):
        fam = self.families[int(self.rng.integers(0, len(self.families)))]
        shared = int(self.rng.integers(0, 5)) * PAGE
        tail = [int(x) for x in self.rng.integers(30000, 60000, int(self.rng.integers(1, 3 * PAGE)))]
        tokens = fam[:shared] + tail
        need = -(-len(tokens) // PAGE) + 1                         # one spare block for the first decoded tokens
        if len(self.free) < need:
            return None
        rid = f"r{self.next_id}"
        self.next_id += 1
        self.leader.create_slot(KvbmRequest(rid), tokens)
        n_ext, is_async = self.leader.get_num_new_matched_tokens(rid, len(tokens), 0)
        assert n_ext % PAGE == 0 and n_ext < len(tokens) and is_async == (n_ext > 0)
        blocks = [self.free.pop() for _ in range(need)]
        k = n_ext // PAGE
        if k:
            self.leader.update_state_after_alloc(rid, blocks[:k], n_ext)
            self.leader.update_state_after_alloc(rid, blocks[k:], 0)
        else:
            self.leader.update_state_after_alloc(rid, blocks, 0)
        self.reqs[rid] = dict(tokens=tokens, blocks=blocks, computed=n_ext, phase="onboarding" if k else "new",
                              budget=int(self.rng.integers(0, 2 * PAGE)), onboarded=k)
        return rid

    def step(self):
        if self.rng.integers(0, 3) == 0 or not self.reqs:
            self.admit()
        so = SchedulerOutput()
        for rid, r in self.reqs.items():
            if r["phase"] == "onboarding":
                continue                                            # vLLM does not list it until the load has finished
            if r["phase"] == "new":
                if self.rng.integers(0, 5) == 0:
                    continue                                        # not scheduled this iteration (stays Initialized)
                sched = len(r["tokens"]) - r["computed"]
                so.add_new_request(rid, r["tokens"], r["blocks"], r["computed"])
                so.add_num_scheduled_tokens({rid: sched})
                self.write_kv(r["tokens"], r["blocks"], r["computed"] // PAGE, len(r["tokens"]) // PAGE)
                r["computed"], r["phase"] = len(r["tokens"]), "decode"
            elif r["phase"] == "decode":
                if self.rng.integers(0, 6) == 0:
                    continue                                        # skipped this iteration
                tok = int(self.rng.integers(60000, 90000))
                new_blocks = []
                if len(r["tokens"]) + 1 > len(r["blocks"]) * PAGE:
                    if not self.free:
                        continue
                    new_blocks = [self.free.pop()]
                before = len(r["tokens"])
                r["tokens"].append(tok)
                r["blocks"].extend(new_blocks)
                so.add_cached_request(rid, False, [tok], new_blocks, before)
                so.add_num_scheduled_tokens({rid: 1})
                self.write_kv(r["tokens"], r["blocks"], before // PAGE, len(r["tokens"]) // PAGE)
                r["computed"] = len(r["tokens"])
                r["budget"] -= 1
        md = self.leader.build_connector_metadata(so)
        ConnectorMetadata.from_bytes(md)                            # well-formed JSON in the reference's shape
        self.forward(md)
        self.settle()
        # finished decodes
        done = [rid for rid, r in self.reqs.items() if r["phase"] == "decode" and r["budget"] <= 0 and self.rng.integers(0, 2)]
        for rid in done:
            r = self.reqs.pop(rid)
            assert self.leader.request_finished(rid, r["blocks"]) is True
            self.waiting_free[rid] = r["blocks"]
        sending, recving = self.w.get_finished(done)
        for rid in sending:
            self.free.extend(self.waiting_free.pop(rid))
        for rid in recving:                                          # onboarding finished: the blocks must hold the matched KV
            r = self.reqs[rid]
            hs = _hashes(r["tokens"])
            for i in range(r["onboarded"]):
                assert self.device_block_matches(r["blocks"][i], hs[i]), (rid, i)
                self.checked_onboards += 1
            r["phase"] = "new"
        self.settle()
        self.check_host_tier()

    def check_host_tier(self):
        lh = self.leader.host
        assert len(lh.free) + len(lh.registered) + len(lh.in_flight) == HOSTB
        assert len(set(lh.free) | set(lh.registered.values()) | lh.in_flight) == HOSTB      # no block in two places
        for h, hb in lh.registered.items():
            for l in range(NL):
                for o in range(2):
                    got = np.frombuffer(self.host.region_bytes(hb, l, o), dtype=np.int16)
                    assert np.array_equal(got, _pattern(h, l, o)), ("host block does not hold the KV of its hash", hb, l, o)
            self.checked_host_blocks += 1

    def drain(self):
        for _ in range(400):
            if not self.reqs and not self.waiting_free:
                break
            for r in self.reqs.values():
                r["budget"] = 0
            self.next_admit = False
            # no new admissions while draining
            saved, self.admit = self.admit, lambda: None
            try:
                self.step()
            finally:
                self.admit = saved
        assert not self.reqs and not self.waiting_free, (list(self.reqs), list(self.waiting_free))


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_schedules_keep_host_tier_content_and_state_machines_consistent(seed):
    hx = Harness(seed)
    for _ in range(120):
        hx.step()
    hx.drain()
    assert hx.leader.slots == {} and hx.leader.inflight_requests == set() and hx.leader._ops == {}
    assert hx.w.slots == {} and not hx.w._pending and not hx.w._inflight
    assert sorted(hx.free) == list(range(NB))
    assert not hx.leader.host.in_flight
    assert hx.checked_host_blocks > 50 and hx.checked_onboards > 0, (hx.checked_host_blocks, hx.checked_onboards)
    hx.w.close()

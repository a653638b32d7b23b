"""Connector worker on a GPU: per-layer events become device-side ready flags, ONE gated transfer launch streams the
layers to the host pool while the "forward pass" is still running, and nothing blocks the host
(contrast: worker.rs:341 `event_sync_blocking` on the last layer)."""
import uuid

import numpy as np
import pytest
import torch

from dynamo_b200.connector import (DEVICE, HOST, SCHEDULED, STORE, BlockTransferRequest, ConnectorMetadata, KvConnectorWorker,
                                   LeaderTransferRequest, SchedulerRequirement, WorkerTransferRequest)
from dynamo_b200.physical import TransferOptions
from oracle import oracle as O

pytestmark = pytest.mark.gpu
NB, NL, PAGE, HEADS, HD = 32, 6, 16, 8, 128


def test_layer_streamed_store_without_host_sync():
    caches = [(f"layers.{l}", torch.zeros(2, NB, PAGE, HEADS, HD, dtype=torch.bfloat16, device="cuda")) for l in range(NL)]
    events = [torch.cuda.Event() for _ in range(NL)]
    main = torch.cuda.current_stream()
    for e in events:
        e.record(main)                      # materialise the handles (vLLM creates them once, connector_worker.py:150-160)
    # warm every torch kernel the "forward pass" below uses BEFORE a gated transfer starts spinning (lazy module loading)
    warm = torch.randint(-30000, 30000, (8,), dtype=torch.int16, device="cuda").view(torch.bfloat16)
    caches[0][1].view(-1)[:8].copy_(warm)
    caches[0][1].zero_()
    torch.cuda.synchronize()
    w = KvConnectorWorker(None, "gpu-worker", host_blocks=16)
    w.register_kv_caches(NB, PAGE, 0, 2, caches, [e.cuda_event for e in events])
    u = str(uuid.uuid4())
    md = ConnectorMetadata(1)
    md.create_slot("req", 0)
    md.add_operations([WorkerTransferRequest("req", u, STORE, SCHEDULED)])
    w.bind_connector_metadata(md.to_bytes())
    blocks = [(3, 0), (9, 1), (30, 2), (17, 3)]
    # the gated transfer is launched BEFORE the forward pass: it may read layer l only after that layer's flag
    opts = TransferOptions(layer_ready_flags=w.ready_flags_ptr(), epoch=w._epoch, max_ctas=8, gate_timeout_ms=20000)
    w.handle_block_transfer(BlockTransferRequest(DEVICE, HOST, blocks, LeaderTransferRequest("req", u, None, SCHEDULED)), opts)
    assert not w.is_complete("req") or len(w.slots["req"].operations) == 0
    g = torch.Generator(device="cuda").manual_seed(5)
    try:
        for l, (name, t) in enumerate(caches):   # the "forward pass": layer l's KV is produced, its event recorded, saved
            t.copy_(torch.randint(-30000, 30000, t.shape, dtype=torch.int16, device="cuda", generator=g).view(torch.bfloat16))
            events[l].record(main)
            w.save_kv_layer(name, t)             # returns immediately: no cuEventSynchronize
    finally:
        w.clear_connector_metadata()
    for _ in range(2000):
        off, on = w.get_finished(["req"])
        if off:
            break
        torch.cuda._sleep(100000)
        torch.cuda.synchronize()
    assert off == {"req"}
    dev = O.Layout(O.LW, NB, NL, 2, PAGE, HEADS * HD, 2, block_dim=O.BLOCK_IS_SECOND_DIM)
    for hb, (_, t) in zip(dev.buffers, caches):
        hb[:] = t.view(torch.uint8).reshape(-1).cpu().numpy()
    host = O.Layout(O.FC, 16, NL, 2, PAGE, HEADS * HD, 2, bases=[w._host_mem.data_ptr()])
    for s, d in blocks:
        assert host.block_checksum(d) == dev.block_checksum(s)
    w.close()

"""Decode-side PULL of the KV hand-off (NIXL-READ direction, components/src/dynamo/vllm/handlers.py:2076-2083), optionally
with the fp8 -> bf16 up-cast fused on the RECEIVER: every decode rank maps rank 0's pool (CUDA IPC) and runs the block-table
kernel itself -- loads cross NVLink (fp8: half the bytes of the bf16 it writes locally), stores stay local.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 5 --master-addr 127.0.0.1 benchmarks/pull_bench.py --cast fp8

BASELINE configs[2] geometry by default (Llama-3-8B, 4k ctx, 256 blocks).  Device-timed (CUDA events on each puller's
stream), max over ranks.  Compare with `bench.py --gpus 5 --cast fp8` (push: the source casts, bf16 crosses NVLink).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamo_b200 import kernels as K  # noqa: E402
from dynamo_b200.physical import BlockDimension, LayoutConfig, StorageKind, TransferManager, TransferOptions  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cast", choices=["none", "fp8"], default="fp8")
ap.add_argument("--blocks", type=int, default=256)
ap.add_argument("--pool", type=int, default=1024)
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--out", default="")
a = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
assert world >= 2, "needs one source rank and at least one puller"
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
NL, NO, PAGE, INNER, NB, n = a.layers, 2, 16, 1024, a.pool, a.blocks
src_w = 1 if a.cast == "fp8" else 2
cast = K.CastMode.FP8E4M3_TO_BF16 if a.cast == "fp8" else K.CastMode.NONE
mgr = TransferManager(device=local, worker_id=rank + 1)
cfg_src = LayoutConfig(NB, NL, NO, PAGE, INNER, dtype_width_bytes=src_w, allow_fp8=src_w == 1)
cfg_dst = LayoutConfig(NB, NL, NO, PAGE, INNER, dtype_width_bytes=2)

blob = b""
if rank == 0:
    src = [torch.empty(NO * NB * PAGE * INNER * src_w, dtype=torch.uint8, device=dev).random_(0, 256) for _ in range(NL)]
    h_src = mgr.register_layer_separate(cfg_src, [b.data_ptr() for b in src], [b.numel() for b in src],
                                        BlockDimension.BlockIsSecondDim, StorageKind.Device, local)
    blob = mgr.export_metadata(h_src)
torch.cuda.synchronize()
box = [blob]
dist.broadcast_object_list(box, src=0)
sid = np.random.default_rng(10).permutation(NB)[:n].astype(np.uint64)
did = np.random.default_rng(100 + rank).permutation(NB)[:n].astype(np.uint64)
stream = torch.cuda.Stream()
sp = int(stream.cuda_stream)
ms = torch.zeros(1, device=dev)
probe = None
if rank > 0:
    h_remote = mgr.import_metadata(box[0])
    dst = [torch.zeros(NO * NB * PAGE * INNER * 2, dtype=torch.uint8, device=dev) for _ in range(NL)]
    h_dst = mgr.register_layer_separate(cfg_dst, [b.data_ptr() for b in dst], [b.numel() for b in dst],
                                        BlockDimension.BlockIsSecondDim, StorageKind.Device, local)
    opts = TransferOptions(cast_mode=cast, cuda_stream=sp)   # caller's stream: launches queue back to back, events time them

    def step():
        mgr.execute_transfer(h_remote, sid, h_dst, did, opts)
    for _ in range(a.warmup):
        step()
    stream.synchronize()
dist.barrier()
if rank > 0:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        step()
    e1.record(stream)
    stream.synchronize()
    ms[0] = e0.elapsed_time(e1) / a.steps
    probe = int(dst[NL - 1].view(NO, NB, -1)[:, int(did[7])].sum(dtype=torch.int64).item())
dist.barrier()
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
probes = [None] * world
dist.all_gather_object(probes, probe)
if rank == 0:
    t = float(ms.item())
    delivered = n * NL * NO * PAGE * INNER * 2 * (world - 1)
    wire = n * NL * NO * PAGE * INNER * src_w * (world - 1)
    # the pulled block's bf16 bytes summed must agree across pullers (same source block, same cast)
    want = None
    if a.cast == "none":
        want = int(src[NL - 1].view(NO, NB, -1)[:, int(sid[7])].sum(dtype=torch.int64).item())
    ok = len(set(probes[1:])) == 1 and (want is None or probes[1] == want)
    line = {"bench": "decode-side pull" + (" + receiver-side fp8->bf16 up-cast" if a.cast == "fp8" else ""), "n_gpus": world,
            "pullers": world - 1, "ms_per_step": round(t, 5), "delivered_gbs_all_destinations": round(delivered / t / 1e6, 2),
            "nvlink_egress_gbs_of_source": round(wire / t / 1e6, 2), "bytes_on_wire_per_destination": wire // (world - 1),
            "bytes_delivered_per_destination": delivered // (world - 1), "steps": a.steps, "warmup": a.warmup, "probe_ok": ok}
    print(json.dumps(line), flush=True)
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump(line, open(a.out, "w"), indent=1)
dist.barrier()
mgr.close()
dist.destroy_process_group()

"""BASELINE configs[4] (+[2]): 1 prefill -> N-1 decode fan-out bandwidth sweep over context length, one process per
GPU (torchrun), pools mapped through CUDA IPC.  Mixtral-8x7B KV geometry (= Llama-3-8B: 32 layers, 8 KV heads,
head_dim 128, bf16 -> 32 KiB regions).  For every ctx in --ctx and every mode in {distinct, replicate}
(and prefix-hit rates: only the non-hit suffix ceil((1-h)*blocks) is moved) prints one JSON line:
GB/s of destination bytes summed over destinations, timed on the source GPU with CUDA events (median of --iters).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        benchmarks/fanout_sweep.py --ctx 1024,4096,16384,65536 --out gpurun_out/fanout_n8.jsonl
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamo_b200 import kernels as K  # noqa: E402
from dynamo_b200.disagg import HandoffGroup  # noqa: E402
from dynamo_b200.physical import BlockDimension, LayoutConfig, StorageKind, TransferManager  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ctx", default="1024,4096,16384,65536")
ap.add_argument("--hit", default="0", help="comma list of prefix-hit rates")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--cast", default="none", choices=["none", "fp8"])
ap.add_argument("--ctas", default="0", help="comma list of CTA caps (0 = whole GPU)")
ap.add_argument("--out", default="gpurun_out/fanout.jsonl")
a = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
NL, OUTER, PAGE, INNER = 32, 2, 16, 1024
cast = 1 if a.cast == "fp8" else 0
dst_region = PAGE * INNER * 2
src_region = dst_region // 2 if cast else dst_region
ctxs = [int(x) for x in a.ctx.split(",")]
max_blocks = max(ctxs) // PAGE
pool_blocks = max(1024, max_blocks + max_blocks // 4)
n_dst = world - 1

mgr = TransferManager(device=local, worker_id=rank + 1)
grp = HandoffGroup(mgr, rank, world, "fanout")


def pool(region, dtype_bytes, fp8=False):
    bufs = [torch.empty(OUTER * pool_blocks * region, dtype=torch.uint8, device=dev) for _ in range(NL)]
    cfg = LayoutConfig(pool_blocks, NL, OUTER, PAGE, INNER, dtype_width_bytes=dtype_bytes, allow_fp8=fp8)
    h = mgr.register_layer_separate(cfg, [b.data_ptr() for b in bufs], [b.numel() for b in bufs],
                                    BlockDimension.BlockIsSecondDim, StorageKind.Device, local)
    return bufs, h


if rank == 0:
    src_bufs, h_src = pool(src_region, 1 if cast else 2, bool(cast))
    for b in src_bufs:
        b.random_(0, 256)
    grp.publish(None)
else:
    dst_bufs, h_dst = pool(dst_region, 2)
    for b in dst_bufs:
        b.zero_()
    grp.publish(h_dst)
torch.cuda.synchronize()
dist.barrier()

lines = []
if rank == 0:
    def desc(h, region):
        t = torch.tensor([mgr.memory_region(h, 0, l, 0)[0] for l in range(NL)], dtype=torch.int64, device=dev)
        return t, K.PagedLayout(t.data_ptr(), region, region * pool_blocks, region, NL, OUTER, pool_blocks)
    keep = []
    t_s, d_src = desc(h_src, src_region)
    dd = []
    for r in range(1, world):
        t, d = desc(grp.remote[r], dst_region)
        keep.append(t)
        dd.append(d)
    stream = torch.cuda.Stream()
    sp = int(stream.cuda_stream)
    for ctx in ctxs:
        for hit in [float(x) for x in a.hit.split(",")]:
            n = max(1, math.ceil((1.0 - hit) * (ctx // PAGE)))
            for mode in ("distinct", "replicate"):
                for ctas in [int(x) for x in a.ctas.split(",")]:
                    rng = np.random.default_rng(ctx + 7)
                    s_shared = torch.from_numpy(rng.permutation(pool_blocks)[:n].astype(np.int32)).to(dev)
                    dsts = []
                    for d in range(n_dst):
                        s_ids = s_shared if mode == "replicate" else torch.from_numpy(rng.permutation(pool_blocks)[:n].astype(np.int32)).to(dev)
                        d_ids = torch.from_numpy(rng.permutation(pool_blocks)[:n].astype(np.int32)).to(dev)
                        keep += [s_ids, d_ids]
                        dsts.append(K.PagedDst(dd[d], s_ids.data_ptr(), d_ids.data_ptr(), 0, 0))
                    opts = K.PagedCopyOpts(max_ctas=ctas)
                    ts = []
                    with torch.cuda.stream(stream):
                        for i in range(a.iters + 2):
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record(stream)
                            K.check(K.paged_copy(d_src, dsts, n, 0, NL, cast, opts, sp))
                            e1.record(stream)
                            e1.synchronize()
                            if i >= 2:
                                ts.append(e0.elapsed_time(e1))
                    ms = float(np.median(ts))
                    dst_bytes = n * NL * OUTER * dst_region * n_dst
                    src_bytes = n * NL * OUTER * src_region * (1 if mode == "replicate" else n_dst)
                    line = {"n_gpus": world, "fan_out": n_dst, "ctx": ctx, "prefix_hit": hit, "blocks_moved": n, "mode": mode, "cast": a.cast,
                            "max_ctas": ctas, "ms": round(ms, 4), "nvlink_egress_gbs": round(dst_bytes / ms / 1e6, 1),
                            "per_destination_gbs": round(dst_bytes / n_dst / ms / 1e6, 1),
                            "source_hbm_read_gbs": round(src_bytes / ms / 1e6, 1), "frac_of_770": round(dst_bytes / ms / 1e6 / 770.0, 3)}
                    print(json.dumps(line), flush=True)
                    lines.append(line)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "a") as f:
        for ln in lines:
            f.write(json.dumps(ln) + "\n")
dist.barrier()
# spot check on every destination: the last transfer's first block equals the source bytes (raw copy only)
mgr.close()
dist.destroy_process_group()

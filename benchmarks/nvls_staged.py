"""Staged NVLS broadcast at full size (BASELINE configs[4] "same payload to all 7", but every receiver scatters into ITS OWN
block table): rank-less, one process drives all visible GPUs.

  root (GPU 0)   multicast-writes the request's 256 source blocks ONCE into blocks 0..255 of the staging pool every GPU bound
                 to the multicast group, setting per-layer flags in every receiver's memory        (disagg.staged_send)
  receiver d     one gated launch on its own GPU scatters staging -> its own 1024-block pool with its own random block
                 table, layer by layer as the flags arrive, then tells the root the staging pool is free   (staged_receive)

Timing: K steps back to back (the root waits on the device for every receiver's `free` flag before it overwrites the
staging pool), host wall clock from first launch to all streams idle, / K.  Every receiver pool is compared with the
source blocks afterwards.  Compare with bench.py --replicate (unicast replicate) and --replicate --nvls (identical tables).
    python benchmarks/nvls_staged.py --steps 20 --out gpurun_out/r02_nvls_staged.json
"""
import argparse
import json
import os
import sys
import time

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")   # gated receivers spin in ONE launch (see kvbm_kernels.h gate_mode)
import numpy as np  # noqa: E402
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamo_b200.disagg import staged_receive, staged_send  # noqa: E402
from dynamo_b200.physical import BlockDimension, LayoutConfig, MulticastGroup, StorageKind, TransferManager  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--blocks", type=int, default=256)
ap.add_argument("--pool", type=int, default=1024)
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--recv-ctas", type=int, default=32)
ap.add_argument("--out", default="gpurun_out/nvls_staged.json")
a = ap.parse_args()
NL, NO, PAGE, INNER, DT, n, NB = a.layers, 2, 16, 1024, 2, a.blocks, a.pool
REGION = PAGE * INNER * DT
nd = torch.cuda.device_count()
assert nd >= 2
stage_layer = NO * n * REGION
stage_total = NL * stage_layer
pool_layer = NO * NB * REGION


class _Raw:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


for d in range(nd):
    torch.zeros(1, device=f"cuda:{d}")
torch.cuda.set_device(0)
g = MulticastGroup.create(nd, stage_total)
for d in range(nd):
    g.add_device(d)
stage = []
for d in range(nd):
    with torch.cuda.device(d):
        t = torch.as_tensor(_Raw(g.bind_local(d), stage_total), device=f"cuda:{d}")
        t.zero_()
        stage.append(t)
for d in range(nd):
    torch.cuda.synchronize(d)
mc = g.map(0)
stage_cfg = LayoutConfig(n, NL, NO, PAGE, INNER, dtype_width_bytes=DT)
pool_cfg = LayoutConfig(NB, NL, NO, PAGE, INNER, dtype_width_bytes=DT)
root = TransferManager(device=0, worker_id=1)
receivers = list(range(1, nd))
for d in receivers:
    root.enable_peer_access(d)
src = [torch.randint(0, 256, (pool_layer,), dtype=torch.uint8, device="cuda:0") for _ in range(NL)]
h_src = root.register_layer_separate(pool_cfg, [b.data_ptr() for b in src], [pool_layer] * NL, BlockDimension.BlockIsSecondDim, StorageKind.Device, 0)
h_mc = root.register_layer_separate(stage_cfg, [mc + l * stage_layer for l in range(NL)], [stage_layer] * NL, BlockDimension.BlockIsSecondDim,
                                    StorageKind.Device, 0)
free = torch.zeros(nd, dtype=torch.int32, device="cuda:0")
side = torch.cuda.Stream(device="cuda:0")
mgrs, h_stage, h_dst, dsts, ready = {}, {}, {}, {}, {}
for d in receivers:
    with torch.cuda.device(d):
        m = TransferManager(device=d, worker_id=1 + d)
        m.enable_peer_access(0)
        mgrs[d] = m
        h_stage[d] = m.register_layer_separate(stage_cfg, [stage[d].data_ptr() + l * stage_layer for l in range(NL)], [stage_layer] * NL,
                                               BlockDimension.BlockIsSecondDim, StorageKind.Device, d)
        dsts[d] = [torch.zeros(pool_layer, dtype=torch.uint8, device=f"cuda:{d}") for _ in range(NL)]
        h_dst[d] = m.register_layer_separate(pool_cfg, [b.data_ptr() for b in dsts[d]], [pool_layer] * NL, BlockDimension.BlockIsSecondDim,
                                             StorageKind.Device, d)
        ready[d] = torch.zeros(NL, dtype=torch.int32, device=f"cuda:{d}")
for d in range(nd):
    torch.cuda.synchronize(d)
sid = np.random.default_rng(10).permutation(NB)[:n].astype(np.uint64)
dids = {d: np.random.default_rng(100 + d).permutation(NB)[:n].astype(np.uint64) for d in receivers}
ready_ptrs = [ready[d].data_ptr() for d in receivers]
free_ptrs = [free[d:].data_ptr() for d in receivers]
sp = int(side.cuda_stream)


def step(epoch):
    notes = []
    for d in receivers:
        with torch.cuda.device(d):
            notes.append(staged_receive(mgrs[d], h_stage[d], h_dst[d], dids[d], ready[d].data_ptr(), epoch, free_flag=free[d:].data_ptr(),
                                        max_ctas=a.recv_ctas))
    staged_send(root, h_src, sid, h_mc, ready_ptrs, epoch, sp, receiver_free_flags=free_ptrs)
    return notes


def drain(notes):
    for nt in notes:
        nt.wait(60.0)
    side.synchronize()


epoch = 0
for _ in range(3):
    epoch += 1
    drain(step(epoch))
t0 = time.perf_counter()
notes = []
for _ in range(a.steps):
    epoch += 1
    notes = step(epoch)
drain(notes)
for d in range(nd):
    torch.cuda.synchronize(d)
ms = 1e3 * (time.perf_counter() - t0) / a.steps
# one isolated step (latency a single request sees)
lat = []
for _ in range(5):
    epoch += 1
    t1 = time.perf_counter()
    drain(step(epoch))
    lat.append(1e3 * (time.perf_counter() - t1))
ok = True
rows = torch.as_tensor(sid.astype(np.int64), device="cuda:0")
for d in receivers:
    drows = torch.as_tensor(dids[d].astype(np.int64), device=f"cuda:{d}")
    for l in range(NL):
        want = src[l].view(NO, NB, REGION)[:, rows].to(f"cuda:{d}")
        ok = ok and bool(torch.equal(dsts[d][l].view(NO, NB, REGION)[:, drows], want))
payload = n * NL * NO * REGION
res = {"what": "staged NVLS broadcast: multicast into a group-bound staging pool + gated local scatter into each receiver's own block table",
       "gpus": nd, "receivers": len(receivers), "payload_bytes_per_receiver": payload, "steps": a.steps,
       "ms_per_step_pipelined": round(ms, 4), "ms_single_step_median": round(float(np.median(lat)), 4),
       "delivered_gbs_all_receivers": round(payload * len(receivers) / (ms * 1e-3) / 1e9, 1),
       "source_egress_gbs": round(payload / (ms * 1e-3) / 1e9, 1), "receiver_ctas": a.recv_ctas, "all_blocks_bit_exact": ok}
print(json.dumps(res), flush=True)
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
sys.stdout.flush()
os._exit(0 if ok else 1)

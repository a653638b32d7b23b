"""GPU baselines BESIDE ours for the cfg2 hand-off (SURVEY.md §8d "Reference beside it"), same tables, same pools:

  ours_paged        TransferManager.execute_transfer (host block ids -> one block-table kernel)            [e2e]
  ref_k1_flow       the reference's execute_fc_lw_vectorized flow (kvbm-physical executor/cuda.rs:234-327) with ITS OWN
                    kernel recompiled (oracle/_ref): build 2 x nb*nl*no addresses on the host, upload both tables,
                    launch K1, block the host on an event
  ours_k1_flow      the same flow through OUR drop-in K1 symbol (what a swapped libkvbm_kernels.so gives with no other change)
  memcpy_per_chunk  one cudaMemcpyAsync per chunk (v1 D2D path / UCX cuda_ipc behaviour, block/transfer/cuda.rs:299-391)
  memcpy_peer_whole one cudaMemcpyPeerAsync of the same number of bytes, contiguous (the DMA-engine ceiling)

Wall-clock per transfer (host work included, stream synchronised), medians.  Destination pool on GPU 1 when visible.
    python benchmarks/baselines.py --out gpurun_out/baselines.json
The grouped-ncclBcast baseline is the separate binary benchmarks/nccl_bcast_baseline (needs no Python).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamo_b200 import kernels as K  # noqa: E402
from dynamo_b200.physical import BlockDimension, LayoutConfig, StorageKind, TransferManager  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--blocks", type=int, default=256)
ap.add_argument("--pool", type=int, default=1024)
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--iters", type=int, default=15)
ap.add_argument("--out", default="gpurun_out/baselines.json")
a = ap.parse_args()

NL, NO, REGION, NB, n = a.layers, 2, 32768, a.pool, a.blocks
torch.cuda.set_device(0)
peer = torch.cuda.device_count() > 1
ddev = 1 if peer else 0
mgr = TransferManager(device=0)
if peer:
    mgr.enable_peer_access(1)
src = [torch.empty(NO * NB * REGION, dtype=torch.uint8, device="cuda:0").random_(0, 256) for _ in range(NL)]
dst = [torch.zeros(NO * NB * REGION, dtype=torch.uint8, device=f"cuda:{ddev}") for _ in range(NL)]
cfg = LayoutConfig(NB, NL, NO, 16, 1024, dtype_width_bytes=2)
h_src = mgr.register_layer_separate(cfg, [b.data_ptr() for b in src], [b.numel() for b in src], BlockDimension.BlockIsSecondDim, StorageKind.Device, 0)
h_dst = mgr.register_layer_separate(cfg, [b.data_ptr() for b in dst], [b.numel() for b in dst], BlockDimension.BlockIsSecondDim, StorageKind.Device, ddev)
rng = np.random.default_rng(0)
sid = rng.permutation(NB)[:n].astype(np.uint64)
did = rng.permutation(NB)[:n].astype(np.uint64)
stream = torch.cuda.Stream()
sp = int(stream.cuda_stream)
payload = n * NL * NO * REGION

sbase = np.array([b.data_ptr() for b in src], dtype=np.uint64)
dbase = np.array([b.data_ptr() for b in dst], dtype=np.uint64)


def host_tables():
    """The reference's host loop (cuda.rs:258-283): (block, layer, outer) order, 2 x nb*nl*no addresses.  numpy stands in
    for 32 768 `memory_region()` calls -- generous to the reference."""
    o = np.arange(NO, dtype=np.uint64)
    s = (sbase[None, :, None] + sid[:, None, None] * np.uint64(REGION) + o[None, None, :] * np.uint64(REGION * NB)).reshape(-1)
    d = (dbase[None, :, None] + did[:, None, None] * np.uint64(REGION) + o[None, None, :] * np.uint64(REGION * NB)).reshape(-1)
    return s, d


npairs = n * NL * NO
pin_s = torch.empty(npairs, dtype=torch.int64).pin_memory()
pin_d = torch.empty(npairs, dtype=torch.int64).pin_memory()
dev_s = torch.empty(npairs, dtype=torch.int64, device="cuda:0")
dev_d = torch.empty(npairs, dtype=torch.int64, device="cuda:0")
done = torch.cuda.Event()


def k1_flow(launch):
    s, d = host_tables()
    pin_s.numpy()[:] = s.view(np.int64)
    pin_d.numpy()[:] = d.view(np.int64)
    with torch.cuda.stream(stream):
        dev_s.copy_(pin_s, non_blocking=True)
        dev_d.copy_(pin_d, non_blocking=True)
        done.record(stream)
    done.synchronize()                                   # pointers_transfered_event.synchronize() (cuda.rs:324)
    assert launch(dev_s.data_ptr(), dev_d.data_ptr(), REGION, npairs, sp) == 0
    stream.synchronize()                                 # the transfer's completion


ref_path = os.path.join(ROOT, "oracle", "_ref", "libkvbm_kernels_ref.so")
RLIB = None
if os.path.exists(ref_path):
    RLIB = C.CDLL(ref_path)
    RLIB.kvbm_kernels_launch_vectorized_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
HL = C.CDLL(os.path.join(ROOT, "benchmarks", "libstandin.so"))
HL.memcpy_per_chunk.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]


def per_chunk():
    s, d = host_tables()
    assert HL.memcpy_per_chunk(s.ctypes.data, d.ctypes.data, npairs, REGION, sp) == 0
    stream.synchronize()


whole_s = torch.empty(payload, dtype=torch.uint8, device="cuda:0")
whole_d = torch.empty(payload, dtype=torch.uint8, device=f"cuda:{ddev}")


def peer_whole():
    with torch.cuda.stream(stream):
        whole_d.copy_(whole_s, non_blocking=True)
    stream.synchronize()


def ours_paged():
    mgr.execute_transfer(h_src, sid, h_dst, did).wait(60.0)


def check(tag):
    for l in (0, NL - 1):
        got = dst[l].view(NO, NB, REGION)[:, torch.from_numpy(did.astype(np.int64)).to(dst[l].device)].cpu()
        want = src[l].view(NO, NB, REGION)[:, torch.from_numpy(sid.astype(np.int64)).cuda()].cpu()
        assert torch.equal(got, want), tag
    for b in dst:
        b.zero_()
    torch.cuda.synchronize()


def t_ms(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


res = {"payload_bytes": payload, "chunks": npairs, "peer": peer}
cases = [("ours_paged", ours_paged, True), ("ours_k1_flow", lambda: k1_flow(K.lib().kvbm_kernels_launch_vectorized_copy), True),
         ("memcpy_per_chunk", per_chunk, True), ("memcpy_peer_whole", peer_whole, False)]
if RLIB is not None:
    cases.insert(1, ("ref_k1_flow", lambda: k1_flow(RLIB.kvbm_kernels_launch_vectorized_copy), True))
for name, fn, verify in cases:
    fn()
    torch.cuda.synchronize()
    if verify:
        check(name)
    ms = t_ms(fn)
    res[name] = {"ms": ms, "gbs": payload / ms / 1e6}
    print(name, res[name], flush=True)
t0 = time.perf_counter()
for _ in range(20):
    host_tables()
res["host_table_build_ms_numpy"] = (time.perf_counter() - t0) / 20 * 1e3
print(json.dumps(res, indent=1))
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)

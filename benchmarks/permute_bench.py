"""Layout-transforming hand-off, timed: kvbm_kernels_paged_permute (layouts + block tables, no pointer tables) against the flow the
reference's K2 needs (host builds nb*nl*no chunk pointers + nb universal pointers, uploads them, launches
kvbm_kernels_launch_universal_from_block -- tensor_kernels.rs:177-215) and against the plain paged copy of the same bytes.
Llama-3-70B KV geometry by default.  Prints one JSON line; roofline = bytes read + written / measured HBM copy peak."""
import argparse
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dynamo_b200 import kernels as K  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--blocks", type=int, default=64)
p.add_argument("--pool", type=int, default=96)
p.add_argument("--layers", type=int, default=80)
p.add_argument("--heads", type=int, default=8)
p.add_argument("--head-dim", type=int, default=128)
p.add_argument("--page", type=int, default=16)
p.add_argument("--iters", type=int, default=20)
p.add_argument("--peer", type=int, default=-1, help="put one pool on this GPU (the launch stays on GPU 0); -1 = same GPU")
p.add_argument("--peer-side", choices=["dst", "src"], default="dst", help="dst: universal pool on the peer (push); src: engine pool on the peer (pull)")
a = p.parse_args()
nl, no, nt, nh, hd, elem = a.layers, 2, a.page, a.heads, a.head_dim, 2
row, region = hd * elem, nt * nh * hd * elem
nb, n = a.pool, a.blocks
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
ddev = torch.device(f"cuda:{a.peer}") if a.peer >= 0 and a.peer_side == "dst" else dev
sdev = torch.device(f"cuda:{a.peer}") if a.peer >= 0 and a.peer_side == "src" else dev
if a.peer >= 0:
    import ctypes
    rt = ctypes.CDLL("libcudart.so")
    rt.cudaDeviceEnablePeerAccess(a.peer, 0)
peak = 6462.4
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

# engine pool: layer-separate, block-is-second-dim (vLLM), NHD inside a region; universal pool: fully contiguous
op = [torch.randint(0, 256, (no * nb * region,), dtype=torch.uint8, device=sdev) for _ in range(nl)]
torch.cuda.synchronize(sdev)
uni = torch.zeros(nb * nl * no * region, dtype=torch.uint8, device=ddev)
op_base = torch.tensor([t.data_ptr() for t in op], dtype=torch.int64, device=dev)
op_desc = K.PagedLayout(op_base.data_ptr(), region, nb * region, region, nl, no, nb)
bs = nl * no * region
uni_base = torch.tensor([uni.data_ptr() + l * no * region for l in range(nl)], dtype=torch.int64, device=dev)
uni_desc = K.PagedLayout(uni_base.data_ptr(), bs, region, region, nl, no, nb)
rng = np.random.default_rng(0)
sid, did = rng.permutation(nb)[:n].astype(np.int32), rng.permutation(nb)[:n].astype(np.int32)
sid_d, did_d = torch.from_numpy(sid).to(dev), torch.from_numpy(did).to(dev)
stream = torch.cuda.current_stream()
sp = int(stream.cuda_stream)
moved = n * bs


def timed(fn, iters=a.iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def paged():
    K.check(K.paged_permute(K.PermuteSide(op_desc, sid_d.data_ptr(), 4), K.PermuteSide(uni_desc, did_d.data_ptr(), 1), n, 0, nl, nh, nt, row, stream=sp))


def paged_back():
    K.check(K.paged_permute(K.PermuteSide(uni_desc, did_d.data_ptr(), 1), K.PermuteSide(op_desc, sid_d.data_ptr(), 4), n, 0, nl, nh, nt, row, stream=sp))


chunk_tab = torch.empty(n * nl * no, dtype=torch.int64).pin_memory()
uni_tab = torch.empty(n, dtype=torch.int64).pin_memory()
chunk_dev, uni_dev = torch.empty(n * nl * no, dtype=torch.int64, device=dev), torch.empty(n, dtype=torch.int64, device=dev)
op_ptrs = np.array([t.data_ptr() for t in op], dtype=np.int64)


def legacy_kernel_only():
    K.check(K.universal_from_block(uni_dev.data_ptr(), chunk_dev.data_ptr(), n, nh, nl, no, nt, hd, K.TensorDataType.BF16, K.BlockLayout.NHD, sp))


def legacy_flow():
    """what a caller of the reference's K2 does per transfer: build both pointer tables on the host (vectorised here; the
    reference loops in Rust), upload them, launch, and -- like executor/cuda.rs:324 -- wait for the upload before returning."""
    ct = chunk_tab.numpy().reshape(n, nl, no)
    ct[:] = op_ptrs[None, :, None] + (np.arange(no, dtype=np.int64)[None, None, :] * nb + sid.astype(np.int64)[:, None, None]) * region
    uni_tab.numpy()[:] = uni.data_ptr() + did.astype(np.int64) * bs
    chunk_dev.copy_(chunk_tab, non_blocking=True)
    uni_dev.copy_(uni_tab, non_blocking=True)
    legacy_kernel_only()


def plain_copy():
    dd = [K.PagedDst(uni_desc, sid_d.data_ptr(), did_d.data_ptr(), 0, 0)]
    K.check(K.paged_copy(op_desc, dd, n, 0, nl, 0, None, sp))


legacy_flow()
torch.cuda.synchronize()
a_ref = uni.clone()
uni.zero_()
torch.cuda.synchronize(ddev)      # the zero fill runs on the pool's own device: order it before the launch on GPU 0
paged()
torch.cuda.synchronize()
same = bool(torch.equal(a_ref, uni))


def wall(fn, iters=a.iters):
    ts = []
    for _ in range(iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts)


res = {"geometry": {"blocks": n, "layers": nl, "outer": no, "page": nt, "heads": nh, "head_dim": hd, "elem": elem, "MiB_moved": moved / 2**20},
       "peer": a.peer, "peer_side": a.peer_side if a.peer >= 0 else None, "paged_equals_legacy": same}
for name, fn in (("paged_permute_to_universal", paged), ("paged_permute_from_universal", paged_back), ("legacy_k2_kernel_only", legacy_kernel_only),
                 ("plain_paged_copy", plain_copy)):
    ms = timed(fn)
    res[name] = {"ms": round(ms, 4), "GBps_read_plus_write": round(2 * moved / ms / 1e6, 1), "frac_of_hbm_peak": round(2 * moved / ms / 1e6 / peak, 3)}
res["wall_ms_host_to_done"] = {"paged_permute": round(wall(paged), 4), "legacy_k2_flow_tables_h2d_launch": round(wall(legacy_flow), 4)}
print(json.dumps(res))

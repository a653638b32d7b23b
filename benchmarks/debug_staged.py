"""2-GPU debug: does the multicast launch write per-layer flags into the receiver's memory?"""
import os, sys
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamo_b200 import kernels as K
from dynamo_b200.physical import BlockDimension, LayoutConfig, MulticastGroup, StorageKind, TransferManager, TransferOptions

NB, NL, NO, PAGE, INNER, DT = 64, 4, 2, 16, 1024, 2
REGION = PAGE * INNER * DT; PER_LAYER = NO * NB * REGION; TOTAL = NL * PER_LAYER
class _Raw:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
nd = torch.cuda.device_count()
for d in range(nd): torch.zeros(1, device=f"cuda:{d}")
torch.cuda.set_device(0)
g = MulticastGroup.create(nd, TOTAL)
for d in range(nd): g.add_device(d)
pools = []
for d in range(nd):
    with torch.cuda.device(d):
        t = torch.as_tensor(_Raw(g.bind_local(d), TOTAL), device=f"cuda:{d}"); t.zero_(); pools.append(t)
for d in range(nd): torch.cuda.synchronize(d)
mc = g.map(0)
cfg = LayoutConfig(NB, NL, NO, PAGE, INNER, dtype_width_bytes=DT)
root = TransferManager(device=0, worker_id=1)
for d in range(1, nd): root.enable_peer_access(d)
src = [torch.randint(0, 256, (PER_LAYER,), dtype=torch.uint8, device="cuda:0") for _ in range(NL)]
h_src = root.register_layer_separate(cfg, [b.data_ptr() for b in src], [PER_LAYER] * NL, BlockDimension.BlockIsSecondDim, StorageKind.Device, 0)
h_mc = root.register_layer_separate(cfg, [mc + l * PER_LAYER for l in range(NL)], [PER_LAYER] * NL, BlockDimension.BlockIsSecondDim, StorageKind.Device, 0)
ready_local = torch.zeros(NL, dtype=torch.int32, device="cuda:0")
ready_remote = torch.zeros(NL, dtype=torch.int32, device="cuda:1")
done_remote = torch.zeros(1, dtype=torch.int32, device="cuda:1")
n = 20; sid = list(range(n)); stage = list(range(n))
side = torch.cuda.Stream(device="cuda:0"); sp = int(side.cuda_stream)
def show(tag):
    for d in range(nd): torch.cuda.synchronize(d)
    print(tag, "local", ready_local.tolist(), "remote", ready_remote.tolist(), "done_remote", done_remote.tolist(), flush=True)
# 1. single destination, flags through the plain option, LOCAL flag memory
root.execute_transfer(h_src, sid, h_mc, stage, TransferOptions(multicast=1, epoch=1, layer_done_flags=ready_local.data_ptr())).wait(20)
show("mc + layer_done_flags(local)")
# 2. remote flag memory through the plain option
root.execute_transfer(h_src, sid, h_mc, stage, TransferOptions(multicast=1, epoch=2, layer_done_flags=ready_remote.data_ptr(), done_flag=done_remote.data_ptr())).wait(20)
show("mc + layer_done_flags(remote)")
# 3. per-destination arrays through fan-out, caller stream
root.execute_fanout(h_src, [h_mc], [sid], [stage], True, TransferOptions(multicast=1, epoch=3, cuda_stream=sp, per_dst_layer_done_flags=[ready_remote.data_ptr()]))
side.synchronize()
show("mc fanout per_dst (remote)")
# 4. no multicast: unicast to the remote bound pool with remote flags
h_uc = root.register_layer_separate(cfg, [pools[1].data_ptr() + l * PER_LAYER for l in range(NL)], [PER_LAYER] * NL, BlockDimension.BlockIsSecondDim, StorageKind.Device, 1)
root.execute_transfer(h_src, sid, h_uc, stage, TransferOptions(epoch=4, layer_done_flags=ready_remote.data_ptr())).wait(20)
show("unicast + layer_done_flags(remote)")
print("data ok:", torch.equal(pools[1][:PER_LAYER].view(NO, NB, REGION)[:, :n].cpu(), src[0].view(NO, NB, REGION)[:, :n].cpu()))
os._exit(0)

"""Probe: does this box support NVLink multicast objects (NVLS) from user space?  Bench-only, uses cuda-python."""
import json
import sys

from cuda.bindings import driver as drv


def ck(r):
    err, *rest = r
    if err != drv.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"{err}")
    return rest[0] if len(rest) == 1 else rest


out = {}
try:
    ck(drv.cuInit(0))
    n = ck(drv.cuDeviceGetCount())
    out["devices"] = n
    devs = [ck(drv.cuDeviceGet(i)) for i in range(n)]
    A = drv.CUdevice_attribute
    out["multicast_supported"] = [ck(drv.cuDeviceGetAttribute(A.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d)) for d in devs]
    out["posix_fd"] = [ck(drv.cuDeviceGetAttribute(A.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, d)) for d in devs]
    out["fabric"] = [ck(drv.cuDeviceGetAttribute(A.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, d)) for d in devs]
    ctxs = [ck(drv.cuDevicePrimaryCtxRetain(d)) for d in devs]
    ck(drv.cuCtxSetCurrent(ctxs[0]))
    prop = drv.CUmulticastObjectProp()
    prop.numDevices = n
    prop.size = 1 << 29
    prop.handleTypes = int(drv.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR)
    prop.flags = 0
    out["gran_min"] = ck(drv.cuMulticastGetGranularity(prop, drv.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_MINIMUM))
    out["gran_rec"] = ck(drv.cuMulticastGetGranularity(prop, drv.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED))
    mc = ck(drv.cuMulticastCreate(prop))
    out["create"] = "ok"
    for d in devs:
        ck(drv.cuMulticastAddDevice(mc, d))
    out["add"] = "ok"
    mems = []
    for i, d in enumerate(devs):
        ap = drv.CUmemAllocationProp()
        ap.type = drv.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
        ap.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        ap.location.id = i
        ap.requestedHandleTypes = drv.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        out.setdefault("mem_gran", ck(drv.cuMemGetAllocationGranularity(ap, drv.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_RECOMMENDED)))
        h = ck(drv.cuMemCreate(prop.size, ap, 0))
        mems.append(h)
        ck(drv.cuMulticastBindMem(mc, 0, h, 0, prop.size, 0))
    out["bind"] = "ok"
    va = ck(drv.cuMemAddressReserve(prop.size, out["gran_rec"], 0, 0))
    ck(drv.cuMemMap(va, prop.size, 0, mc, 0))
    ad = drv.CUmemAccessDesc()
    ad.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    ad.location.id = 0
    ad.flags = drv.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
    ck(drv.cuMemSetAccess(va, prop.size, [ad], 1))
    out["map"] = "ok"
    fd = ck(drv.cuMemExportToShareableHandle(mc, drv.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0))
    out["export_fd"] = int(fd)
except Exception as e:  # noqa: BLE001
    out["error"] = repr(e)
print(json.dumps(out))
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/mc_probe.json", "w"))

#!/usr/bin/env python
"""Does this box support NVSwitch multicast objects (the NVLS path: one `multimem.st` lands in every GPU's memory)?

One process, all visible GPUs: queries CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, then tries the whole set-up sequence a
multicast result arena would need -- cuMulticastCreate, cuMulticastAddDevice per GPU, one physical allocation per GPU bound
with cuMulticastBindMem, and a virtual mapping of the multicast handle on every GPU.  Prints one JSON line."""
import json
import sys

from cuda.bindings import driver as cu


def ck(res, what):
    err = res[0]
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"{what}: {cu.cuGetErrorName(err)[1].decode()}")
    return res[1] if len(res) == 2 else res[1:]


def main():
    out = {}
    try:
        ck(cu.cuInit(0), "cuInit")
        ndev = ck(cu.cuDeviceGetCount(), "count")
        out["devices"] = ndev
        devs = [ck(cu.cuDeviceGet(i), "get") for i in range(ndev)]
        out["multicast_supported"] = [int(ck(cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d), "attr")) for d in devs]
        out["fabric_handle_supported"] = [int(ck(cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, d), "attr")) for d in devs]
        out["posix_fd_handle_supported"] = [int(ck(cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, d), "attr")) for d in devs]
        if ndev < 2 or not all(out["multicast_supported"]):
            out["result"] = "not attempted"
            print(json.dumps(out)); return
        ctxs = [ck(cu.cuDevicePrimaryCtxRetain(d), "ctx") for d in devs]
        size = 64 << 20
        prop = cu.CUmulticastObjectProp()
        prop.numDevices = ndev
        prop.size = size
        prop.handleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        prop.flags = 0
        gran = ck(cu.cuMulticastGetGranularity(prop, cu.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED), "mc granularity")
        out["granularity"] = int(gran)
        size = (size + gran - 1) // gran * gran
        prop.size = size
        mc = ck(cu.cuMulticastCreate(prop), "cuMulticastCreate")
        out["create"] = "ok"
        for d in devs:
            ck(cu.cuMulticastAddDevice(mc, d), "cuMulticastAddDevice")
        out["add_devices"] = "ok"
        handles = []
        for i, d in enumerate(devs):
            ck(cu.cuCtxSetCurrent(ctxs[i]), "setctx")
            ap = cu.CUmemAllocationProp()
            ap.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
            ap.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
            ap.location.id = i
            ap.requestedHandleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
            h = ck(cu.cuMemCreate(size, ap, 0), "cuMemCreate")
            handles.append(h)
            ck(cu.cuMulticastBindMem(mc, 0, h, 0, size, 0), "cuMulticastBindMem")
        out["bind"] = "ok"
        # map the multicast object into the address space and give every device access
        va = ck(cu.cuMemAddressReserve(size, gran, 0, 0), "reserve")
        ck(cu.cuMemMap(va, size, 0, mc, 0), "cuMemMap(mc)")
        descs = []
        for i in range(ndev):
            ad = cu.CUmemAccessDesc()
            ad.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
            ad.location.id = i
            ad.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
            descs.append(ad)
        ck(cu.cuMemSetAccess(va, size, descs, ndev), "cuMemSetAccess(mc)")
        out["map"] = "ok"
        out["result"] = "multicast arena set-up works"
        # the same with FABRIC handles (64-byte opaque handles that travel through any host transport, like CUDA IPC handles)
        try:
            ck(cu.cuCtxSetCurrent(ctxs[0]), "setctx")
            ap = cu.CUmemAllocationProp()
            ap.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
            ap.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
            ap.location.id = 0
            ap.requestedHandleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_FABRIC
            g2 = ck(cu.cuMemGetAllocationGranularity(ap, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM), "gran")
            out["alloc_granularity_min"] = int(g2)
            hf = ck(cu.cuMemCreate(size, ap, 0), "cuMemCreate(fabric)")
            fh = ck(cu.cuMemExportToShareableHandle(hf, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_FABRIC, 0), "export(fabric)")
            out["fabric_alloc_export"] = "ok"
            prop2 = cu.CUmulticastObjectProp()
            prop2.numDevices = ndev; prop2.size = size; prop2.flags = 0
            prop2.handleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_FABRIC
            mc2 = ck(cu.cuMulticastCreate(prop2), "cuMulticastCreate(fabric)")
            fh2 = ck(cu.cuMemExportToShareableHandle(mc2, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_FABRIC, 0), "export mc(fabric)")
            out["fabric_multicast_export"] = "ok"
        except Exception as e:       # noqa: BLE001
            out["fabric"] = f"failed: {e}"
        try:
            gmin = ck(cu.cuMulticastGetGranularity(prop, cu.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_MINIMUM), "mc granularity min")
            out["granularity_min"] = int(gmin)
            fd = ck(cu.cuMemExportToShareableHandle(mc, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "export mc(fd)")
            out["posix_fd_multicast_export"] = "ok"
            import os
            pidfd = os.pidfd_open(os.getpid())
            import ctypes
            libc = ctypes.CDLL(None, use_errno=True)
            r = libc.syscall(438, pidfd, int(fd), 0)
            out["pidfd_getfd"] = "ok" if r >= 0 else f"errno {ctypes.get_errno()}"
        except Exception as e:       # noqa: BLE001
            out["fd_export"] = f"failed: {e}"
    except Exception as e:       # noqa: BLE001 -- a probe: report whatever stopped it
        out["result"] = f"failed: {e}"
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())

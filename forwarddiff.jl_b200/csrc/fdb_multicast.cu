// fdb_multicast.cu -- NVSwitch multicast (NVLS) form of the result exchange.
//
// fdb_exchange.cu repeats every extraction store once per peer: a rank's NVLink egress is (W-1) x its result slice and
// the storing thread issues W stores per entry.  NVSwitch can replicate a write itself: the ranks bind their arenas into
// one *multicast object*; a store to the multicast mapping at offset `off` is delivered by the switch to offset `off` of
// every bound arena.  A rank then sends each entry ONCE (egress = its slice) and the kernels issue one store per entry.
//
// The arenas are virtual-memory-management allocations (cuMemCreate) because only those can be bound to a multicast
// object; they are shared between the processes as POSIX file descriptors.  A descriptor number means nothing in another
// process, so the handle that travels through the caller's host transport is {pid, fd} and the importer duplicates the
// descriptor out of the exporting process with pidfd_getfd(2) (same user, same node).  Fabric handles would be plain
// 64-byte blobs but need an IMEX channel (cuMemCreate: CUDA_ERROR_NOT_PERMITTED on the boxes this was developed on).
//
// Sequence (every step is collective over the ranks; "host barrier" = any barrier of the caller's transport):
//   fdb_exchange_create_multicast      -> arena
//   fdb_exchange_share_handle          -> {pid, fd} of the arena, all-gathered by the caller
//   fdb_exchange_multicast_create      -> rank 0 only: the multicast object, {pid, fd} broadcast by the caller
//   fdb_exchange_connect_multicast     -> maps the peers' arenas (unicast: barrier flags, fdb_exchange_push), adds this
//                                         rank's device to the multicast object
//   host barrier                          (every device must have been added before memory is bound)
//   fdb_exchange_multicast_bind        -> binds the local arena, maps the multicast object
//   host barrier                          (nobody stores through the multicast mapping before every arena is bound)
// The driver API is reached through cudaGetDriverEntryPoint (libfdb200.so does not link libcuda: it must load on a
// machine without a driver for the symbol checks of the CPU test suite).
#include <cuda.h>
#include <errno.h>
#include <string.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "fdb_internal.h"

namespace fdb {

static constexpr size_t MC_FLAG_PAGE = 256;       // = FLAG_PAGE of fdb_exchange.cu: the barrier flags sit behind the data

struct Drv {
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*MemExport)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImport)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    CUresult (*MemGetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*McCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
    CUresult (*McAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*McBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
    CUresult (*McUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
    CUresult (*McGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*GetErrorName)(CUresult, const char**) = nullptr;
    bool ok = false;
};

template <typename F>
static bool entry(const char* name, F& fn) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) { cudaGetLastError(); return false; }
    fn = reinterpret_cast<F>(p);
    return true;
}

static const Drv& drv() {
    static Drv d;
    static bool tried = false;
    if (!tried) {
        tried = true;
        d.ok = entry("cuMemCreate", d.MemCreate) && entry("cuMemRelease", d.MemRelease) && entry("cuMemAddressReserve", d.MemAddressReserve) &&
               entry("cuMemAddressFree", d.MemAddressFree) && entry("cuMemMap", d.MemMap) && entry("cuMemUnmap", d.MemUnmap) &&
               entry("cuMemSetAccess", d.MemSetAccess) && entry("cuMemExportToShareableHandle", d.MemExport) &&
               entry("cuMemImportFromShareableHandle", d.MemImport) && entry("cuMemGetAllocationGranularity", d.MemGetGranularity) &&
               entry("cuMulticastCreate", d.McCreate) && entry("cuMulticastAddDevice", d.McAddDevice) && entry("cuMulticastBindMem", d.McBindMem) &&
               entry("cuMulticastUnbind", d.McUnbind) && entry("cuMulticastGetGranularity", d.McGetGranularity) && entry("cuDeviceGet", d.DeviceGet) &&
               entry("cuDeviceGetAttribute", d.DeviceGetAttribute) && entry("cuGetErrorName", d.GetErrorName);
    }
    return d;
}

static const char* cu_name(CUresult r) {
    const char* s = nullptr;
    if (drv().GetErrorName && drv().GetErrorName(r, &s) == CUDA_SUCCESS && s) return s;
    return "CUDA driver error";
}
#define FDB_CU(ctx, call)                                                                                     \
    do {                                                                                                      \
        CUresult cr__ = (call);                                                                               \
        if (cr__ != CUDA_SUCCESS) return fail((ctx), FDB_ERR_CUDA, "%s: %s", #call, cu_name(cr__));           \
    } while (0)

static CUmemAllocationProp arena_prop(int device) {
    CUmemAllocationProp p;
    memset(&p, 0, sizeof p);
    p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    p.location.id = device;
    p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return p;
}
static CUmulticastObjectProp mc_prop(int world, size_t bytes) {
    CUmulticastObjectProp p;
    memset(&p, 0, sizeof p);
    p.numDevices = (unsigned)world;
    p.size = bytes;
    p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return p;
}

// map `h` (an allocation or a multicast object) read-write for this context's device
static int map_handle(fdb_ctx* ctx, CUmemGenericAllocationHandle h, size_t bytes, size_t gran, char** out) {
    const Drv& d = drv();
    CUdeviceptr va = 0;
    FDB_CU(ctx, d.MemAddressReserve(&va, bytes, gran, 0, 0));
    CUresult cr = d.MemMap(va, bytes, 0, h, 0);
    if (cr != CUDA_SUCCESS) { d.MemAddressFree(va, bytes); return fail(ctx, FDB_ERR_CUDA, "cuMemMap: %s", cu_name(cr)); }
    CUmemAccessDesc ad;
    memset(&ad, 0, sizeof ad);
    ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ad.location.id = ctx->device;
    ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    cr = d.MemSetAccess(va, bytes, &ad, 1);
    if (cr != CUDA_SUCCESS) { d.MemUnmap(va, bytes); d.MemAddressFree(va, bytes); return fail(ctx, FDB_ERR_CUDA, "cuMemSetAccess: %s", cu_name(cr)); }
    *out = reinterpret_cast<char*>(va);
    return FDB_OK;
}
static void unmap(char* p, size_t bytes) {
    if (!p) return;
    drv().MemUnmap(reinterpret_cast<CUdeviceptr>(p), bytes);
    drv().MemAddressFree(reinterpret_cast<CUdeviceptr>(p), bytes);
}

struct ShareHandle { int pid; int fd; int pad[2]; };
static_assert(sizeof(ShareHandle) == FDB_SHARE_HANDLE_BYTES, "share handle size");

// duplicate descriptor `fd` of process `pid` into this process
static int fetch_fd(fdb_ctx* ctx, const ShareHandle& h, int* out) {
    if (h.pid == (int)getpid()) { *out = dup(h.fd); if (*out < 0) return fail(ctx, FDB_ERR_CUDA, "dup(%d) failed", h.fd); return FDB_OK; }
#if defined(SYS_pidfd_open) && defined(SYS_pidfd_getfd)
    const int pidfd = (int)syscall(SYS_pidfd_open, h.pid, 0);
    if (pidfd < 0) return fail(ctx, FDB_ERR_UNSUPPORTED, "pidfd_open(%d) failed (errno %d): cannot import a peer's memory handle", h.pid, errno);
    const int fd = (int)syscall(SYS_pidfd_getfd, pidfd, h.fd, 0);
    const int err = errno;
    close(pidfd);
    if (fd < 0) return fail(ctx, FDB_ERR_UNSUPPORTED, "pidfd_getfd(pid %d, fd %d) failed (errno %d): ptrace access to the peer process is needed", h.pid, h.fd, err);
    *out = fd;
    return FDB_OK;
#else
    return fail(ctx, FDB_ERR_UNSUPPORTED, "pidfd_getfd is not available on this platform");
#endif
}

int multicast_destroy(fdb_exchange* x) {        // called by fdb_exchange_destroy for vmm arenas
    const Drv& d = drv();
    if (x->mc_base) { unmap(x->mc_base, x->map_bytes); x->mc_base = nullptr; }
    if (x->mc_handle) {
        CUdevice dev;
        if (x->mem_handle && d.DeviceGet(&dev, x->ctx->device) == CUDA_SUCCESS) d.McUnbind(x->mc_handle, dev, 0, x->map_bytes);
        d.MemRelease(x->mc_handle); x->mc_handle = 0;
    }
    for (int r = 0; r < x->world; r++)
        if (r != x->rank && x->peer[r]) { unmap(x->peer[r], x->map_bytes); x->peer[r] = nullptr; }
    if (x->base) { unmap(x->base, x->map_bytes); x->base = nullptr; }
    if (x->mem_handle) { d.MemRelease(x->mem_handle); x->mem_handle = 0; }
    for (int i = 0; i < 2; i++) if (x->exported_fd[i] >= 0) { close(x->exported_fd[i]); x->exported_fd[i] = -1; }
    return FDB_OK;
}

}  // namespace fdb

using namespace fdb;

extern "C" int fdb_multicast_supported(fdb_ctx* ctx, int* supported) {
    FDB_ENTER(ctx);
    if (!ctx || !supported) return FDB_ERR_BADARG;
    *supported = 0;
    const Drv& d = drv();
    if (!d.ok) return FDB_OK;
    CUdevice dev; int v = 0;
    if (d.DeviceGet(&dev, ctx->device) != CUDA_SUCCESS) return FDB_OK;
    if (d.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS && v) {
        int fdok = 0;
        d.DeviceGetAttribute(&fdok, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
        *supported = fdok ? 1 : 0;
    }
    return FDB_OK;
}

extern "C" int fdb_exchange_create_multicast(fdb_ctx* ctx, int64_t n_doubles, int rank, int world, fdb_exchange** out) {
    FDB_ENTER(ctx);
    if (!ctx || !out || n_doubles < 0) return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_create_multicast: bad argument");
    if (world < 1 || world > FDB_MAX_RANKS || rank < 0 || rank >= world)
        return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_create_multicast: rank %d of %d (at most %d ranks, one NVSwitch domain)", rank, world, FDB_MAX_RANKS);
    int sup = 0;
    fdb_multicast_supported(ctx, &sup);
    if (!sup) return fail(ctx, FDB_ERR_UNSUPPORTED, "device %d has no multicast support (CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED) or the driver entry points are missing", ctx->device);
    const Drv& d = drv();
    cudaSetDevice(ctx->device);
    cudaFree(0);                                          // the primary context must exist before driver calls
    const size_t data_bytes = (size_t)round_up(n_doubles > 0 ? n_doubles : 1, 32) * sizeof(double);
    CUmemAllocationProp ap = arena_prop(ctx->device);
    CUmulticastObjectProp mp = mc_prop(world, data_bytes + MC_FLAG_PAGE);
    size_t g_alloc = 0, g_mc = 0;
    FDB_CU(ctx, d.MemGetGranularity(&g_alloc, &ap, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
    FDB_CU(ctx, d.McGetGranularity(&g_mc, &mp, CU_MULTICAST_GRANULARITY_MINIMUM));
    const size_t gran = g_alloc > g_mc ? g_alloc : g_mc;
    const size_t bytes = (data_bytes + MC_FLAG_PAGE + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h = 0;
    CUresult cr = d.MemCreate(&h, bytes, &ap, 0);
    if (cr != CUDA_SUCCESS) return fail(ctx, cr == CUDA_ERROR_OUT_OF_MEMORY ? FDB_ERR_OOM : FDB_ERR_CUDA, "cuMemCreate(%zu bytes): %s", bytes, cu_name(cr));
    fdb_exchange* x = new fdb_exchange();
    x->ctx = ctx; x->rank = rank; x->world = world; x->data_bytes = data_bytes;
    x->vmm = true; x->mem_handle = h; x->map_bytes = bytes;
    int rc = map_handle(ctx, h, bytes, gran, &x->base);
    if (rc) { d.MemRelease(h); delete x; return rc; }
    cudaError_t e = cudaMemset(x->base, 0, bytes);
    if (e == cudaSuccess && !ctx->xch_timeout_flag) {
        e = cudaHostAlloc(reinterpret_cast<void**>(&ctx->xch_timeout_flag), 64, cudaHostAllocMapped | cudaHostAllocPortable);
        if (e == cudaSuccess) *ctx->xch_timeout_flag = 0;
    }
    if (e != cudaSuccess) { multicast_destroy(x); delete x; return fail(ctx, FDB_ERR_CUDA, "arena set-up: %s", cudaGetErrorString(e)); }
    x->peer[rank] = x->base;
    x->connected = (world == 1);
    *out = x;
    return FDB_OK;
}

extern "C" int fdb_exchange_share_handle(fdb_exchange* x, unsigned char handle[FDB_SHARE_HANDLE_BYTES]) {
    FDB_ENTER(x ? x->ctx : nullptr);
    if (!x || !handle) return FDB_ERR_BADARG;
    if (!x->vmm) return fail(x->ctx, FDB_ERR_BADARG, "fdb_exchange_share_handle: the arena was not created by fdb_exchange_create_multicast (use fdb_exchange_handle)");
    if (x->exported_fd[0] < 0) {
        int fd = -1;
        FDB_CU(x->ctx, drv().MemExport(&fd, x->mem_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
        x->exported_fd[0] = fd;
    }
    ShareHandle h = {(int)getpid(), x->exported_fd[0], {0, 0}};
    memcpy(handle, &h, sizeof h);
    return FDB_OK;
}

extern "C" int fdb_exchange_multicast_create(fdb_exchange* x, unsigned char handle[FDB_SHARE_HANDLE_BYTES]) {
    FDB_ENTER(x ? x->ctx : nullptr);
    if (!x || !handle) return FDB_ERR_BADARG;
    fdb_ctx* ctx = x->ctx;
    if (!x->vmm) return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_multicast_create: the arena was not created by fdb_exchange_create_multicast");
    if (x->rank != 0) return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_multicast_create: rank 0 creates the multicast object, the others import it");
    if (x->world < 2) return fail(ctx, FDB_ERR_UNSUPPORTED, "fdb_exchange_multicast_create: a multicast object spans at least two devices");
    const Drv& d = drv();
    if (!x->mc_handle) {
        CUmulticastObjectProp mp = mc_prop(x->world, x->map_bytes);
        CUmemGenericAllocationHandle mc = 0;
        FDB_CU(ctx, d.McCreate(&mc, &mp));
        x->mc_handle = mc;
    }
    if (x->exported_fd[1] < 0) {
        int fd = -1;
        FDB_CU(ctx, d.MemExport(&fd, x->mc_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
        x->exported_fd[1] = fd;
    }
    ShareHandle h = {(int)getpid(), x->exported_fd[1], {0, 0}};
    memcpy(handle, &h, sizeof h);
    return FDB_OK;
}

extern "C" int fdb_exchange_connect_multicast(fdb_exchange* x, const unsigned char* arena_handles, const unsigned char* mc_handle) {
    FDB_ENTER(x ? x->ctx : nullptr);
    if (!x || !arena_handles || !mc_handle) return FDB_ERR_BADARG;
    fdb_ctx* ctx = x->ctx;
    if (!x->vmm) return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_connect_multicast: the arena was not created by fdb_exchange_create_multicast");
    if (x->connected && x->world > 1) return FDB_OK;
    const Drv& d = drv();
    cudaSetDevice(ctx->device);
    CUmemAllocationProp ap = arena_prop(ctx->device);
    size_t gran = 0;
    FDB_CU(ctx, d.MemGetGranularity(&gran, &ap, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
    for (int r = 0; r < x->world; r++) {
        if (r == x->rank) continue;
        ShareHandle sh;
        memcpy(&sh, arena_handles + (size_t)r * FDB_SHARE_HANDLE_BYTES, sizeof sh);
        int fd = -1;
        int rc = fetch_fd(ctx, sh, &fd);
        if (rc) return rc;
        CUmemGenericAllocationHandle h = 0;
        CUresult cr = d.MemImport(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        close(fd);
        if (cr != CUDA_SUCCESS) return fail(ctx, FDB_ERR_CUDA, "cuMemImportFromShareableHandle(rank %d's arena): %s", r, cu_name(cr));
        rc = map_handle(ctx, h, x->map_bytes, gran, &x->peer[r]);
        d.MemRelease(h);                                  // the mapping keeps the allocation alive
        if (rc) return rc;
    }
    if (!x->mc_handle) {
        ShareHandle sh;
        memcpy(&sh, mc_handle, sizeof sh);
        int fd = -1;
        int rc = fetch_fd(ctx, sh, &fd);
        if (rc) return rc;
        CUmemGenericAllocationHandle mc = 0;
        CUresult cr = d.MemImport(&mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        close(fd);
        if (cr != CUDA_SUCCESS) return fail(ctx, FDB_ERR_CUDA, "cuMemImportFromShareableHandle(multicast object): %s", cu_name(cr));
        x->mc_handle = mc;
    }
    CUdevice dev;
    FDB_CU(ctx, d.DeviceGet(&dev, ctx->device));
    FDB_CU(ctx, d.McAddDevice(x->mc_handle, dev));
    x->connected = true;
    return FDB_OK;
}

extern "C" int fdb_exchange_multicast_bind(fdb_exchange* x) {
    FDB_ENTER(x ? x->ctx : nullptr);
    if (!x) return FDB_ERR_BADARG;
    fdb_ctx* ctx = x->ctx;
    if (!x->vmm || !x->mc_handle || !x->connected) return fail(ctx, FDB_ERR_BADARG, "fdb_exchange_multicast_bind: call fdb_exchange_connect_multicast first");
    if (x->mc_base) return FDB_OK;
    const Drv& d = drv();
    cudaSetDevice(ctx->device);
    FDB_CU(ctx, d.McBindMem(x->mc_handle, 0, x->mem_handle, 0, x->map_bytes, 0));
    CUmulticastObjectProp mp = mc_prop(x->world, x->map_bytes);
    size_t gran = 0;
    FDB_CU(ctx, d.McGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_MINIMUM));
    return map_handle(ctx, x->mc_handle, x->map_bytes, gran, &x->mc_base);
}

extern "C" int fdb_exchange_set_multicast(fdb_exchange* x, int on) {
    if (!x) return FDB_ERR_BADARG;
    if (on && !x->mc_base) return fail(x->ctx, FDB_ERR_BADARG, "fdb_exchange_set_multicast: no multicast mapping (fdb_exchange_multicast_bind)");
    x->mc_on = on != 0;
    return FDB_OK;
}

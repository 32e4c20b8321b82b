// psx_nvls.cuh -- NVSwitch multicast (NVLS) objects.  Included at the end of psx.cu.
//
// Two forms:
//   psx_mc_*   all member GPUs in ONE process (micro-benchmarks, tests/test_gpu_nvls.py)
//   psx_mcx_*  one MEMBER per (process, GPU): the multicast object's POSIX file
//              descriptor travels from the creating process to the others over an
//              AF_UNIX socket (SCM_RIGHTS; the Python host does that), every
//              member adds its device, then binds its own VMM allocation.  This
//              is what the NVLS form of psx_round uses (psx_round_bind_mc): the
//              product runs one process per GPU.
//
//   multicast buffer = one cuMemCreate allocation per GPU, all bound at offset 0
//   of ONE multicast object (cuMulticastCreate / AddDevice / BindMem), mapped
//   twice: per-GPU unicast addresses (ordinary loads/stores, peer-accessible) and
//   one multicast address on which
//       multimem.st         stores to EVERY GPU's copy        (PS -> workers)
//       multimem.ld_reduce  returns the SUM over all copies   (workers -> PS)
//   are executed by the switch.
//
// Every driver entry point is resolved through cudaGetDriverEntryPoint, so the
// library still has no link-time dependency on libcuda.
#pragma once

namespace {

struct DriverVmm {
    CUresult (*DeviceGet)(CUdevice *, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int *, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*MulticastGetGranularity)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags) = nullptr;
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *) = nullptr;
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
    CUresult (*MemExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
    bool ok = false;
};
DriverVmm g_drv;
std::once_flag g_drv_once;

template <typename F> bool drv_resolve(const char *name, F *out)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || fn == nullptr)
        return false;
    *out = (F)fn;
    return true;
}

int load_driver_vmm()
{
    std::call_once(g_drv_once, [] {
        bool ok = true;
        ok &= drv_resolve("cuDeviceGet", &g_drv.DeviceGet);
        ok &= drv_resolve("cuDeviceGetAttribute", &g_drv.DeviceGetAttribute);
        ok &= drv_resolve("cuMulticastGetGranularity", &g_drv.MulticastGetGranularity);
        ok &= drv_resolve("cuMulticastCreate", &g_drv.MulticastCreate);
        ok &= drv_resolve("cuMulticastAddDevice", &g_drv.MulticastAddDevice);
        ok &= drv_resolve("cuMulticastBindMem", &g_drv.MulticastBindMem);
        ok &= drv_resolve("cuMulticastUnbind", &g_drv.MulticastUnbind);
        ok &= drv_resolve("cuMemCreate", &g_drv.MemCreate);
        ok &= drv_resolve("cuMemRelease", &g_drv.MemRelease);
        ok &= drv_resolve("cuMemAddressReserve", &g_drv.MemAddressReserve);
        ok &= drv_resolve("cuMemAddressFree", &g_drv.MemAddressFree);
        ok &= drv_resolve("cuMemMap", &g_drv.MemMap);
        ok &= drv_resolve("cuMemUnmap", &g_drv.MemUnmap);
        ok &= drv_resolve("cuMemSetAccess", &g_drv.MemSetAccess);
        ok &= drv_resolve("cuGetErrorString", &g_drv.GetErrorString);
        ok &= drv_resolve("cuMemExportToShareableHandle", &g_drv.MemExportToShareableHandle);
        ok &= drv_resolve("cuMemImportFromShareableHandle", &g_drv.MemImportFromShareableHandle);
        g_drv.ok = ok;
    });
    if (!g_drv.ok) return fail(PSX_ECUDA, "the driver does not expose the VMM / multicast entry points");
    return PSX_OK;
}

int drv_fail(const char *what, CUresult r)
{
    const char *msg = "?";
    if (g_drv.GetErrorString) g_drv.GetErrorString(r, &msg);
    return fail(PSX_ECUDA, "%s failed: CUresult %d (%s)", what, (int)r, msg ? msg : "?");
}
#define DRV_TRY(call, what)                         \
    do {                                            \
        CUresult r_ = (call);                       \
        if (r_ != CUDA_SUCCESS) return drv_fail(what, r_); \
    } while (0)

constexpr int kMcMaxDevices = 16;
struct McBuffer {
    int n = 0;
    int device[kMcMaxDevices] = {};
    size_t size = 0;
    CUmemGenericAllocationHandle mc = 0;
    CUmemGenericAllocationHandle mem[kMcMaxDevices] = {};
    CUdeviceptr uc[kMcMaxDevices] = {};
    CUdeviceptr mcva = 0;
    bool bound[kMcMaxDevices] = {};
};
std::unordered_map<uint64_t, McBuffer *> g_mcs;

void mc_release(McBuffer *b)
{
    for (int i = 0; i < b->n; ++i) {
        if (b->uc[i]) {
            g_drv.MemUnmap(b->uc[i], b->size);
            g_drv.MemAddressFree(b->uc[i], b->size);
        }
    }
    if (b->mcva) {
        g_drv.MemUnmap(b->mcva, b->size);
        g_drv.MemAddressFree(b->mcva, b->size);
    }
    for (int i = 0; i < b->n; ++i) {
        if (b->bound[i]) {
            CUdevice d;
            if (g_drv.DeviceGet(&d, b->device[i]) == CUDA_SUCCESS) g_drv.MulticastUnbind(b->mc, d, 0, b->size);
        }
        if (b->mem[i]) g_drv.MemRelease(b->mem[i]);
    }
    if (b->mc) g_drv.MemRelease(b->mc);
    delete b;
}

// dst (every GPU's copy, through the multicast address) = src
__global__ void __launch_bounds__(256)
k_mc_broadcast(float *mc_dst, const float4 *__restrict__ src, size_t n4)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = psx::ld_stream(src + i);
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_dst + 4 * i),
                     "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
    }
}

// dst = sum over every GPU's copy (reduced in the switch)
__global__ void __launch_bounds__(256)
k_mc_reduce(float4 *__restrict__ dst, const float *mc_src, size_t n4)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (size_t)gridDim.x * blockDim.x) {
        float4 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "l"(mc_src + 4 * i)
                     : "memory");
        psx::st_stream(dst + i, v);
    }
}

}  // namespace

extern "C" {

int psx_nvls_supported(int device, int *out)
{
    if (!out) return fail(PSX_EINVAL, "null out");
    *out = 0;
    int rc = load_driver_vmm();
    if (rc) return rc;
    CU_TRY(cudaFree(0));
    CUdevice d;
    DRV_TRY(g_drv.DeviceGet(&d, device), "cuDeviceGet");
    int v = 0;
    DRV_TRY(g_drv.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d),
            "cuDeviceGetAttribute(MULTICAST_SUPPORTED)");
    *out = v;
    return PSX_OK;
}

int psx_mc_create(const int *devices, int n, uint64_t nbytes, uint64_t *out_id)
{
    if (!devices || !out_id || n < 2 || n > kMcMaxDevices || nbytes == 0)
        return fail(PSX_EINVAL, "psx_mc_create: 2..%d devices and a non-empty size", kMcMaxDevices);
    int rc = load_driver_vmm();
    if (rc) return rc;
    for (int i = 0; i < n; ++i) {
        PSX_DEVICE(devices[i]);
        CU_TRY(cudaFree(0));                       // primary context of every member
        int sup = 0;
        rc = psx_nvls_supported(devices[i], &sup);
        if (rc) return rc;
        if (!sup) return fail(PSX_ECUDA, "device %d does not support NVSwitch multicast", devices[i]);
    }
    CUmulticastObjectProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.numDevices = (unsigned)n;
    prop.size = nbytes;
    prop.handleTypes = 0;
    size_t gran = 0;
    DRV_TRY(g_drv.MulticastGetGranularity(&gran, &prop, CU_MULTICAST_GRANULARITY_RECOMMENDED),
            "cuMulticastGetGranularity");
    const size_t size = (nbytes + gran - 1) / gran * gran;
    prop.size = size;

    McBuffer *b = new McBuffer();
    b->n = n;
    b->size = size;
    for (int i = 0; i < n; ++i) b->device[i] = devices[i];
#define MC_TRY(call, what)                                  \
    do {                                                    \
        CUresult r_ = (call);                               \
        if (r_ != CUDA_SUCCESS) {                           \
            int rc_ = drv_fail(what, r_);                   \
            mc_release(b);                                  \
            return rc_;                                     \
        }                                                   \
    } while (0)
    MC_TRY(g_drv.MulticastCreate(&b->mc, &prop), "cuMulticastCreate");
    CUdevice cu[kMcMaxDevices];
    for (int i = 0; i < n; ++i) {
        MC_TRY(g_drv.DeviceGet(&cu[i], devices[i]), "cuDeviceGet");
        MC_TRY(g_drv.MulticastAddDevice(b->mc, cu[i]), "cuMulticastAddDevice");
    }
    CUmemAccessDesc access[kMcMaxDevices];
    for (int i = 0; i < n; ++i) {
        access[i].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        access[i].location.id = devices[i];
        access[i].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    }
    for (int i = 0; i < n; ++i) {
        CUmemAllocationProp ap;
        memset(&ap, 0, sizeof(ap));
        ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        ap.location.id = devices[i];
        ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_NONE;
        MC_TRY(g_drv.MemCreate(&b->mem[i], size, &ap, 0), "cuMemCreate");
        MC_TRY(g_drv.MulticastBindMem(b->mc, 0, b->mem[i], 0, size, 0), "cuMulticastBindMem");
        b->bound[i] = true;
        MC_TRY(g_drv.MemAddressReserve(&b->uc[i], size, gran, 0, 0), "cuMemAddressReserve(unicast)");
        MC_TRY(g_drv.MemMap(b->uc[i], size, 0, b->mem[i], 0), "cuMemMap(unicast)");
        MC_TRY(g_drv.MemSetAccess(b->uc[i], size, access, (size_t)n), "cuMemSetAccess(unicast)");
    }
    MC_TRY(g_drv.MemAddressReserve(&b->mcva, size, gran, 0, 0), "cuMemAddressReserve(multicast)");
    MC_TRY(g_drv.MemMap(b->mcva, size, 0, b->mc, 0), "cuMemMap(multicast)");
    MC_TRY(g_drv.MemSetAccess(b->mcva, size, access, (size_t)n), "cuMemSetAccess(multicast)");
#undef MC_TRY
    for (int i = 0; i < n; ++i) {
        PSX_DEVICE(devices[i]);
        CU_TRY(cudaMemset((void *)b->uc[i], 0, size));
        CU_TRY(cudaDeviceSynchronize());
    }
    uint64_t id = g_next_id.fetch_add(1);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_mcs[id] = b;
    }
    *out_id = id;
    return PSX_OK;
}

int psx_mc_destroy(uint64_t id)
{
    McBuffer *b = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_mcs.find(id);
        if (it == g_mcs.end()) return fail(PSX_EINVAL, "unknown multicast buffer id");
        b = it->second;
        g_mcs.erase(it);
    }
    for (int i = 0; i < b->n; ++i) {
        cudaSetDevice(b->device[i]);
        cudaDeviceSynchronize();
    }
    mc_release(b);
    return PSX_OK;
}

/* member: index into the device list given at creation */
int psx_mc_ptrs(uint64_t id, int member, void **out_unicast, void **out_multicast, uint64_t *out_size)
{
    McBuffer *b = find(g_mcs, id);
    if (!b || member < 0 || member >= b->n) return fail(PSX_EINVAL, "unknown multicast buffer / member");
    if (out_unicast) *out_unicast = (void *)b->uc[member];
    if (out_multicast) *out_multicast = (void *)b->mcva;
    if (out_size) *out_size = b->size;
    return PSX_OK;
}

int psx_mc_broadcast(uint64_t id, int member, const void *src_dev, uint64_t off_bytes, uint64_t nbytes,
                     void *stream)
{
    McBuffer *b = find(g_mcs, id);
    if (!b || member < 0 || member >= b->n) return fail(PSX_EINVAL, "unknown multicast buffer / member");
    if (off_bytes % 16 || nbytes % 16 || off_bytes + nbytes > b->size || ((uintptr_t)src_dev % 16))
        return fail(PSX_EINVAL, "multicast ranges and sources are 16-byte granular");
    PSX_DEVICE(b->device[member]);
    const size_t n4 = nbytes / 16;
    const int grid = grid_for(n4 ? n4 : 1, 256, sm_count_of(b->device[member]), 8);
    k_mc_broadcast<<<grid, 256, 0, (cudaStream_t)stream>>>((float *)(b->mcva + off_bytes),
                                                           (const float4 *)src_dev, n4);
    LAUNCH_CHECK();
    return PSX_OK;
}

int psx_mc_reduce(uint64_t id, int member, void *dst_dev, uint64_t off_bytes, uint64_t nbytes, void *stream)
{
    McBuffer *b = find(g_mcs, id);
    if (!b || member < 0 || member >= b->n) return fail(PSX_EINVAL, "unknown multicast buffer / member");
    if (off_bytes % 16 || nbytes % 16 || off_bytes + nbytes > b->size || ((uintptr_t)dst_dev % 16))
        return fail(PSX_EINVAL, "multicast ranges and destinations are 16-byte granular");
    PSX_DEVICE(b->device[member]);
    const size_t n4 = nbytes / 16;
    const int grid = grid_for(n4 ? n4 : 1, 256, sm_count_of(b->device[member]), 8);
    k_mc_reduce<<<grid, 256, 0, (cudaStream_t)stream>>>((float4 *)dst_dev, (const float *)(b->mcva + off_bytes), n4);
    LAUNCH_CHECK();
    return PSX_OK;
}

}  // extern "C"

// ------------------------------------------------- multi-process members ----
namespace {

struct McMember {
    int device = 0;
    int n_devices = 0;
    size_t size = 0, gran = 0;
    CUmemGenericAllocationHandle mc = 0, mem = 0;
    CUdeviceptr uc = 0, mcva = 0;
    bool added = false, bound = false;
    int fd = -1;                  // creator only: the exported descriptor (owned here)
};
std::unordered_map<uint64_t, McMember *> g_mcx;

int mcx_prop(int n_devices, uint64_t nbytes, CUmulticastObjectProp *prop, size_t *gran)
{
    memset(prop, 0, sizeof(*prop));
    prop->numDevices = (unsigned)n_devices;
    prop->size = nbytes;
    prop->handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    DRV_TRY(g_drv.MulticastGetGranularity(gran, prop, CU_MULTICAST_GRANULARITY_RECOMMENDED),
            "cuMulticastGetGranularity");
    prop->size = (nbytes + *gran - 1) / *gran * *gran;
    return PSX_OK;
}

void mcx_release(McMember *m)
{
    if (m->uc) {
        g_drv.MemUnmap(m->uc, m->size);
        g_drv.MemAddressFree(m->uc, m->size);
    }
    if (m->mcva) {
        g_drv.MemUnmap(m->mcva, m->size);
        g_drv.MemAddressFree(m->mcva, m->size);
    }
    if (m->bound) {
        CUdevice d;
        if (g_drv.DeviceGet(&d, m->device) == CUDA_SUCCESS) g_drv.MulticastUnbind(m->mc, d, 0, m->size);
    }
    if (m->mem) g_drv.MemRelease(m->mem);
    if (m->mc) g_drv.MemRelease(m->mc);
    if (m->fd >= 0) close(m->fd);
    delete m;
}

int mcx_common(int device, int n_devices, uint64_t nbytes, McMember **out, CUmulticastObjectProp *prop)
{
    if (n_devices < 1 || n_devices > kMcMaxDevices || nbytes == 0)
        return fail(PSX_EINVAL, "multicast member: 1..%d devices and a non-empty size", kMcMaxDevices);
    int rc = load_driver_vmm();
    if (rc) return rc;
    PSX_DEVICE(device);
    CU_TRY(cudaFree(0));
    int sup = 0;
    rc = psx_nvls_supported(device, &sup);
    if (rc) return rc;
    if (!sup) return fail(PSX_ECUDA, "device %d does not support NVSwitch multicast", device);
    size_t gran = 0;
    rc = mcx_prop(n_devices, nbytes, prop, &gran);
    if (rc) return rc;
    McMember *m = new McMember();
    m->device = device;
    m->n_devices = n_devices;
    m->size = prop->size;
    m->gran = gran;
    *out = m;
    return PSX_OK;
}

uint64_t mcx_register(McMember *m)
{
    uint64_t id = g_next_id.fetch_add(1);
    std::lock_guard<std::mutex> lk(g_mu);
    g_mcx[id] = m;
    return id;
}

}  // namespace

extern "C" {

int psx_mcx_create(int device, int n_devices, uint64_t nbytes, int *out_fd, uint64_t *out_id)
{
    if (!out_fd || !out_id) return fail(PSX_EINVAL, "null out pointer");
    McMember *m = nullptr;
    CUmulticastObjectProp prop;
    int rc = mcx_common(device, n_devices, nbytes, &m, &prop);
    if (rc) return rc;
    CUresult r = g_drv.MulticastCreate(&m->mc, &prop);
    if (r != CUDA_SUCCESS) {
        mcx_release(m);
        return drv_fail("cuMulticastCreate", r);
    }
    int fd = -1;
    r = g_drv.MemExportToShareableHandle(&fd, m->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) {
        mcx_release(m);
        return drv_fail("cuMemExportToShareableHandle(multicast)", r);
    }
    m->fd = fd;
    *out_fd = fd;
    *out_id = mcx_register(m);
    return PSX_OK;
}

int psx_mcx_import(int device, int n_devices, uint64_t nbytes, int fd, uint64_t *out_id)
{
    if (!out_id || fd < 0) return fail(PSX_EINVAL, "bad descriptor / null out id");
    McMember *m = nullptr;
    CUmulticastObjectProp prop;
    int rc = mcx_common(device, n_devices, nbytes, &m, &prop);
    if (rc) return rc;
    CUresult r = g_drv.MemImportFromShareableHandle(&m->mc, (void *)(uintptr_t)fd,
                                                    CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    if (r != CUDA_SUCCESS) {
        mcx_release(m);
        return drv_fail("cuMemImportFromShareableHandle(multicast)", r);
    }
    *out_id = mcx_register(m);
    return PSX_OK;
}

int psx_mcx_add_device(uint64_t id)
{
    McMember *m = find(g_mcx, id);
    if (!m) return fail(PSX_EINVAL, "unknown multicast member id");
    if (m->added) return PSX_OK;
    CUdevice d;
    DRV_TRY(g_drv.DeviceGet(&d, m->device), "cuDeviceGet");
    DRV_TRY(g_drv.MulticastAddDevice(m->mc, d), "cuMulticastAddDevice");
    m->added = true;
    return PSX_OK;
}

/* Call after EVERY member has added its device (host barrier in between):
 * cuMulticastBindMem needs the complete team. */
int psx_mcx_bind(uint64_t id, void **out_unicast, void **out_multicast, uint64_t *out_size)
{
    McMember *m = find(g_mcx, id);
    if (!m) return fail(PSX_EINVAL, "unknown multicast member id");
    if (!m->added) return fail(PSX_ESTATE, "psx_mcx_add_device first");
    if (!m->bound) {
        PSX_DEVICE(m->device);
        CUmemAllocationProp ap;
        memset(&ap, 0, sizeof(ap));
        ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        ap.location.id = m->device;
        ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        DRV_TRY(g_drv.MemCreate(&m->mem, m->size, &ap, 0), "cuMemCreate");
        DRV_TRY(g_drv.MulticastBindMem(m->mc, 0, m->mem, 0, m->size, 0), "cuMulticastBindMem");
        m->bound = true;
        CUmemAccessDesc access;
        access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        access.location.id = m->device;
        access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        DRV_TRY(g_drv.MemAddressReserve(&m->uc, m->size, m->gran, 0, 0), "cuMemAddressReserve(unicast)");
        DRV_TRY(g_drv.MemMap(m->uc, m->size, 0, m->mem, 0), "cuMemMap(unicast)");
        DRV_TRY(g_drv.MemSetAccess(m->uc, m->size, &access, 1), "cuMemSetAccess(unicast)");
        DRV_TRY(g_drv.MemAddressReserve(&m->mcva, m->size, m->gran, 0, 0), "cuMemAddressReserve(multicast)");
        DRV_TRY(g_drv.MemMap(m->mcva, m->size, 0, m->mc, 0), "cuMemMap(multicast)");
        DRV_TRY(g_drv.MemSetAccess(m->mcva, m->size, &access, 1), "cuMemSetAccess(multicast)");
        CU_TRY(cudaMemset((void *)m->uc, 0, m->size));
        CU_TRY(cudaDeviceSynchronize());
    }
    if (out_unicast) *out_unicast = (void *)m->uc;
    if (out_multicast) *out_multicast = (void *)m->mcva;
    if (out_size) *out_size = m->size;
    return PSX_OK;
}

int psx_mcx_destroy(uint64_t id)
{
    McMember *m = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_mcx.find(id);
        if (it == g_mcx.end()) return fail(PSX_EINVAL, "unknown multicast member id");
        m = it->second;
        g_mcx.erase(it);
    }
    cudaSetDevice(m->device);
    cudaError_t e = cudaDeviceSynchronize();
    mcx_release(m);
    if (e != cudaSuccess) return fail(PSX_ECUDA, "draining the device before unbinding: %s", cudaGetErrorString(e));
    return PSX_OK;
}

/* The NVLS form of psx_round_bind: instead of W unicast peer mappings, the shard
 * gets the multicast addresses of its range inside the workers' arena --
 * grad_off_bytes / param_off_bytes locate the bucket's gradient and parameter
 * tensors in the arena (identical in every member), elem_off the shard's first
 * element inside the bucket.  psx_round / psx_round_counted then gather with
 * multimem.ld_reduce and scatter with multimem.st.  The shard's device must be a
 * member (it issues the multimem operations through its own mapping). */
int psx_round_bind_mc(uint64_t shard_id, uint64_t mcx_id, uint64_t grad_off_bytes,
                      uint64_t param_off_bytes, uint64_t elem_off, int n_members)
{
    Shard *s = find(g_shards, shard_id);
    if (!s) return fail(PSX_EINVAL, "unknown shard id");
    McMember *m = find(g_mcx, mcx_id);
    if (!m) return fail(PSX_EINVAL, "unknown multicast member id");
    if (!m->bound) return fail(PSX_ESTATE, "psx_mcx_bind first");
    if (m->device != s->device) return fail(PSX_EINVAL, "the multicast member must live on the shard's device");
    if (s->lay.wire != PSX_F32) return fail(PSX_ESTATE, "the NVLS round needs an f32 wire format");
    if (elem_off % 4 || grad_off_bytes % 16 || param_off_bytes % 16)
        return fail(PSX_EINVAL, "offsets must be 16-byte granular");
    if (n_members < 1 || n_members > m->n_devices) return fail(PSX_EINVAL, "n_members %d outside 1..%d", n_members, m->n_devices);
    const uint64_t span = (elem_off + s->lay.nelem_pad) * 4;
    if (grad_off_bytes + span > m->size || param_off_bytes + span > m->size)
        return fail(PSX_EINVAL, "shard range [%llu,+%llu) elements does not fit the %zu-byte arena",
                    (unsigned long long)elem_off, (unsigned long long)s->lay.nelem_pad, m->size);
    s->mc_grad = (const float *)(m->mcva + grad_off_bytes) + elem_off;
    s->mc_param = (float *)(m->mcva + param_off_bytes) + elem_off;
    s->mc_members = n_members;
    return PSX_OK;
}

}  // extern "C"

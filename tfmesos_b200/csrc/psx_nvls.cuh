// psx_nvls.cuh -- NVSwitch multicast (NVLS) primitives, single-process form.
// Included at the end of psx.cu.  EXPERIMENTAL: measured on 2 GPUs only
// (profiles/); not used by psx_round (DESIGN.md section 6 explains why the
// striped PS round gains from NVLS only from N = 4 GPUs and only as a
// tolerance-checked mode).
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

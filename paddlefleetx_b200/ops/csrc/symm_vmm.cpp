// Symmetric-memory runtime on the CUDA virtual-memory-management API (cuMem*) with NVLink-SHARP multicast objects.
//
// One *arena* = one physical allocation per rank (cuMemCreate, exportable as a POSIX file descriptor) that every rank of
// a group maps into its own address space (unicast peer pointers: plain ld/st over NVLink), plus — when the device and
// the fabric support it — ONE multicast object that all ranks bind their allocation to: a store to the multicast address
// is replicated by the NVSwitch into every rank's copy (multimem.st / multimem.red), a multimem.ld_reduce returns the
// in-switch sum of all copies.  The kernels in comm_nvls.cu are written against those two address kinds.
//
// This file is host-only and torch-free; descriptors travel between processes as file descriptors (SCM_RIGHTS over a
// unix socket, done by parallel/symmetric_memory.py).  Driver entry points are resolved at run time so the library
// still loads on a machine without libcuda (the CPU build box).
//
// Role in the reference: none — Paddle's sharding / TP collectives are NCCL calls (SURVEY §2.5, §5.8 describes the
// target design).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "pfx_symm.h"

namespace pfx {
namespace vmm {

namespace {

struct Api {
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
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  bool ok = false, mc_ok = false;
};

template <typename F>
bool resolve(const char* name, F* out) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr) { (void)cudaGetLastError(); return false; }
  *out = reinterpret_cast<F>(p);
  return true;
}

Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    bool ok = true;
    ok &= resolve("cuMemCreate", &a.MemCreate);
    ok &= resolve("cuMemRelease", &a.MemRelease);
    ok &= resolve("cuMemAddressReserve", &a.MemAddressReserve);
    ok &= resolve("cuMemAddressFree", &a.MemAddressFree);
    ok &= resolve("cuMemMap", &a.MemMap);
    ok &= resolve("cuMemUnmap", &a.MemUnmap);
    ok &= resolve("cuMemSetAccess", &a.MemSetAccess);
    ok &= resolve("cuMemExportToShareableHandle", &a.MemExport);
    ok &= resolve("cuMemImportFromShareableHandle", &a.MemImport);
    ok &= resolve("cuMemGetAllocationGranularity", &a.MemGetGranularity);
    ok &= resolve("cuDeviceGetAttribute", &a.DeviceGetAttribute);
    ok &= resolve("cuDeviceGet", &a.DeviceGet);
    resolve("cuGetErrorString", &a.GetErrorString);
    a.ok = ok;
    bool mc = ok;
    mc &= resolve("cuMulticastCreate", &a.McCreate);
    mc &= resolve("cuMulticastAddDevice", &a.McAddDevice);
    mc &= resolve("cuMulticastBindMem", &a.McBindMem);
    mc &= resolve("cuMulticastUnbind", &a.McUnbind);
    mc &= resolve("cuMulticastGetGranularity", &a.McGetGranularity);
    a.mc_ok = mc;
  });
  return a;
}

std::string err_str(const char* what, CUresult r) {
  const char* s = nullptr;
  if (api().GetErrorString) api().GetErrorString(r, &s);
  char buf[256];
  snprintf(buf, sizeof(buf), "%s failed: %s (CUresult %d)", what, s ? s : "?", (int)r);
  return buf;
}

#define VMM_TRY(call, what)                      \
  do {                                           \
    CUresult _r = (call);                        \
    if (_r != CUDA_SUCCESS) {                    \
      if (err) *err = err_str(what, _r);         \
      return false;                              \
    }                                            \
  } while (0)

struct Mapping { CUdeviceptr va; size_t size; CUmemGenericAllocationHandle handle; bool owns_handle; };
std::mutex g_mu;
std::unordered_map<int64_t, Mapping> g_maps;       // by base address: unicast (own + peers) and multicast mappings
std::unordered_map<int64_t, CUmemGenericAllocationHandle> g_mc;   // multicast object handle by id
int64_t g_next_mc = 1;

bool cur_device(CUdevice* dev, int* ordinal, std::string* err) {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) { if (err) *err = "cudaGetDevice failed"; return false; }
  (void)cudaFree(nullptr);        // make sure the primary context exists
  VMM_TRY(api().DeviceGet(dev, d), "cuDeviceGet");
  if (ordinal) *ordinal = d;
  return true;
}

CUmemAllocationProp alloc_prop(int ordinal) {
  CUmemAllocationProp p;
  memset(&p, 0, sizeof(p));
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = ordinal;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

bool map_rw(CUmemGenericAllocationHandle h, size_t size, size_t align, int ordinal, CUdeviceptr* out, std::string* err) {
  CUdeviceptr va = 0;
  VMM_TRY(api().MemAddressReserve(&va, size, align, 0, 0), "cuMemAddressReserve");
  CUresult r = api().MemMap(va, size, 0, h, 0);
  if (r != CUDA_SUCCESS) { api().MemAddressFree(va, size); if (err) *err = err_str("cuMemMap", r); return false; }
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = ordinal;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = api().MemSetAccess(va, size, &acc, 1);
  if (r != CUDA_SUCCESS) { api().MemUnmap(va, size); api().MemAddressFree(va, size); if (err) *err = err_str("cuMemSetAccess", r); return false; }
  *out = va;
  return true;
}

}  // namespace

Caps query_caps() {
  Caps c;
  Api& a = api();
  if (!a.ok) return c;
  CUdevice dev;
  int ord = 0;
  std::string e;
  if (!cur_device(&dev, &ord, &e)) return c;
  int v = 0;
  if (a.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev) == CUDA_SUCCESS) c.vmm = v != 0;
  v = 0;
  if (a.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev) == CUDA_SUCCESS) c.fd_export = v != 0;
  v = 0;
  if (a.mc_ok && a.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS) c.multicast = v != 0;
  CUmemAllocationProp p = alloc_prop(ord);
  size_t g = 0;
  if (a.MemGetGranularity(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS) c.granularity = g;
  return c;
}

size_t multicast_granularity(int world, size_t bytes, bool recommended) {
  Api& a = api();
  if (!a.mc_ok) return 0;
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof(mp));
  mp.numDevices = (unsigned)world;
  mp.size = bytes;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t g = 0;
  if (a.McGetGranularity(&g, &mp, recommended ? CU_MULTICAST_GRANULARITY_RECOMMENDED : CU_MULTICAST_GRANULARITY_MINIMUM) != CUDA_SUCCESS) return 0;
  return g;
}

bool arena_alloc(size_t bytes, int64_t* ptr, int* fd, std::string* err) {
  Api& a = api();
  if (!a.ok) { if (err) *err = "CUDA VMM driver entry points unavailable"; return false; }
  CUdevice dev;
  int ord = 0;
  if (!cur_device(&dev, &ord, err)) return false;
  CUmemAllocationProp p = alloc_prop(ord);
  CUmemGenericAllocationHandle h;
  VMM_TRY(a.MemCreate(&h, bytes, &p, 0), "cuMemCreate");
  int out_fd = -1;
  CUresult r = a.MemExport(&out_fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) { a.MemRelease(h); if (err) *err = err_str("cuMemExportToShareableHandle", r); return false; }
  CUdeviceptr va = 0;
  if (!map_rw(h, bytes, 0, ord, &va, err)) { a.MemRelease(h); return false; }
  if (cudaMemset(reinterpret_cast<void*>(va), 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    if (err) *err = "cudaMemset of a fresh arena failed";
    return false;
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_maps[(int64_t)va] = Mapping{va, bytes, h, true};
  }
  *ptr = (int64_t)va;
  *fd = out_fd;
  return true;
}

bool arena_import(int fd, size_t bytes, int64_t* ptr, std::string* err) {
  Api& a = api();
  CUdevice dev;
  int ord = 0;
  if (!cur_device(&dev, &ord, err)) return false;
  CUmemGenericAllocationHandle h;
  VMM_TRY(a.MemImport(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "cuMemImportFromShareableHandle");
  CUdeviceptr va = 0;
  if (!map_rw(h, bytes, 0, ord, &va, err)) { a.MemRelease(h); return false; }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_maps[(int64_t)va] = Mapping{va, bytes, h, true};
  }
  *ptr = (int64_t)va;
  return true;
}

bool mc_create(size_t bytes, int world, int64_t* mc_id, int* fd, std::string* err) {
  Api& a = api();
  if (!a.mc_ok) { if (err) *err = "multicast driver entry points unavailable"; return false; }
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof(mp));
  mp.numDevices = (unsigned)world;
  mp.size = bytes;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle h;
  VMM_TRY(a.McCreate(&h, &mp), "cuMulticastCreate");
  int out_fd = -1;
  CUresult r = a.MemExport(&out_fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) { a.MemRelease(h); if (err) *err = err_str("cuMemExportToShareableHandle(multicast)", r); return false; }
  std::lock_guard<std::mutex> lk(g_mu);
  *mc_id = g_next_mc++;
  g_mc[*mc_id] = h;
  *fd = out_fd;
  return true;
}

bool mc_import(int fd, int64_t* mc_id, std::string* err) {
  Api& a = api();
  if (!a.mc_ok) { if (err) *err = "multicast driver entry points unavailable"; return false; }
  CUmemGenericAllocationHandle h;
  VMM_TRY(a.MemImport(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "cuMemImportFromShareableHandle(multicast)");
  std::lock_guard<std::mutex> lk(g_mu);
  *mc_id = g_next_mc++;
  g_mc[*mc_id] = h;
  return true;
}

bool mc_add_device(int64_t mc_id, std::string* err) {
  Api& a = api();
  CUdevice dev;
  if (!cur_device(&dev, nullptr, err)) return false;
  CUmemGenericAllocationHandle h;
  { std::lock_guard<std::mutex> lk(g_mu); auto it = g_mc.find(mc_id); if (it == g_mc.end()) { if (err) *err = "unknown multicast id"; return false; } h = it->second; }
  VMM_TRY(a.McAddDevice(h, dev), "cuMulticastAddDevice");
  return true;
}

bool mc_bind_and_map(int64_t mc_id, int64_t local_arena_ptr, size_t bytes, int64_t* mc_ptr, std::string* err) {
  Api& a = api();
  CUdevice dev;
  int ord = 0;
  if (!cur_device(&dev, &ord, err)) return false;
  CUmemGenericAllocationHandle mc, mem;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_mc.find(mc_id);
    auto im = g_maps.find(local_arena_ptr);
    if (it == g_mc.end() || im == g_maps.end()) { if (err) *err = "unknown multicast id / arena"; return false; }
    mc = it->second; mem = im->second.handle;
  }
  VMM_TRY(a.McBindMem(mc, 0, mem, 0, bytes, 0), "cuMulticastBindMem");
  CUdeviceptr va = 0;
  if (!map_rw(mc, bytes, 0, ord, &va, err)) return false;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_maps[(int64_t)va] = Mapping{va, bytes, mc, false};
  }
  *mc_ptr = (int64_t)va;
  return true;
}

void unmap(int64_t ptr) {
  Api& a = api();
  Mapping m;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_maps.find(ptr);
    if (it == g_maps.end()) return;
    m = it->second;
    g_maps.erase(it);
  }
  a.MemUnmap(m.va, m.size);
  a.MemAddressFree(m.va, m.size);
  if (m.owns_handle) a.MemRelease(m.handle);
}

}  // namespace vmm
}  // namespace pfx

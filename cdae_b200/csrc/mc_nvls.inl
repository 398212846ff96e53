// cdae_b200/csrc/mc_nvls.inl — host side of the NVLS (NVSwitch multicast) mode of the combine step,
// included at the end of api.cu.  The gradient buffers, the item-side parameter buffer and the barrier
// counters of every rank live in ONE virtual-memory-management allocation per GPU, all bound at offset 0
// of one multicast object, so that inside p2p::mc_step_kernel
//   multimem.ld_reduce [mc + off]  returns the SUM over all ranks of the word at `off` (reduced in the switch:
//                                  a rank pulls 1/G of the buffer once instead of 1/G from each of G-1 peers),
//   multimem.st        [mc + off]  writes every rank's copy (the all-gather of the updated parameters),
//   multimem.red       [mc + off]  bumps every rank's barrier counter.
// Driver entry points are fetched with cudaGetDriverEntryPoint (no link-time dependency on libcuda).
#include <cuda.h>

namespace {
struct McApi {
  bool loaded = false;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemExport)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImport)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
};
McApi g_mc;

template <class F>
static bool mc_sym(F& fn, const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
  fn = reinterpret_cast<F>(p);
  return true;
}
static int mc_load() {
  if (g_mc.loaded) return 0;
  bool ok = mc_sym(g_mc.MulticastCreate, "cuMulticastCreate") && mc_sym(g_mc.MulticastAddDevice, "cuMulticastAddDevice") &&
            mc_sym(g_mc.MulticastBindMem, "cuMulticastBindMem") && mc_sym(g_mc.MulticastGetGranularity, "cuMulticastGetGranularity") &&
            mc_sym(g_mc.MemCreate, "cuMemCreate") && mc_sym(g_mc.MemRelease, "cuMemRelease") &&
            mc_sym(g_mc.MemExport, "cuMemExportToShareableHandle") && mc_sym(g_mc.MemImport, "cuMemImportFromShareableHandle") &&
            mc_sym(g_mc.MemAddressReserve, "cuMemAddressReserve") && mc_sym(g_mc.MemAddressFree, "cuMemAddressFree") &&
            mc_sym(g_mc.MemMap, "cuMemMap") && mc_sym(g_mc.MemUnmap, "cuMemUnmap") && mc_sym(g_mc.MemSetAccess, "cuMemSetAccess") &&
            mc_sym(g_mc.MemGetAllocationGranularity, "cuMemGetAllocationGranularity") &&
            mc_sym(g_mc.DeviceGet, "cuDeviceGet") && mc_sym(g_mc.DeviceGetAttribute, "cuDeviceGetAttribute");
  if (!ok) return set_error(CDAE_E_STATE, "this CUDA driver lacks the multicast / virtual-memory entry points");
  g_mc.loaded = true;
  return 0;
}
#define MC(call)                                                                                        \
  do {                                                                                                  \
    CUresult r__ = (call);                                                                              \
    if (r__ != CUDA_SUCCESS) return set_error(CDAE_E_STATE, "%s failed: CUresult %d (multicast / NVLS unavailable?)", #call, (int)r__); \
  } while (0)

// layout of the per-rank allocation (bytes): [flags 4 KB | item-side parameters | gradients x2], padded to the granularity
static size_t mc_bytes_needed(cdae_handle* h) { return 4096 + 3 * sizeof(float) * h->grad_floats; }

static int mc_props(cdae_handle* h, CUmulticastObjectProp* mp, size_t* size_out) {
  memset(mp, 0, sizeof(*mp));
  mp->numDevices = (unsigned)h->world;
  mp->handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  mp->size = mc_bytes_needed(h);
  size_t gran = 0;
  MC(g_mc.MulticastGetGranularity(&gran, mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
  if (gran == 0) gran = 2u << 20;
  mp->size = (mp->size + gran - 1) / gran * gran;
  *size_out = mp->size;
  return 0;
}
static int mc_check_support(cdae_handle* h) {
  TRY(mc_load());
  if (h->world < 2 || h->world > p2p::MAX_RANKS) return set_error(CDAE_E_STATE, "needs a process group of 2..%d ranks (cdae_dist_init first)", p2p::MAX_RANKS);
  CUdevice dev;
  MC(g_mc.DeviceGet(&dev, h->cfg.device));
  int sup = 0;
  MC(g_mc.DeviceGetAttribute(&sup, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev));
  if (!sup) return set_error(CDAE_E_STATE, "device %d does not support multicast (no NVSwitch / NVLS)", h->cfg.device);
  return 0;
}
}  // namespace

extern "C" {

int cdae_dist_mc_create(cdae_handle* h, int32_t* fd_out) {
  if (!h || !fd_out) return set_error(CDAE_E_INVALID, "NULL argument");
  CU(cudaSetDevice(h->cfg.device));
  TRY(mc_check_support(h));
  CUmulticastObjectProp mp;
  size_t size = 0;
  TRY(mc_props(h, &mp, &size));
  CUmemGenericAllocationHandle mc = 0;
  MC(g_mc.MulticastCreate(&mc, &mp));
  int fd = -1;
  MC(g_mc.MemExport(&fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  h->mc_handle = (unsigned long long)mc;
  h->mc_size = size;
  h->mc_creator = true;
  *fd_out = fd;
  return 0;
}

int cdae_dist_mc_attach(cdae_handle* h, int32_t fd) {
  if (!h) return set_error(CDAE_E_INVALID, "handle is NULL");
  CU(cudaSetDevice(h->cfg.device));
  TRY(mc_check_support(h));
  if (!h->mc_creator) {
    if (fd < 0) return set_error(CDAE_E_INVALID, "need the file descriptor exported by cdae_dist_mc_create on rank 0");
    CUmulticastObjectProp mp;
    size_t size = 0;
    TRY(mc_props(h, &mp, &size));
    CUmemGenericAllocationHandle mc = 0;
    MC(g_mc.MemImport(&mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    h->mc_handle = (unsigned long long)mc;
    h->mc_size = size;
  }
  CUdevice dev;
  MC(g_mc.DeviceGet(&dev, h->cfg.device));
  MC(g_mc.MulticastAddDevice((CUmemGenericAllocationHandle)h->mc_handle, dev));
  h->mc_attached = true;
  return 0;
}

int cdae_dist_mc_bind(cdae_handle* h) {
  if (!h) return set_error(CDAE_E_INVALID, "handle is NULL");
  if (!h->mc_attached) return set_error(CDAE_E_STATE, "cdae_dist_mc_attach has not run on this rank");
  if (h->p2p_on) return set_error(CDAE_E_STATE, "peer-memory mode is already active");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));
  const size_t size = h->mc_size;
  CUdevice dev;
  MC(g_mc.DeviceGet(&dev, h->cfg.device));
  CUmemAllocationProp ap;
  memset(&ap, 0, sizeof(ap));
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = dev;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle phys = 0;
  MC(g_mc.MemCreate(&phys, size, &ap, 0));
  MC(g_mc.MulticastBindMem((CUmemGenericAllocationHandle)h->mc_handle, 0, phys, 0, size, 0));
  CUmemAccessDesc ad;
  memset(&ad, 0, sizeof(ad));
  ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ad.location.id = dev;
  ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CUdeviceptr uc = 0, mcva = 0;
  MC(g_mc.MemAddressReserve(&uc, size, size < (512u << 20) ? (2u << 20) : 0, 0, 0));
  MC(g_mc.MemMap(uc, size, 0, phys, 0));
  MC(g_mc.MemSetAccess(uc, size, &ad, 1));
  MC(g_mc.MemAddressReserve(&mcva, size, size < (512u << 20) ? (2u << 20) : 0, 0, 0));
  MC(g_mc.MemMap(mcva, size, 0, (CUmemGenericAllocationHandle)h->mc_handle, 0));
  MC(g_mc.MemSetAccess(mcva, size, &ad, 1));
  h->mc_phys = (unsigned long long)phys;
  h->mc_uc = (char*)uc;
  h->mc_mc = (char*)mcva;
  // move the item side into the new allocation: [flags | parameters | gradients x2]
  CU(cudaMemsetAsync(h->mc_uc, 0, size, h->stream));
  float* new_params = reinterpret_cast<float*>(h->mc_uc + 4096);
  float* new_grad = new_params + h->grad_floats;
  CU(cudaMemcpyAsync(new_params, h->item_params.p, sizeof(float) * h->grad_floats, cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->item_params.release();
  h->grad.release();
  h->item_params.p = new_params; h->item_params.cap = h->grad_floats;   // owned by the VMM block (mc_active): never cudaFree'd
  h->grad.p = new_grad; h->grad.cap = 2 * h->grad_floats;
  h->mc_active = true;
  h->p2p_fused = true;
  h->p2p_parity = 0;
  h->m.steps_slot = 0;
  h->m.direct_lambda = 1;
  point_item_side(h, h->grad.p);
  CU(cudaMalloc(&h->p2p_done, sizeof(unsigned int)));
  CU(cudaMemset(h->p2p_done, 0, sizeof(unsigned int)));
  h->p2p_epoch = 0;
  h->p2p_on = true;       // callers must run a barrier between the last rank's bind and the first training call
  return 0;
}

}  // extern "C"

// (called from cdae_destroy)
static void mc_release(cdae_handle* h) {
  if (!h->mc_active && !h->mc_attached) return;
  if (h->mc_active) {
    h->item_params.p = nullptr; h->item_params.cap = 0;
    h->grad.p = nullptr; h->grad.cap = 0;
    if (h->p2p_done) cudaFree(h->p2p_done);
    h->p2p_done = nullptr;
    g_mc.MemUnmap((CUdeviceptr)h->mc_mc, h->mc_size);
    g_mc.MemUnmap((CUdeviceptr)h->mc_uc, h->mc_size);
    g_mc.MemAddressFree((CUdeviceptr)h->mc_mc, h->mc_size);
    g_mc.MemAddressFree((CUdeviceptr)h->mc_uc, h->mc_size);
    g_mc.MemRelease((CUmemGenericAllocationHandle)h->mc_phys);
  }
  if (h->mc_handle && !h->mc_shared) g_mc.MemRelease((CUmemGenericAllocationHandle)h->mc_handle);
  h->mc_active = h->mc_attached = false;
}

// Host runtime of libtstore_b200: error strings, per-device copy streams, events, region
// export/import (CUDA IPC + same-process P2P), the HBM arena of the storage volume and pinned
// host staging.  Everything here is plumbing around the copy_rects kernel; see
// include/tstore_b200.h for the reference call sites each entry point replaces.

#include <cuda.h>
#include <unistd.h>

#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <random>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "tsb_internal.h"

namespace tsb {

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string t_last_error;

void set_error(const std::string& msg) { t_last_error = msg; }
int fail(int code, const std::string& msg) {
  t_last_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  t_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  cudaGetLastError();  // clear the sticky-less error so later calls start clean
  return TSB_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// DeviceGuard
// ---------------------------------------------------------------------------------------------
namespace {
typedef CUresult (*PFN_cuCtxGetCurrent)(CUcontext*);
PFN_cuCtxGetCurrent ctx_get_current() {
  static PFN_cuCtxGetCurrent fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuCtxGetCurrent", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<PFN_cuCtxGetCurrent>(p);
  }();
  return fn;
}
}  // namespace

DeviceGuard::DeviceGuard(int device) {
  if (PFN_cuCtxGetCurrent get = ctx_get_current()) {
    CUcontext c = nullptr;
    had_context = get(&c) == CUDA_SUCCESS && c != nullptr;
  }
  err = cudaGetDevice(&prev);
  if (err != cudaSuccess) { ok = false; return; }
  // always set: on a thread that never touched CUDA (e.g. the actor-server thread) this is what
  // binds the primary context, which the driver-API calls below rely on
  err = cudaSetDevice(device);
  if (err != cudaSuccess) ok = false;
  cur = device;
}

// ---------------------------------------------------------------------------------------------
// global state
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kMaxDevices = 64;

struct DeviceState {
  bool stream_ready = false;
  cudaStream_t stream = nullptr;
  int sm_count = 0;
};

struct ImportEntry {
  void* base = nullptr;  // mapped base of the exporter's allocation
  int opened_on = -1;
  uint64_t epoch = 0;
  std::array<uint64_t, 3> where{};  // (boot_id, exporter device, exporter base address)
};

struct Arena {
  int device = 0;
  char* base = nullptr;
  uint64_t capacity = 0;
  uint64_t in_use = 0;
  uint64_t high_water = 0;
  uint64_t cursor = 0;
  std::map<uint64_t, uint64_t> free_blocks;           // offset -> size
  std::unordered_map<uint64_t, uint64_t> used_blocks; // offset -> size
};

struct Global {
  std::mutex mu;
  bool inited = false;
  int ndev = 0;
  uint64_t boot_id = 0;
  DeviceState dev[kMaxDevices];
  std::set<std::pair<int, int>> peers;  // (device, peer) enabled by us or already enabled
  std::map<std::array<uint8_t, 72>, ImportEntry> imports;  // boot_id(8) + handle(64)
  // (boot_id, exporter device, exporter base address) -> key in `imports`: one live allocation per
  // address, so a second handle for the same address means the first allocation is gone
  std::map<std::array<uint64_t, 3>, std::array<uint8_t, 72>> import_by_address;
  uint64_t stale_evictions = 0;
  std::unordered_map<uint64_t, Arena*> arenas;
  uint64_t next_id = 1;
};

Global& G() {
  static Global g;
  return g;
}

typedef CUresult (*PFN_cuMemGetAddressRange)(CUdeviceptr*, size_t*, CUdeviceptr);
typedef CUresult (*PFN_cuPointerGetAttribute)(void*, CUpointer_attribute, CUdeviceptr);

template <typename Fn>
Fn driver_fn(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return reinterpret_cast<Fn>(fn);
}

int ensure_init_locked(Global& g) {
  if (g.inited) return TSB_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  if (n <= 0) return fail(TSB_ERR_CUDA, "tstore_b200: no CUDA device visible (this library has no CPU fallback)");
  if (n > kMaxDevices) n = kMaxDevices;
  g.ndev = n;
  std::random_device rd;
  g.boot_id = (static_cast<uint64_t>(rd()) << 32) ^ rd() ^ (static_cast<uint64_t>(getpid()) << 20);
  if (g.boot_id == 0) g.boot_id = 1;
  g.inited = true;
  return TSB_OK;
}

int check_device(Global& g, int device) {
  if (device < 0 || device >= g.ndev) return fail(TSB_ERR_INVALID, "invalid device ordinal " + std::to_string(device));
  return TSB_OK;
}

int enable_peer_locked(Global& g, int device, int peer) {
  if (device == peer) return TSB_OK;
  if (g.peers.count({device, peer})) return TSB_OK;
  int can = 0;
  TSB_CUDA(cudaDeviceCanAccessPeer(&can, device, peer));
  if (!can) return fail(TSB_ERR_UNSUPPORTED, "device " + std::to_string(device) + " cannot access peer " + std::to_string(peer));
  DeviceGuard guard(device);
  if (!guard.ok) return cuda_fail(guard.err, "cudaSetDevice");
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
  } else if (e != cudaSuccess) {
    return cuda_fail(e, "cudaDeviceEnablePeerAccess");
  }
  g.peers.insert({device, peer});
  return TSB_OK;
}

}  // namespace

int device_sm_count(int device, int* out) {
  Global& g = G();
  std::lock_guard<std::mutex> lk(g.mu);
  int st = ensure_init_locked(g);
  if (st) return st;
  if ((st = check_device(g, device))) return st;
  if (g.dev[device].sm_count == 0) {
    int n = 0;
    TSB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
    g.dev[device].sm_count = n;
  }
  *out = g.dev[device].sm_count;
  return TSB_OK;
}

int copy_stream(int device, cudaStream_t* out) {
  Global& g = G();
  std::lock_guard<std::mutex> lk(g.mu);
  int st = ensure_init_locked(g);
  if (st) return st;
  if ((st = check_device(g, device))) return st;
  DeviceState& d = g.dev[device];
  if (!d.stream_ready) {
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(guard.err, "cudaSetDevice");
    TSB_CUDA(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    d.stream_ready = true;
  }
  *out = d.stream;
  return TSB_OK;
}

cudaStream_t resolve_stream(int device, void* stream, int* status) {
  *status = TSB_OK;
  if (stream != nullptr) return reinterpret_cast<cudaStream_t>(stream);
  cudaStream_t s = nullptr;
  *status = copy_stream(device, &s);
  return s;
}

}  // namespace tsb

using namespace tsb;

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int tsb_abi_version(void) { return TSB_ABI_VERSION; }

const char* tsb_last_error(void) { return t_last_error.c_str(); }

int tsb_init(void) {
  Global& g = G();
  std::lock_guard<std::mutex> lk(g.mu);
  return ensure_init_locked(g);
}

int tsb_device_count(int* out_n) {
  if (!out_n) return fail(TSB_ERR_INVALID, "out_n is NULL");
  Global& g = G();
  std::lock_guard<std::mutex> lk(g.mu);
  int st = ensure_init_locked(g);
  if (st) return st;
  *out_n = g.ndev;
  return TSB_OK;
}

int tsb_enable_peer_access(int device, int peer_device) {
  Global& g = G();
  std::lock_guard<std::mutex> lk(g.mu);
  int st = ensure_init_locked(g);
  if (st) return st;
  if ((st = check_device(g, device))) return st;
  if ((st = check_device(g, peer_device))) return st;
  if ((st = enable_peer_locked(g, device, peer_device))) return st;
  return enable_peer_locked(g, peer_device, device);
}

// ---------------------------------------------------------------------------------------------
// regions
// ---------------------------------------------------------------------------------------------
int tsb_export_region(const void* ptr, uint64_t nbytes, tsb_region_t* out) {
  if (!ptr || !out) return fail(TSB_ERR_INVALID, "tsb_export_region: NULL argument");
  Global& g = G();
  {
    std::lock_guard<std::mutex> lk(g.mu);
    int st = ensure_init_locked(g);
    if (st) return st;
  }
  cudaPointerAttributes attr;
  TSB_CUDA(cudaPointerGetAttributes(&attr, ptr));
  if (attr.type != cudaMemoryTypeDevice)
    return fail(TSB_ERR_INVALID, "tsb_export_region: pointer is not device memory");

  static PFN_cuMemGetAddressRange p_range = driver_fn<PFN_cuMemGetAddressRange>("cuMemGetAddressRange");
  static PFN_cuPointerGetAttribute p_attr = driver_fn<PFN_cuPointerGetAttribute>("cuPointerGetAttribute");
  if (!p_range) return fail(TSB_ERR_CUDA, "cuMemGetAddressRange entry point not found");

  DeviceGuard guard(attr.device);
  if (!guard.ok) return cuda_fail(guard.err, "cudaSetDevice");

  CUdeviceptr base = 0;
  size_t size = 0;
  CUresult r = p_range(&base, &size, reinterpret_cast<CUdeviceptr>(ptr));
  if (r != CUDA_SUCCESS) return fail(TSB_ERR_CUDA, "cuMemGetAddressRange failed with CUresult " + std::to_string(static_cast<int>(r)));
  const uint64_t off = reinterpret_cast<uint64_t>(ptr) - static_cast<uint64_t>(base);
  if (off + nbytes > size)
    return fail(TSB_ERR_INVALID, "tsb_export_region: [ptr, ptr+nbytes) crosses the end of its allocation");

  memset(out, 0, sizeof(*out));
  out->offset = off;
  out->nbytes = nbytes;
  out->alloc_bytes = size;
  out->local_ptr = reinterpret_cast<uint64_t>(ptr);
  out->device = attr.device;
  out->pid = static_cast<int32_t>(getpid());
  out->boot_id = g.boot_id;
  if (p_attr) {
    unsigned long long id = 0;
    if (p_attr(&id, CU_POINTER_ATTRIBUTE_BUFFER_ID, base) == CUDA_SUCCESS) out->epoch = id;
  }

  int legacy_ok = 1;
  if (p_attr) {
    int v = 0;
    if (p_attr(&v, CU_POINTER_ATTRIBUTE_IS_LEGACY_CUDA_IPC_CAPABLE, base) == CUDA_SUCCESS) legacy_ok = v;
  }
  if (!legacy_ok) {
    // Still usable inside this process (single-process multi-GPU); cross-process import will
    // fail loudly because the handle is all zeros.
    return TSB_OK;
  }
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));
  if (e != cudaSuccess) {
    cudaGetLastError();
    // leave the handle zeroed: same-process use keeps working
    return TSB_OK;
  }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out->ipc_handle, &h, 64);
  return TSB_OK;
}

int tsb_import_region(const tsb_region_t* region, int device, void** out_ptr) {
  if (!region || !out_ptr) return fail(TSB_ERR_INVALID, "tsb_import_region: NULL argument");
  Global& g = G();
  std::lock_guard<std::mutex> lk(g.mu);
  int st = ensure_init_locked(g);
  if (st) return st;
  if ((st = check_device(g, device))) return st;

  if (region->pid == static_cast<int32_t>(getpid()) && region->boot_id == g.boot_id) {
    // same process: no IPC, just make sure `device` can dereference the owner's memory
    if (region->device != device) {
      if ((st = check_device(g, region->device))) return st;
      if ((st = enable_peer_locked(g, device, region->device))) return st;
    }
    *out_ptr = reinterpret_cast<void*>(region->local_ptr);
    return TSB_OK;
  }

  bool zero = true;
  for (int i = 0; i < 64; ++i) zero = zero && region->ipc_handle[i] == 0;
  if (zero)
    return fail(TSB_ERR_UNSUPPORTED,
                "tsb_import_region: exporter could not create a CUDA IPC handle for this memory "
                "(VMM / expandable_segments allocations are not legacy-IPC capable; allocate the "
                "tensors with the default caching allocator or from a tstore arena)");

  std::array<uint8_t, 72> key;
  memcpy(key.data(), &region->boot_id, 8);
  memcpy(key.data() + 8, region->ipc_handle, 64);
  auto it = g.imports.find(key);
  if (it == g.imports.end()) {
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(guard.err, "cudaSetDevice");
    const std::array<uint64_t, 3> where{region->boot_id, static_cast<uint64_t>(region->device), region->local_ptr - region->offset};
    auto old = g.import_by_address.find(where);
    if (old != g.import_by_address.end()) {
      // the exporter freed the allocation we mapped and its address now belongs to a new one
      // (new handle, new epoch): the old mapping only pins dead memory -- drop it
      auto stale = g.imports.find(old->second);
      if (stale != g.imports.end()) {
        DeviceGuard g2(stale->second.opened_on);
        if (cudaIpcCloseMemHandle(stale->second.base) != cudaSuccess) cudaGetLastError();
        g.imports.erase(stale);
        ++g.stale_evictions;
      }
      g.import_by_address.erase(old);
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, region->ipc_handle, 64);
    void* base = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle");
    ImportEntry ent;
    ent.base = base;
    ent.opened_on = device;
    ent.epoch = region->epoch;
    ent.where = where;
    it = g.imports.emplace(key, ent).first;
    g.import_by_address[where] = key;
  }
  *out_ptr = static_cast<char*>(it->second.base) + region->offset;
  return TSB_OK;
}

int tsb_release_region(const tsb_region_t* region) {
  if (!region) return fail(TSB_ERR_INVALID, "tsb_release_region: NULL argument");
  Global& g = G();
  std::lock_guard<std::mutex> lk(g.mu);
  if (!g.inited) return TSB_OK;
  if (region->pid == static_cast<int32_t>(getpid()) && region->boot_id == g.boot_id) return TSB_OK;
  std::array<uint8_t, 72> key;
  memcpy(key.data(), &region->boot_id, 8);
  memcpy(key.data() + 8, region->ipc_handle, 64);
  auto it = g.imports.find(key);
  if (it == g.imports.end()) return TSB_OK;
  DeviceGuard guard(it->second.opened_on);
  cudaError_t e = cudaIpcCloseMemHandle(it->second.base);
  g.import_by_address.erase(it->second.where);
  g.imports.erase(it);
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcCloseMemHandle");
  return TSB_OK;
}

int tsb_import_stats(uint64_t* out_live, uint64_t* out_stale_evictions) {
  Global& g = G();
  std::lock_guard<std::mutex> lk(g.mu);
  if (out_live) *out_live = g.imports.size();
  if (out_stale_evictions) *out_stale_evictions = g.stale_evictions;
  return TSB_OK;
}

int tsb_release_all(void) {
  Global& g = G();
  std::lock_guard<std::mutex> lk(g.mu);
  if (!g.inited) return TSB_OK;
  int st = TSB_OK;
  for (auto& kv : g.imports) {
    DeviceGuard guard(kv.second.opened_on);
    cudaError_t e = cudaIpcCloseMemHandle(kv.second.base);
    if (e != cudaSuccess) st = cuda_fail(e, "cudaIpcCloseMemHandle");
  }
  g.imports.clear();
  g.import_by_address.clear();
  return st;
}

// ---------------------------------------------------------------------------------------------
// streams / events
// ---------------------------------------------------------------------------------------------
int tsb_stream_create(int device, void** out_stream) {
  if (!out_stream) return fail(TSB_ERR_INVALID, "out_stream is NULL");
  int n = 0;
  int st = tsb_device_count(&n);
  if (st) return st;
  if (device < 0 || device >= n) return fail(TSB_ERR_INVALID, "invalid device");
  DeviceGuard guard(device);
  if (!guard.ok) return cuda_fail(guard.err, "cudaSetDevice");
  cudaStream_t s;
  TSB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *out_stream = s;
  return TSB_OK;
}

int tsb_stream_destroy(void* stream) {
  if (!stream) return TSB_OK;
  TSB_CUDA(cudaStreamDestroy(reinterpret_cast<cudaStream_t>(stream)));
  return TSB_OK;
}

int tsb_copy_stream(int device, void** out_stream) {
  if (!out_stream) return fail(TSB_ERR_INVALID, "out_stream is NULL");
  cudaStream_t s;
  int st = copy_stream(device, &s);
  if (st) return st;
  *out_stream = s;
  return TSB_OK;
}

int tsb_stream_sync(int device, void* stream) {
  int st;
  cudaStream_t s = resolve_stream(device, stream, &st);
  if (st) return st;
  DeviceGuard guard(device);
  if (!guard.ok) return cuda_fail(guard.err, "cudaSetDevice");
  TSB_CUDA(cudaStreamSynchronize(s));
  return TSB_OK;
}

int tsb_event_create(int device, int timing, void** out_event) {
  if (!out_event) return fail(TSB_ERR_INVALID, "out_event is NULL");
  int n = 0;
  int st = tsb_device_count(&n);
  if (st) return st;
  if (device < 0 || device >= n) return fail(TSB_ERR_INVALID, "invalid device");
  DeviceGuard guard(device);
  if (!guard.ok) return cuda_fail(guard.err, "cudaSetDevice");
  cudaEvent_t ev;
  TSB_CUDA(cudaEventCreateWithFlags(&ev, timing ? cudaEventDefault : cudaEventDisableTiming));
  *out_event = ev;
  return TSB_OK;
}

int tsb_event_record(void* event, int device, void* stream) {
  if (!event) return fail(TSB_ERR_INVALID, "event is NULL");
  int st;
  cudaStream_t s = resolve_stream(device, stream, &st);
  if (st) return st;
  DeviceGuard guard(device);
  if (!guard.ok) return cuda_fail(guard.err, "cudaSetDevice");
  TSB_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(event), s));
  return TSB_OK;
}

int tsb_stream_wait_event(int device, void* stream, void* event) {
  if (!event) return fail(TSB_ERR_INVALID, "event is NULL");
  int st;
  cudaStream_t s = resolve_stream(device, stream, &st);
  if (st) return st;
  DeviceGuard guard(device);
  if (!guard.ok) return cuda_fail(guard.err, "cudaSetDevice");
  TSB_CUDA(cudaStreamWaitEvent(s, reinterpret_cast<cudaEvent_t>(event), 0));
  return TSB_OK;
}

int tsb_event_query(void* event, int* out_done) {
  if (!event || !out_done) return fail(TSB_ERR_INVALID, "NULL argument");
  cudaError_t e = cudaEventQuery(reinterpret_cast<cudaEvent_t>(event));
  if (e == cudaSuccess) {
    *out_done = 1;
    return TSB_OK;
  }
  if (e == cudaErrorNotReady) {
    cudaGetLastError();
    *out_done = 0;
    return TSB_OK;
  }
  return cuda_fail(e, "cudaEventQuery");
}

int tsb_event_sync(void* event) {
  if (!event) return fail(TSB_ERR_INVALID, "event is NULL");
  TSB_CUDA(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(event)));
  return TSB_OK;
}

int tsb_event_elapsed_ms(void* start, void* stop, float* out_ms) {
  if (!start || !stop || !out_ms) return fail(TSB_ERR_INVALID, "NULL argument");
  TSB_CUDA(cudaEventElapsedTime(out_ms, reinterpret_cast<cudaEvent_t>(start), reinterpret_cast<cudaEvent_t>(stop)));
  return TSB_OK;
}

int tsb_event_destroy(void* event) {
  if (!event) return TSB_OK;
  TSB_CUDA(cudaEventDestroy(reinterpret_cast<cudaEvent_t>(event)));
  return TSB_OK;
}

// ---------------------------------------------------------------------------------------------
// arena
// ---------------------------------------------------------------------------------------------
int tsb_arena_create(int device, uint64_t capacity_bytes, tsb_arena_t* out) {
  if (!out || capacity_bytes == 0) return fail(TSB_ERR_INVALID, "tsb_arena_create: bad argument");
  Global& g = G();
  std::lock_guard<std::mutex> lk(g.mu);
  int st = ensure_init_locked(g);
  if (st) return st;
  if ((st = check_device(g, device))) return st;
  DeviceGuard guard(device);
  if (!guard.ok) return cuda_fail(guard.err, "cudaSetDevice");
  void* base = nullptr;
  cudaError_t e = cudaMalloc(&base, capacity_bytes);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    return fail(TSB_ERR_NOMEM, "tsb_arena_create: cudaMalloc of " + std::to_string(capacity_bytes) + " bytes failed");
  }
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  Arena* a = new Arena();
  a->device = device;
  a->base = static_cast<char*>(base);
  a->capacity = capacity_bytes;
  a->free_blocks[0] = capacity_bytes;
  uint64_t id = g.next_id++;
  g.arenas[id] = a;
  *out = id;
  return TSB_OK;
}

int tsb_arena_alloc(tsb_arena_t arena, uint64_t nbytes, uint64_t align, void** out_ptr) {
  if (!out_ptr) return fail(TSB_ERR_INVALID, "out_ptr is NULL");
  Global& g = G();
  std::lock_guard<std::mutex> lk(g.mu);
  auto it = g.arenas.find(arena);
  if (it == g.arenas.end()) return fail(TSB_ERR_NOTFOUND, "unknown arena");
  Arena* a = it->second;
  if (align == 0) align = 256;
  if (align & (align - 1)) return fail(TSB_ERR_INVALID, "alignment must be a power of two");
  if (nbytes == 0) nbytes = 1;
  const uint64_t base_addr = reinterpret_cast<uint64_t>(a->base);
  // first fit starting at the rotating cursor, then wrap: ring-buffer behaviour for the steady
  // state of put/overwrite/delete cycles
  auto try_block = [&](std::map<uint64_t, uint64_t>::iterator b) -> bool {
    const uint64_t off = b->first, size = b->second;
    const uint64_t aligned = ((base_addr + off + align - 1) & ~(align - 1)) - base_addr;
    const uint64_t padded_end = aligned + ((nbytes + 255) & ~uint64_t(255));
    if (padded_end > off + size) return false;
    a->free_blocks.erase(b);
    if (aligned > off) a->free_blocks[off] = aligned - off;
    if (padded_end < off + size) a->free_blocks[padded_end] = off + size - padded_end;
    a->used_blocks[aligned] = padded_end - aligned;
    a->in_use += padded_end - aligned;
    if (a->in_use > a->high_water) a->high_water = a->in_use;
    a->cursor = padded_end;
    *out_ptr = a->base + aligned;
    return true;
  };
  auto start = a->free_blocks.lower_bound(a->cursor);
  for (auto b = start; b != a->free_blocks.end(); ++b)
    if (try_block(b)) return TSB_OK;
  for (auto b = a->free_blocks.begin(); b != start; ++b)
    if (try_block(b)) return TSB_OK;
  return fail(TSB_ERR_NOMEM, "arena exhausted: requested " + std::to_string(nbytes) + " bytes, in use " +
                                 std::to_string(a->in_use) + " of " + std::to_string(a->capacity));
}

int tsb_arena_free(tsb_arena_t arena, void* ptr) {
  Global& g = G();
  std::lock_guard<std::mutex> lk(g.mu);
  auto it = g.arenas.find(arena);
  if (it == g.arenas.end()) return fail(TSB_ERR_NOTFOUND, "unknown arena");
  Arena* a = it->second;
  const uint64_t off = static_cast<uint64_t>(static_cast<char*>(ptr) - a->base);
  auto u = a->used_blocks.find(off);
  if (u == a->used_blocks.end()) return fail(TSB_ERR_NOTFOUND, "tsb_arena_free: pointer was not allocated from this arena");
  uint64_t size = u->second;
  a->used_blocks.erase(u);
  a->in_use -= size;
  uint64_t start = off, end = off + size;
  auto next = a->free_blocks.lower_bound(start);
  if (next != a->free_blocks.end() && next->first == end) {
    end += next->second;
    next = a->free_blocks.erase(next);
  }
  if (next != a->free_blocks.begin()) {
    auto prev = std::prev(next);
    if (prev->first + prev->second == start) {
      start = prev->first;
      a->free_blocks.erase(prev);
    }
  }
  a->free_blocks[start] = end - start;
  return TSB_OK;
}

int tsb_arena_stats(tsb_arena_t arena, tsb_arena_stats_t* out) {
  if (!out) return fail(TSB_ERR_INVALID, "out is NULL");
  Global& g = G();
  std::lock_guard<std::mutex> lk(g.mu);
  auto it = g.arenas.find(arena);
  if (it == g.arenas.end()) return fail(TSB_ERR_NOTFOUND, "unknown arena");
  Arena* a = it->second;
  out->capacity = a->capacity;
  out->in_use = a->in_use;
  out->high_water = a->high_water;
  out->num_blocks = a->used_blocks.size();
  out->base = reinterpret_cast<uint64_t>(a->base);
  return TSB_OK;
}

int tsb_arena_destroy(tsb_arena_t arena) {
  Global& g = G();
  std::lock_guard<std::mutex> lk(g.mu);
  auto it = g.arenas.find(arena);
  if (it == g.arenas.end()) return fail(TSB_ERR_NOTFOUND, "unknown arena");
  Arena* a = it->second;
  g.arenas.erase(it);
  DeviceGuard guard(a->device);
  cudaError_t e = cudaFree(a->base);
  delete a;
  if (e != cudaSuccess) return cuda_fail(e, "cudaFree");
  return TSB_OK;
}

// ---------------------------------------------------------------------------------------------
// host staging
// ---------------------------------------------------------------------------------------------
int tsb_host_alloc(uint64_t nbytes, void** out_ptr) {
  if (!out_ptr) return fail(TSB_ERR_INVALID, "out_ptr is NULL");
  int st = tsb_init();
  if (st) return st;
  TSB_CUDA(cudaHostAlloc(out_ptr, nbytes, cudaHostAllocPortable));
  return TSB_OK;
}

int tsb_host_free(void* ptr) {
  if (!ptr) return TSB_OK;
  TSB_CUDA(cudaFreeHost(ptr));
  return TSB_OK;
}

int tsb_host_register(void* ptr, uint64_t nbytes) {
  int st = tsb_init();
  if (st) return st;
  cudaError_t e = cudaHostRegister(ptr, nbytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return TSB_OK;
  }
  if (e != cudaSuccess) return cuda_fail(e, "cudaHostRegister");
  return TSB_OK;
}

int tsb_host_unregister(void* ptr) {
  cudaError_t e = cudaHostUnregister(ptr);
  if (e == cudaErrorHostMemoryNotRegistered) {
    cudaGetLastError();
    return TSB_OK;
  }
  if (e != cudaSuccess) return cuda_fail(e, "cudaHostUnregister");
  return TSB_OK;
}

int tsb_memcpy_async(int device, void* dst, const void* src, uint64_t nbytes, int kind, void* stream) {
  int st;
  cudaStream_t s = resolve_stream(device, stream, &st);
  if (st) return st;
  cudaMemcpyKind k;
  switch (kind) {
    case TSB_H2D: k = cudaMemcpyHostToDevice; break;
    case TSB_D2H: k = cudaMemcpyDeviceToHost; break;
    case TSB_D2D: k = cudaMemcpyDeviceToDevice; break;
    default: return fail(TSB_ERR_INVALID, "tsb_memcpy_async: bad kind");
  }
  DeviceGuard guard(device);
  if (!guard.ok) return cuda_fail(guard.err, "cudaSetDevice");
  TSB_CUDA(cudaMemcpyAsync(dst, src, nbytes, k, s));
  return TSB_OK;
}

}  // extern "C"

// shutdown lives here because it tears down state owned by this file; plan.cu registers its own
// cleanup through tsb_plans_shutdown.
namespace tsb {
int plans_shutdown();
}

extern "C" int tsb_shutdown(void) {
  int st = tsb::plans_shutdown();
  int st2 = tsb_release_all();
  Global& g = G();
  std::lock_guard<std::mutex> lk(g.mu);
  for (auto& kv : g.arenas) {
    DeviceGuard guard(kv.second->device);
    cudaFree(kv.second->base);
    delete kv.second;
  }
  g.arenas.clear();
  for (int d = 0; d < g.ndev; ++d) {
    if (g.dev[d].stream_ready) {
      DeviceGuard guard(d);
      cudaStreamDestroy(g.dev[d].stream);
      g.dev[d].stream_ready = false;
    }
  }
  return st ? st : st2;
}

// Internal declarations shared by the translation units of libtstore_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "tstore_b200.h"

namespace tsb {

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define TSB_CUDA(expr)                                   \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return ::tsb::cuda_fail(_e, #expr); \
  } while (0)

// RAII: switch the calling thread to `device`, restore on scope exit (torch owns the
// thread's current device; we must not disturb it).
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  bool had_context = true;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device);  // runtime.cu
  ~DeviceGuard() {
    // restore only if this thread had a context before: on a fresh thread (the actor-server thread)
    // "restoring" device 0 would create a primary context on GPU 0 in every rank's process
    if (ok && had_context && prev >= 0 && prev != cur) cudaSetDevice(prev);
  }
  int cur = -1;
};

// ---- per-device state -----------------------------------------------------------------------
int device_sm_count(int device, int* out);
int copy_stream(int device, cudaStream_t* out);
cudaStream_t resolve_stream(int device, void* stream, int* status);

// ---- kernel-side structures -------------------------------------------------------------------
// Copy modes of a compiled rect: how one "unit" (the thing a thread moves per step) is defined.
enum : uint32_t {
  MODE_B1 = 0,   // unit = 1 byte
  MODE_B2 = 1,
  MODE_B4 = 2,
  MODE_B8 = 3,
  MODE_B16 = 4,  // unit = 16 bytes (LDG.128 / STG.128)
  // casts: unit = 8 elements (vector) or 1 element (scalar)
  MODE_F32_BF16_V8 = 8,
  MODE_F32_BF16_S = 9,
  MODE_F32_F16_V8 = 10,
  MODE_F32_F16_S = 11,
  MODE_BF16_F32_V8 = 12,
  MODE_BF16_F32_S = 13,
  MODE_F16_F32_V8 = 14,
  MODE_F16_F32_S = 15,
  MODE_BF16_F16_V8 = 16,
  MODE_BF16_F16_S = 17,
  MODE_F16_BF16_V8 = 18,
  MODE_F16_BF16_S = 19,
  MODE_F64_F32_S = 20,
  MODE_F32_F64_S = 21,
};

constexpr int kMaxOuter = TSB_MAX_DIMS;  // outer (row) dims after splitting off the run

// Device-resident compiled rect.  16-byte aligned, size multiple of 16 so that a few lanes can
// stage it into shared memory with cp.async.
struct alignas(16) DevRect {
  uint64_t src;
  uint64_t dst;
  int64_t src_stride[kMaxOuter];  // bytes, outermost first; only [0, n_outer) valid
  int64_t dst_stride[kMaxOuter];
  uint32_t ext[kMaxOuter];        // extents of the outer dims
  uint32_t n_outer;               // 0 => a single row
  uint32_t rows;                  // prod(ext[0..n_outer))
  uint32_t units_per_row;         // U
  uint32_t magic;                 // ceil(2^32 / U) (0 when U == 1)
  uint32_t wide;                  // 1: tiles are segments of one row; 0: tiles are groups of whole rows
  uint32_t split;                 // wide: tiles per row; narrow: rows per tile
  uint32_t mode;
  uint32_t src_unit_bytes;        // bytes one unit spans in src
  uint32_t dst_unit_bytes;        // bytes one unit spans in dst
  uint32_t tile_units;            // tile size of THIS rect (link rects use the TMA stage size)
  uint32_t link;                  // 1: tiles of this rect are in the link queue (TMA bulk over NVLink)
  uint32_t pad_[3];
};
static_assert(sizeof(DevRect) == 192, "DevRect layout changed: keep it a multiple of 16 bytes");

struct DevTile {
  uint32_t rect;
  uint32_t tile_in_rect;
};

// kernel specialisations (see copy_rects.cu)
enum : uint32_t { KIND_GENERIC = 0, KIND_B16 = 1, KIND_F32_BF16 = 2 };

// Two work queues per plan (see copy_rects.cu):
//   copy queue  tiles moved by the CTA's 8 copy warps with LDG.128/STG.128 (local HBM sources,
//               casts, misaligned rects, peer destinations)
//   link queue  tiles whose source is another GPU's HBM and that move 16-byte units: driven by one
//               extra "link warp" per CTA through a ring of TMA bulk copies
//               (peer global -> shared -> local global), fully asynchronous to the copy warps
struct LaunchParams {
  const DevTile* tiles;       // copy queue
  const DevRect* rects;
  const DevTile* link_tiles;  // link queue (nullptr when empty)
  uint32_t num_tiles;
  uint32_t num_link_tiles;
  uint32_t kind;
  uint32_t link_stage_bytes;  // bytes per ring stage (== tile size of link rects)
  uint32_t link_stages;       // ring depth S (stores trail loads by S-2 stages)
  uint32_t* sched;  // {copy claim, finished CTAs, link claim, pad}; nullptr = fully static striding (no link queue)
};

constexpr uint32_t kCopyThreads = 256;
constexpr uint32_t kLinkThreads = 32;
constexpr uint32_t kLinkBatch = 8;  // link tiles claimed per atomic

// defined in copy_rects.cu
int launch_copy_rects(const LaunchParams& p, uint32_t grid, cudaStream_t stream);
void count_launch();
int max_ctas_per_sm(uint32_t kind, bool with_link, uint32_t link_smem_bytes, int* out);

}  // namespace tsb

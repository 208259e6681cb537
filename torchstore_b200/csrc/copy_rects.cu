// copy_rects: the reshard kernel of libtstore_b200 (sm_100a).
//
// One persistent launch moves every rectangle of a compiled plan: local-shard slice -> (peer or
// local) gather into the destination shard, with the optional dtype cast fused into the store.
// It replaces, in the reference,
//   * asyncio.gather(rdma_buffer.read_into(...))          direct_weight_sync.py:338-340
//   * dest[dest_slices].copy_(recv[src_slices])            direct_weight_sync.py:343-350
//   * client_tensor.copy_(shm_tensor) / shm.copy_(tensor)  transport/shared_memory.py:374,475
//   * local.to(transfer_dtype) / staging.copy_(src)        direct_weight_sync.py:133,167-168
//
// Work decomposition (built on the host by plan.cu):
//   rect  -> rows x units   (unit = 16 B for byte moves, 8 elements for vector casts)
//   tile  = <= tile_units units: either a segment of one wide row or a group of whole narrow rows
//   tiles are listed in an order that round-robins over source GPUs (NVSwitch port fairness)
//   CTA b processes tiles b, b+grid, b+2*grid, ...; the next tile's rect header is staged into
//   shared memory with cp.async while the current tile is being moved, so the index math never
//   waits on HBM.
// Data path: LDG.128 (non-coherent, no L1 allocate) from local HBM or the peer-mapped NVLink
// aperture, UNROLL independent loads in flight per thread, then coalesced STG.128.

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <atomic>

#include "tsb_internal.h"

namespace tsb {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint4 ldg16(const char* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg16(char* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// ---- movers: how one unit travels from src to dst ---------------------------------------------
struct MoveB16 {
  static constexpr int kUnroll = 4;
  using Reg = uint4;
  static __device__ __forceinline__ Reg ld(const char* p) { return ldg16(p); }
  static __device__ __forceinline__ void st(char* p, const Reg& v) { stg16(p, v); }
};
template <typename T>
struct MoveSmall {
  static constexpr int kUnroll = 4;
  using Reg = T;
  static __device__ __forceinline__ Reg ld(const char* p) { return __ldg(reinterpret_cast<const T*>(p)); }
  static __device__ __forceinline__ void st(char* p, const Reg& v) { *reinterpret_cast<T*>(p) = v; }
};

// scalar conversions, all round-to-nearest-even like torch's .to()
template <typename S, typename D>
__device__ __forceinline__ D convert(S v);
template <> __device__ __forceinline__ __nv_bfloat16 convert<float, __nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half convert<float, __half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float convert<__nv_bfloat16, float>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float convert<__half, float>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ __half convert<__nv_bfloat16, __half>(__nv_bfloat16 v) { return __float2half_rn(__bfloat162float(v)); }
template <> __device__ __forceinline__ __nv_bfloat16 convert<__half, __nv_bfloat16>(__half v) { return __float2bfloat16_rn(__half2float(v)); }
template <> __device__ __forceinline__ float convert<double, float>(double v) { return __double2float_rn(v); }
template <> __device__ __forceinline__ double convert<float, double>(float v) { return static_cast<double>(v); }

template <typename S, typename D>
struct CastScalar {
  static constexpr int kUnroll = 4;
  using Reg = S;
  static __device__ __forceinline__ Reg ld(const char* p) { return *reinterpret_cast<const S*>(p); }
  static __device__ __forceinline__ void st(char* p, const Reg& v) { *reinterpret_cast<D*>(p) = convert<S, D>(v); }
};

// 8 elements per unit; 4-byte -> 2-byte (32 B in, 16 B out)
template <typename D>
struct CastNarrowV8 {
  static constexpr int kUnroll = 2;
  struct Reg { uint4 a, b; };
  static __device__ __forceinline__ Reg ld(const char* p) { return Reg{ldg16(p), ldg16(p + 16)}; }
  static __device__ __forceinline__ uint32_t pack(uint32_t lo, uint32_t hi) {
    D l = convert<float, D>(__uint_as_float(lo));
    D h = convert<float, D>(__uint_as_float(hi));
    return static_cast<uint32_t>(*reinterpret_cast<unsigned short*>(&l)) |
           (static_cast<uint32_t>(*reinterpret_cast<unsigned short*>(&h)) << 16);
  }
  static __device__ __forceinline__ void st(char* p, const Reg& v) {
    uint4 o;
    o.x = pack(v.a.x, v.a.y);
    o.y = pack(v.a.z, v.a.w);
    o.z = pack(v.b.x, v.b.y);
    o.w = pack(v.b.z, v.b.w);
    stg16(p, o);
  }
};
// 8 elements per unit; 2-byte -> 4-byte (16 B in, 32 B out)
template <typename S>
struct CastWidenV8 {
  static constexpr int kUnroll = 4;
  using Reg = uint4;
  static __device__ __forceinline__ Reg ld(const char* p) { return ldg16(p); }
  static __device__ __forceinline__ uint32_t up(uint32_t bits16) {
    unsigned short b = static_cast<unsigned short>(bits16);
    S s = *reinterpret_cast<S*>(&b);
    return __float_as_uint(convert<S, float>(s));
  }
  static __device__ __forceinline__ void st(char* p, const Reg& v) {
    uint4 lo, hi;
    lo.x = up(v.x & 0xffffu); lo.y = up(v.x >> 16); lo.z = up(v.y & 0xffffu); lo.w = up(v.y >> 16);
    hi.x = up(v.z & 0xffffu); hi.y = up(v.z >> 16); hi.z = up(v.w & 0xffffu); hi.w = up(v.w >> 16);
    stg16(p, lo);
    stg16(p + 16, hi);
  }
};
// 8 elements per unit; 2-byte -> 2-byte (bf16 <-> f16)
template <typename S, typename D>
struct CastHalfV8 {
  static constexpr int kUnroll = 4;
  using Reg = uint4;
  static __device__ __forceinline__ Reg ld(const char* p) { return ldg16(p); }
  static __device__ __forceinline__ uint32_t cv(uint32_t w) {
    unsigned short b0 = static_cast<unsigned short>(w & 0xffffu), b1 = static_cast<unsigned short>(w >> 16);
    D d0 = convert<S, D>(*reinterpret_cast<S*>(&b0));
    D d1 = convert<S, D>(*reinterpret_cast<S*>(&b1));
    return static_cast<uint32_t>(*reinterpret_cast<unsigned short*>(&d0)) |
           (static_cast<uint32_t>(*reinterpret_cast<unsigned short*>(&d1)) << 16);
  }
  static __device__ __forceinline__ void st(char* p, const Reg& v) {
    uint4 o{cv(v.x), cv(v.y), cv(v.z), cv(v.w)};
    stg16(p, o);
  }
};

// ---- index math --------------------------------------------------------------------------------
__device__ __forceinline__ void row_offsets(const DevRect& r, uint32_t row, int64_t& so, int64_t& dof) {
  const uint32_t n = r.n_outer;
  if (n <= 1) {  // n == 0: rows == 1, row == 0
    so = n ? static_cast<int64_t>(row) * r.src_stride[0] : 0;
    dof = n ? static_cast<int64_t>(row) * r.dst_stride[0] : 0;
    return;
  }
  so = 0;
  dof = 0;
  for (int d = static_cast<int>(n) - 1; d >= 1; --d) {
    const uint32_t e = r.ext[d];
    const uint32_t q = row / e;
    const uint32_t rem = row - q * e;
    so += static_cast<int64_t>(rem) * r.src_stride[d];
    dof += static_cast<int64_t>(rem) * r.dst_stride[d];
    row = q;
  }
  so += static_cast<int64_t>(row) * r.src_stride[0];
  dof += static_cast<int64_t>(row) * r.dst_stride[0];
}

// A segment of a single (wide) row: pure streaming copy, no per-unit index math.
template <typename M>
__device__ __forceinline__ void move_wide(const DevRect& r, uint32_t tile_in_rect, uint32_t tile_units) {
  constexpr int U = M::kUnroll;
  const uint32_t sub = r.src_unit_bytes, dub = r.dst_unit_bytes;
  const uint32_t row = tile_in_rect / r.split;
  const uint32_t seg = tile_in_rect - row * r.split;
  const uint32_t ustart = seg * tile_units;
  const uint32_t ucount = min(tile_units, r.units_per_row - ustart);
  int64_t so, dof;
  row_offsets(r, row, so, dof);
  const char* __restrict__ src = reinterpret_cast<const char*>(r.src) + so + static_cast<int64_t>(ustart) * sub;
  char* __restrict__ dst = reinterpret_cast<char*>(r.dst) + dof + static_cast<int64_t>(ustart) * dub;
  for (uint32_t i = threadIdx.x; i < ucount; i += kThreads * U) {
    typename M::Reg v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const uint32_t idx = i + k * kThreads;
      if (idx < ucount) v[k] = M::ld(src + static_cast<size_t>(idx) * sub);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const uint32_t idx = i + k * kThreads;
      if (idx < ucount) M::st(dst + static_cast<size_t>(idx) * dub, v[k]);
    }
  }
}

// A group of whole narrow rows of a 2-D rect (one outer dim): unit index -> (row, col) with a
// multiply-high instead of a divide.  This is the FSDP->TP hot case (1 KiB .. 28 KiB rows).
template <typename M>
__device__ __forceinline__ void move_narrow(const DevRect& r, uint32_t tile_in_rect) {
  constexpr int U = M::kUnroll;
  const uint32_t sub = r.src_unit_bytes, dub = r.dst_unit_bytes;
  const uint32_t upr = r.units_per_row;
  const uint32_t magic = r.magic;
  const uint32_t row0 = tile_in_rect * r.split;
  const uint32_t nrows = min(r.split, r.rows - row0);
  const uint32_t total = nrows * upr;
  const int64_t ss0 = r.n_outer ? r.src_stride[0] : 0;
  const int64_t ds0 = r.n_outer ? r.dst_stride[0] : 0;
  const char* __restrict__ src = reinterpret_cast<const char*>(r.src) + static_cast<int64_t>(row0) * ss0;
  char* __restrict__ dst = reinterpret_cast<char*>(r.dst) + static_cast<int64_t>(row0) * ds0;
  for (uint32_t i = threadIdx.x; i < total; i += kThreads * U) {
    typename M::Reg v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const uint32_t idx = i + k * kThreads;
      if (idx < total) {
        const uint32_t rr = magic ? __umulhi(idx, magic) : idx;
        const uint32_t c = idx - rr * upr;
        v[k] = M::ld(src + static_cast<int64_t>(rr) * ss0 + static_cast<int64_t>(c) * sub);
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const uint32_t idx = i + k * kThreads;
      if (idx < total) {
        const uint32_t rr = magic ? __umulhi(idx, magic) : idx;
        const uint32_t c = idx - rr * upr;
        M::st(dst + static_cast<int64_t>(rr) * ds0 + static_cast<int64_t>(c) * dub, v[k]);
      }
    }
  }
}

// General N-D narrow rows (3 or more collapsed dims): rare, kept simple.
template <typename M>
__device__ __noinline__ void move_narrow_nd(const DevRect& r, uint32_t tile_in_rect) {
  const uint32_t sub = r.src_unit_bytes, dub = r.dst_unit_bytes;
  const uint32_t upr = r.units_per_row;
  const uint32_t magic = r.magic;
  const uint32_t row0 = tile_in_rect * r.split;
  const uint32_t nrows = min(r.split, r.rows - row0);
  const uint32_t total = nrows * upr;
  const char* __restrict__ src = reinterpret_cast<const char*>(r.src);
  char* __restrict__ dst = reinterpret_cast<char*>(r.dst);
  for (uint32_t idx = threadIdx.x; idx < total; idx += kThreads) {
    const uint32_t rr = magic ? __umulhi(idx, magic) : idx;
    const uint32_t c = idx - rr * upr;
    int64_t so, dof;
    row_offsets(r, row0 + rr, so, dof);
    typename M::Reg v = M::ld(src + so + static_cast<int64_t>(c) * sub);
    M::st(dst + dof + static_cast<int64_t>(c) * dub, v);
  }
}

// Out-of-line wrappers for the generic (mixed-mode) kernel: one ABI call per tile keeps the
// 60-odd instantiations from sharing one register allocation.
template <typename M>
__device__ __noinline__ void move_wide_ool(const DevRect& r, uint32_t t, uint32_t tile_units) { move_wide<M>(r, t, tile_units); }
template <typename M>
__device__ __noinline__ void move_narrow_ool(const DevRect& r, uint32_t t) { move_narrow<M>(r, t); }

template <typename M, bool kInline>
__device__ __forceinline__ void move_tile(const DevRect& r, uint32_t tile_in_rect, uint32_t tile_units) {
  if (r.wide) {
    if (kInline) move_wide<M>(r, tile_in_rect, tile_units);
    else move_wide_ool<M>(r, tile_in_rect, tile_units);
  } else if (r.n_outer <= 1) {
    if (kInline) move_narrow<M>(r, tile_in_rect);
    else move_narrow_ool<M>(r, tile_in_rect);
  } else {
    move_narrow_nd<M>(r, tile_in_rect);
  }
}

// KIND selects a kernel specialisation chosen by the plan compiler:
//   KIND_GENERIC  any mix of modes (per-tile dispatch, out-of-line movers)
//   KIND_B16      every rect moves 16-byte units: the weight-sync case, fully inlined
//   KIND_F32_BF16 every rect is the vector fp32->bf16 cast (transfer_dtype=bf16), fully inlined
template <int KIND>
__device__ __forceinline__ void process_tile(const DevRect& r, uint32_t tile_in_rect, uint32_t tile_units) {
  if (KIND == KIND_B16) {
    move_tile<MoveB16, true>(r, tile_in_rect, tile_units);
    return;
  }
  if (KIND == KIND_F32_BF16) {
    move_tile<CastNarrowV8<__nv_bfloat16>, true>(r, tile_in_rect, tile_units);
    return;
  }
  switch (r.mode) {
    case MODE_B16: move_tile<MoveB16, false>(r, tile_in_rect, tile_units); break;
    case MODE_B8: move_tile<MoveSmall<uint2>, false>(r, tile_in_rect, tile_units); break;
    case MODE_B4: move_tile<MoveSmall<uint32_t>, false>(r, tile_in_rect, tile_units); break;
    case MODE_B2: move_tile<MoveSmall<uint16_t>, false>(r, tile_in_rect, tile_units); break;
    case MODE_B1: move_tile<MoveSmall<uint8_t>, false>(r, tile_in_rect, tile_units); break;
    case MODE_F32_BF16_V8: move_tile<CastNarrowV8<__nv_bfloat16>, false>(r, tile_in_rect, tile_units); break;
    case MODE_F32_BF16_S: move_tile<CastScalar<float, __nv_bfloat16>, false>(r, tile_in_rect, tile_units); break;
    case MODE_F32_F16_V8: move_tile<CastNarrowV8<__half>, false>(r, tile_in_rect, tile_units); break;
    case MODE_F32_F16_S: move_tile<CastScalar<float, __half>, false>(r, tile_in_rect, tile_units); break;
    case MODE_BF16_F32_V8: move_tile<CastWidenV8<__nv_bfloat16>, false>(r, tile_in_rect, tile_units); break;
    case MODE_BF16_F32_S: move_tile<CastScalar<__nv_bfloat16, float>, false>(r, tile_in_rect, tile_units); break;
    case MODE_F16_F32_V8: move_tile<CastWidenV8<__half>, false>(r, tile_in_rect, tile_units); break;
    case MODE_F16_F32_S: move_tile<CastScalar<__half, float>, false>(r, tile_in_rect, tile_units); break;
    case MODE_BF16_F16_V8: move_tile<CastHalfV8<__nv_bfloat16, __half>, false>(r, tile_in_rect, tile_units); break;
    case MODE_BF16_F16_S: move_tile<CastScalar<__nv_bfloat16, __half>, false>(r, tile_in_rect, tile_units); break;
    case MODE_F16_BF16_V8: move_tile<CastHalfV8<__half, __nv_bfloat16>, false>(r, tile_in_rect, tile_units); break;
    case MODE_F16_BF16_S: move_tile<CastScalar<__half, __nv_bfloat16>, false>(r, tile_in_rect, tile_units); break;
    case MODE_F64_F32_S: move_tile<CastScalar<double, float>, false>(r, tile_in_rect, tile_units); break;
    case MODE_F32_F64_S: move_tile<CastScalar<float, double>, false>(r, tile_in_rect, tile_units); break;
    default: break;
  }
}

// Stage one DevRect (global -> shared) with 16-byte cp.async issued by the first few lanes.
__device__ __forceinline__ void stage_rect(DevRect* sdst, const DevRect* gsrc) {
  constexpr uint32_t kChunks = sizeof(DevRect) / 16;
  if (threadIdx.x < kChunks) {
    const uint32_t saddr = static_cast<uint32_t>(__cvta_generic_to_shared(reinterpret_cast<char*>(sdst) + threadIdx.x * 16));
    const char* g = reinterpret_cast<const char*>(gsrc) + threadIdx.x * 16;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
  }
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Tile scheduling.  Every CTA owns the first three tiles of its column statically
// (b, b+grid, b+2*grid: their headers can be prefetched at once); after that tiles are claimed
// from a global atomic counter, so CTAs that happened to draw slow tiles (NVLink sources, short
// remainders) simply claim fewer -- no tail imbalance when tile costs are heterogeneous.  The
// claim for tile k+3 is issued before tile k's payload moves and only consumed after it, so the
// atomic's round trip is never exposed.  The last CTA to finish resets the counters, which keeps a
// sync at one kernel launch (a plan must not run concurrently with itself).
template <int KIND>
__global__ void __launch_bounds__(kThreads, 3) copy_rects_kernel(LaunchParams p) {
  __shared__ DevRect srect[2];
  __shared__ uint32_t s_claim;
  const uint32_t grid = gridDim.x;
  const uint32_t n = p.num_tiles;
  const bool dynamic = p.sched != nullptr;
  uint32_t i0 = blockIdx.x;          // index (into p.tiles) of the current tile
  uint32_t i1 = i0 + grid;           // next
  uint32_t i2 = i1 + grid;           // the one after
  if (i0 < n) {
    DevTile cur = p.tiles[i0];
    DevTile nxt = cur;
    if (i1 < n) nxt = p.tiles[i1];
    stage_rect(&srect[0], &p.rects[cur.rect]);
    stage_wait();
    __syncthreads();

    int buf = 0;
    while (true) {
      const bool has_next = i1 < n;
      DevTile nn = nxt;
      uint32_t claimed = i2 + grid;  // static fallback: keep striding
      if (has_next) {
        // header pipeline: rect of tile k+1 -> smem, tile entry k+2 -> registers, claim of tile
        // k+3 -> in flight, all while tile k's payload moves
        stage_rect(&srect[buf ^ 1], &p.rects[nxt.rect]);
        if (i2 < n) nn = p.tiles[i2];
        if (dynamic && threadIdx.x == 0) claimed = atomicAdd(p.sched, 1u) + 3u * grid;
      }
      process_tile<KIND>(srect[buf], cur.tile_in_rect, p.tile_units);
      if (!has_next) break;
      if (dynamic && threadIdx.x == 0) s_claim = claimed;
      stage_wait();
      __syncthreads();
      if (dynamic) claimed = s_claim;
      cur = nxt;
      nxt = nn;
      i1 = i2;
      i2 = claimed;
      buf ^= 1;
    }
  }
  if (dynamic && threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(p.sched + 1, 1u) == grid - 1) {  // last CTA out: re-arm for the next launch
      p.sched[0] = 0;
      p.sched[1] = 0;
      __threadfence();
    }
  }
}

std::atomic<uint64_t> g_launches{0};

}  // namespace

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int max_ctas_per_sm(uint32_t kind, int* out) {
  int n = 0;
  cudaError_t e;
  switch (kind) {
    case KIND_B16: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, copy_rects_kernel<KIND_B16>, kThreads, 0); break;
    case KIND_F32_BF16: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, copy_rects_kernel<KIND_F32_BF16>, kThreads, 0); break;
    default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, copy_rects_kernel<KIND_GENERIC>, kThreads, 0); break;
  }
  if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
  *out = n;
  return TSB_OK;
}

int launch_copy_rects(const LaunchParams& p, uint32_t grid, uint32_t block, cudaStream_t stream) {
  if (p.num_tiles == 0) return TSB_OK;
  if (block != kThreads) return fail(TSB_ERR_INVALID, "copy_rects: block must be 256");
  switch (p.kind) {
    case KIND_B16: copy_rects_kernel<KIND_B16><<<grid, kThreads, 0, stream>>>(p); break;
    case KIND_F32_BF16: copy_rects_kernel<KIND_F32_BF16><<<grid, kThreads, 0, stream>>>(p); break;
    default: copy_rects_kernel<KIND_GENERIC><<<grid, kThreads, 0, stream>>>(p); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "copy_rects_kernel launch");
  count_launch();
  return TSB_OK;
}

}  // namespace tsb

extern "C" uint64_t tsb_launch_count(void) { return tsb::g_launches.load(std::memory_order_relaxed); }

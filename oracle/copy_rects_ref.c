/*
 * copy_rects_ref.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped, never on the product path).
 *
 * Plain-C restatement of the byte movement the reference performs on the weight-sync path:
 *   - dest[dest_slices].copy_(recv[src_slices])       direct_weight_sync.py:343-350
 *   - dest_byte_view.copy_(source_bytes)              tests/test_direct_weight_sync.py:33-34
 *                                                      (MockRDMABuffer: what read_into must equal)
 *   - client_tensor.copy_(shm_tensor)                 transport/shared_memory.py:473-476
 *   - shm_tensor.copy_(tensor)                        transport/shared_memory.py:373-374
 *   - local.to(transfer_dtype) / staging.copy_(src)   direct_weight_sync.py:133,167-168
 * i.e. "copy an N-D strided rectangle element by element, converting dtype with torch's
 * round-to-nearest-even .to() semantics when the dtypes differ".
 *
 * It consumes the SAME descriptor struct as the product's C-ABI (include/tstore_b200.h
 * tsb_rect_t), with host pointers, so a test can hand identical descriptors to the CUDA kernel
 * and to this file and compare bytes.
 *
 * The float conversions restate c10's CPU algorithms (torch 2.11: c10/util/BFloat16.h
 * round_to_nearest_even, c10/util/Half.h fp16_ieee_from_fp32_value) and are pinned against
 * torch itself by tests/golden/cast_vectors.npz.  nan_mode selects the NaN encoding:
 *   0 = torch CPU (bf16 NaN -> 0x7FC0, f16 NaN -> sign|0x7E00)
 *   1 = CUDA cvt.rn (bf16/f16 NaN -> 0x7FFF, f32 NaN from f64 -> 0x7FFFFFFF), which is what
 *       torch's CUDA .to() produces and what the reference would see with GPU-resident params.
 *
 * Build: see oracle/Makefile (gcc -O2 -shared -fPIC -pthread).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "tstore_b200.h"

static uint32_t dtype_size(uint32_t dt) {
  switch (dt) {
    case TSB_U8: return 1;
    case TSB_U16: case TSB_F16: case TSB_BF16: return 2;
    case TSB_U32: case TSB_F32: return 4;
    case TSB_U64: case TSB_F64: return 8;
    default: return 0;
  }
}

/* ---- scalar conversions ------------------------------------------------------------------- */
static uint16_t f32_to_bf16(uint32_t x, int nan_mode) {
  if ((x & 0x7fffffffu) > 0x7f800000u) return nan_mode ? 0x7fffu : 0x7fc0u;
  uint32_t bias = ((x >> 16) & 1u) + 0x7fffu;
  return (uint16_t)((x + bias) >> 16);
}

static uint16_t f32_to_f16(uint32_t x, int nan_mode) {
  const uint32_t sign = (x >> 16) & 0x8000u;
  const uint32_t abs = x & 0x7fffffffu;
  if (abs > 0x7f800000u) return nan_mode ? 0x7fffu : (uint16_t)(sign | 0x7e00u);
  if (abs == 0x7f800000u) return (uint16_t)(sign | 0x7c00u);
  int32_t exp = (int32_t)(abs >> 23) - 127;
  uint32_t man = abs & 0x7fffffu;
  if (exp > 15) return (uint16_t)(sign | 0x7c00u); /* overflow -> inf */
  if (exp >= -14) {
    /* normal half: keep 10 mantissa bits, round to nearest even on the 13 dropped bits */
    uint32_t h = ((uint32_t)(exp + 15) << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++; /* may carry into the exponent / inf */
    return (uint16_t)(sign | h);
  }
  if (exp < -25) return (uint16_t)sign; /* underflow to zero (below half of the smallest denormal) */
  /* subnormal half */
  man |= 0x800000u;
  uint32_t shift = (uint32_t)(-14 - exp) + 13; /* 14..24 */
  uint32_t h = man >> shift;
  uint32_t rem = man & ((1u << shift) - 1u);
  uint32_t half = 1u << (shift - 1);
  if (rem > half || (rem == half && (h & 1u))) h++;
  return (uint16_t)(sign | h);
}

static uint32_t bf16_to_f32(uint16_t x) { return (uint32_t)x << 16; }

static uint32_t f16_to_f32(uint16_t h) {
  const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu;
  uint32_t man = h & 0x3ffu;
  if (exp == 0x1f) return sign | 0x7f800000u | (man << 13);
  if (exp == 0) {
    if (man == 0) return sign;
    /* normalise the subnormal */
    int e = -1;
    do {
      man <<= 1;
      e++;
    } while (!(man & 0x400u));
    man &= 0x3ffu;
    return sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
  }
  return sign | ((exp + 112u) << 23) | (man << 13);
}

static uint32_t f64_to_f32(uint64_t x, int nan_mode) {
  double d;
  memcpy(&d, &x, 8);
  if (d != d && nan_mode) return 0x7fffffffu;
  float f = (float)d; /* IEEE round-to-nearest-even */
  uint32_t o;
  memcpy(&o, &f, 4);
  return o;
}

static uint64_t f32_to_f64(uint32_t x) {
  float f;
  memcpy(&f, &x, 4);
  double d = (double)f;
  uint64_t o;
  memcpy(&o, &d, 8);
  return o;
}

/* convert one element; returns 0 on success, -1 for an unsupported pair */
static int convert_elem(const uint8_t* s, uint32_t sdt, uint8_t* d, uint32_t ddt, int nan_mode) {
  if (sdt == ddt) {
    memcpy(d, s, dtype_size(sdt));
    return 0;
  }
  uint16_t h;
  uint32_t w;
  uint64_t q;
  if (sdt == TSB_F32 && ddt == TSB_BF16) { memcpy(&w, s, 4); h = f32_to_bf16(w, nan_mode); memcpy(d, &h, 2); return 0; }
  if (sdt == TSB_F32 && ddt == TSB_F16) { memcpy(&w, s, 4); h = f32_to_f16(w, nan_mode); memcpy(d, &h, 2); return 0; }
  if (sdt == TSB_BF16 && ddt == TSB_F32) { memcpy(&h, s, 2); w = bf16_to_f32(h); memcpy(d, &w, 4); return 0; }
  if (sdt == TSB_F16 && ddt == TSB_F32) { memcpy(&h, s, 2); w = f16_to_f32(h); memcpy(d, &w, 4); return 0; }
  if (sdt == TSB_BF16 && ddt == TSB_F16) { memcpy(&h, s, 2); h = f32_to_f16(bf16_to_f32(h), nan_mode); memcpy(d, &h, 2); return 0; }
  if (sdt == TSB_F16 && ddt == TSB_BF16) { memcpy(&h, s, 2); h = f32_to_bf16(f16_to_f32(h), nan_mode); memcpy(d, &h, 2); return 0; }
  if (sdt == TSB_F64 && ddt == TSB_F32) { memcpy(&q, s, 8); w = f64_to_f32(q, nan_mode); memcpy(d, &w, 4); return 0; }
  if (sdt == TSB_F32 && ddt == TSB_F64) { memcpy(&w, s, 4); q = f32_to_f64(w); memcpy(d, &q, 8); return 0; }
  return -1;
}

/* Flat conversion of n elements (staging.copy_(src), direct_weight_sync.py:167-168). */
int oracle_convert(const void* src, uint32_t sdt, void* dst, uint32_t ddt, uint64_t n, int nan_mode) {
  const uint32_t es = dtype_size(sdt), ed = dtype_size(ddt);
  if (!es || !ed) return -1;
  const uint8_t* s = (const uint8_t*)src;
  uint8_t* d = (uint8_t*)dst;
  for (uint64_t i = 0; i < n; ++i)
    if (convert_elem(s + i * es, sdt, d + i * ed, ddt, nan_mode)) return -1;
  return 0;
}

/* ---- one rectangle, rows [row_begin, row_end) of its outer iteration space ---------------- */
static int copy_rect_rows(const tsb_rect_t* r, int nan_mode, uint64_t row_begin, uint64_t row_end) {
  const uint32_t es = dtype_size(r->src_dtype), ed = dtype_size(r->dst_dtype);
  if (!es || !ed || r->ndim < 1 || r->ndim > TSB_MAX_DIMS) return -1;
  const uint32_t nd = r->ndim;
  const int64_t inner = r->extent[nd - 1];
  const int inner_contig = (r->src_stride[nd - 1] == (int64_t)es) && (r->dst_stride[nd - 1] == (int64_t)ed);
  const uint8_t* sbase = (const uint8_t*)(uintptr_t)r->src;
  uint8_t* dbase = (uint8_t*)(uintptr_t)r->dst;
  for (uint64_t row = row_begin; row < row_end; ++row) {
    /* decompose row into the outer indices (all dims but the last) */
    uint64_t rem = row;
    int64_t so = 0, dof = 0;
    for (int d = (int)nd - 2; d >= 0; --d) {
      uint64_t e = (uint64_t)r->extent[d];
      uint64_t idx = rem % e;
      rem /= e;
      so += (int64_t)idx * r->src_stride[d];
      dof += (int64_t)idx * r->dst_stride[d];
    }
    const uint8_t* s = sbase + so;
    uint8_t* d = dbase + dof;
    if (r->src_dtype == r->dst_dtype && inner_contig) {
      memcpy(d, s, (size_t)inner * es);
    } else {
      for (int64_t i = 0; i < inner; ++i)
        if (convert_elem(s + i * r->src_stride[nd - 1], r->src_dtype, d + i * r->dst_stride[nd - 1], r->dst_dtype, nan_mode))
          return -1;
    }
  }
  return 0;
}

static uint64_t rect_rows(const tsb_rect_t* r) {
  uint64_t rows = 1;
  for (uint32_t d = 0; d + 1 < r->ndim; ++d) rows *= (uint64_t)r->extent[d];
  for (uint32_t d = 0; d < r->ndim; ++d)
    if (r->extent[d] == 0) return 0;
  return rows;
}

/* ---- threaded driver ------------------------------------------------------------------------- */
typedef struct {
  const tsb_rect_t* rects;
  uint64_t n;
  int nan_mode;
  int tid, nthreads;
  uint64_t total_bytes;
  int status;
  int cpu; /* >= 0: pin this worker to that core (bench.py's baseline: stable NUMA placement) */
} job_t;

static uint64_t rect_bytes(const tsb_rect_t* r) {
  uint64_t e = dtype_size(r->dst_dtype);
  for (uint32_t d = 0; d < r->ndim; ++d) e *= (uint64_t)r->extent[d];
  return e;
}

/* Each thread takes a contiguous byte-balanced share of the flattened (rect, row) space, the way
 * at::parallel_for splits a big copy_ across intra-op threads. */
static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  if (j->cpu >= 0) {
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET(j->cpu, &set);
    pthread_setaffinity_np(pthread_self(), sizeof(set), &set); /* best effort */
  }
  const uint64_t lo = j->total_bytes / (uint64_t)j->nthreads * (uint64_t)j->tid;
  const uint64_t hi = (j->tid == j->nthreads - 1) ? j->total_bytes : j->total_bytes / (uint64_t)j->nthreads * (uint64_t)(j->tid + 1);
  uint64_t pos = 0;
  for (uint64_t i = 0; i < j->n; ++i) {
    const tsb_rect_t* r = &j->rects[i];
    const uint64_t rows = rect_rows(r);
    if (!rows) continue;
    const uint64_t bytes = rect_bytes(r);
    const uint64_t per_row = bytes / rows;
    const uint64_t begin = pos, end = pos + bytes;
    pos = end;
    if (end <= lo || begin >= hi) continue;
    uint64_t rb = (lo > begin) ? (lo - begin + per_row - 1) / per_row : 0;
    uint64_t re = (hi < end) ? (hi - begin + per_row - 1) / per_row : rows;
    if (rb > rows) rb = rows;
    if (re > rows) re = rows;
    if (rb < re && copy_rect_rows(r, j->nan_mode, rb, re)) j->status = -1;
  }
  return NULL;
}

static int copy_rects_impl(const tsb_rect_t* rects, uint64_t n, int nan_mode, int nthreads, int pin);

/* Move every rect.  nthreads <= 1 runs inline.  Returns 0 or -1 (bad descriptor / unsupported cast). */
int oracle_copy_rects(const tsb_rect_t* rects, uint64_t n, int nan_mode, int nthreads) {
  return copy_rects_impl(rects, n, nan_mode, nthreads, 0);
}

/* Same, with worker t pinned to the t-th core of the caller's affinity mask: every worker owns the
 * same byte range of the same rect list on every call, so after the first touch its pages are
 * NUMA-local (what a long-lived shm segment + a pinned intra-op pool give the reference). */
int oracle_copy_rects_pinned(const tsb_rect_t* rects, uint64_t n, int nan_mode, int nthreads) {
  return copy_rects_impl(rects, n, nan_mode, nthreads, 1);
}

static int copy_rects_impl(const tsb_rect_t* rects, uint64_t n, int nan_mode, int nthreads, int pin) {
  uint64_t total = 0;
  for (uint64_t i = 0; i < n; ++i) total += rect_bytes(&rects[i]);
  if (nthreads <= 1 || total < (1u << 20)) {
    for (uint64_t i = 0; i < n; ++i) {
      uint64_t rows = rect_rows(&rects[i]);
      if (rows && copy_rect_rows(&rects[i], nan_mode, 0, rows)) return -1;
    }
    return 0;
  }
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  job_t jobs[256];
  int cpus[1024], ncpu = 0;
  if (pin) {
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
      for (int c = 0; c < CPU_SETSIZE && ncpu < 1024; ++c)
        if (CPU_ISSET(c, &allowed)) cpus[ncpu++] = c;
  }
  for (int t = 0; t < nthreads; ++t) {
    jobs[t] = (job_t){rects, n, nan_mode, t, nthreads, total, 0, ncpu ? cpus[t % ncpu] : -1};
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  int status = 0;
  for (int t = 0; t < nthreads; ++t) {
    pthread_join(th[t], NULL);
    if (jobs[t].status) status = -1;
  }
  return status;
}

/* ---- interpreter for the product's compiled plan tables ----------------------------------------
 * Replays (DevRect[192 B], DevTile{rect, tile_in_rect}) tables produced by tsb_plan_compile_host
 * on host memory, unit by unit, with the same arithmetic the CUDA kernel uses (multiply-high
 * division included).  Lets the CPU test-suite validate the plan compiler's index math against
 * oracle_copy_rects without a GPU.  Layout mirrors torchstore_b200/csrc/tsb_internal.h DevRect. */
typedef struct {
  uint64_t src, dst;
  int64_t src_stride[6];
  int64_t dst_stride[6];
  uint32_t ext[6];
  uint32_t n_outer, rows, units_per_row, magic, wide, split, mode, src_unit_bytes, dst_unit_bytes;
  uint32_t tile_units; /* tile size of this rect (link-queue rects use the TMA stage size) */
  uint32_t link;       /* 1: the rect's tiles are in the link queue (moved by TMA bulk copies) */
  uint32_t pad_[3];
} dev_rect_t;

static void mode_dtypes(uint32_t mode, uint32_t* sdt, uint32_t* ddt, uint32_t* elems) {
  *elems = 1;
  switch (mode) {
    case 0: case 1: case 2: case 3: case 4: *sdt = *ddt = TSB_U8; *elems = 1u << mode; break;
    case 8: *sdt = TSB_F32; *ddt = TSB_BF16; *elems = 8; break;
    case 9: *sdt = TSB_F32; *ddt = TSB_BF16; break;
    case 10: *sdt = TSB_F32; *ddt = TSB_F16; *elems = 8; break;
    case 11: *sdt = TSB_F32; *ddt = TSB_F16; break;
    case 12: *sdt = TSB_BF16; *ddt = TSB_F32; *elems = 8; break;
    case 13: *sdt = TSB_BF16; *ddt = TSB_F32; break;
    case 14: *sdt = TSB_F16; *ddt = TSB_F32; *elems = 8; break;
    case 15: *sdt = TSB_F16; *ddt = TSB_F32; break;
    case 16: *sdt = TSB_BF16; *ddt = TSB_F16; *elems = 8; break;
    case 17: *sdt = TSB_BF16; *ddt = TSB_F16; break;
    case 18: *sdt = TSB_F16; *ddt = TSB_BF16; *elems = 8; break;
    case 19: *sdt = TSB_F16; *ddt = TSB_BF16; break;
    case 20: *sdt = TSB_F64; *ddt = TSB_F32; break;
    case 21: *sdt = TSB_F32; *ddt = TSB_F64; break;
    default: *sdt = *ddt = 0xff; break;
  }
}

static void dev_row_offsets(const dev_rect_t* r, uint32_t row, int64_t* so, int64_t* dof) {
  *so = 0;
  *dof = 0;
  if (r->n_outer == 0) return;
  for (int d = (int)r->n_outer - 1; d >= 1; --d) {
    uint32_t e = r->ext[d], q = row / e, rem = row - q * e;
    *so += (int64_t)rem * r->src_stride[d];
    *dof += (int64_t)rem * r->dst_stride[d];
    row = q;
  }
  *so += (int64_t)row * r->src_stride[0];
  *dof += (int64_t)row * r->dst_stride[0];
}

int oracle_replay_plan(const void* rect_table, uint64_t n_rects, const uint32_t* tiles, uint64_t n_tiles,
                       uint32_t tile_units, int nan_mode) {
  const dev_rect_t* rects = (const dev_rect_t*)rect_table;
  for (uint64_t t = 0; t < n_tiles; ++t) {
    const uint32_t ri = tiles[2 * t], tir = tiles[2 * t + 1];
    if (ri >= n_rects) return -2;
    const dev_rect_t* r = &rects[ri];
    uint32_t sdt, ddt, elems;
    mode_dtypes(r->mode, &sdt, &ddt, &elems);
    if (sdt == 0xff) return -3;
    const uint32_t es = dtype_size(sdt), ed = dtype_size(ddt);
    if (es * elems != r->src_unit_bytes || ed * elems != r->dst_unit_bytes) return -4;
    if (r->tile_units) tile_units = r->tile_units; /* per-rect tile size wins over the caller's default */
    if (r->link) {
      /* what the link warp's bulk copies can express: 16-byte units, rows addressed by one stride,
       * a tile never larger than one ring stage */
      if (r->mode != 4 || (!r->wide && r->n_outer > 1)) return -10;
      uint64_t tile_bytes = r->wide ? (uint64_t)tile_units * 16 : (uint64_t)r->split * r->units_per_row * 16;
      if (tile_bytes > (uint64_t)tile_units * 16) return -11;
    }
    if (r->wide) {
      uint32_t row = tir / r->split, seg = tir - row * r->split;
      if (row >= r->rows) return -5;
      uint32_t ustart = seg * tile_units;
      if (ustart >= r->units_per_row) return -6;
      uint32_t ucount = r->units_per_row - ustart;
      if (ucount > tile_units) ucount = tile_units;
      int64_t so, dof;
      dev_row_offsets(r, row, &so, &dof);
      const uint8_t* s = (const uint8_t*)(uintptr_t)r->src + so + (int64_t)ustart * r->src_unit_bytes;
      uint8_t* d = (uint8_t*)(uintptr_t)r->dst + dof + (int64_t)ustart * r->dst_unit_bytes;
      if (oracle_convert(s, sdt, d, ddt, (uint64_t)ucount * elems, nan_mode)) return -7;
    } else {
      uint32_t row0 = tir * r->split;
      if (row0 >= r->rows) return -8;
      uint32_t nrows = r->rows - row0;
      if (nrows > r->split) nrows = r->split;
      uint32_t total = nrows * r->units_per_row;
      for (uint32_t idx = 0; idx < total; ++idx) {
        uint32_t rr = r->magic ? (uint32_t)(((uint64_t)idx * r->magic) >> 32) : idx;
        uint32_t c = idx - rr * r->units_per_row;
        if (c >= r->units_per_row) return -9; /* multiply-high division must be exact */
        int64_t so, dof;
        dev_row_offsets(r, row0 + rr, &so, &dof);
        const uint8_t* s = (const uint8_t*)(uintptr_t)r->src + so + (int64_t)c * r->src_unit_bytes;
        uint8_t* d = (uint8_t*)(uintptr_t)r->dst + dof + (int64_t)c * r->dst_unit_bytes;
        if (oracle_convert(s, sdt, d, ddt, elems, nan_mode)) return -7;
      }
    }
  }
  return 0;
}

// Plan compiler: turns a list of tsb_rect_t (N-D rectangles, element extents, byte strides) into
// the device-resident tables the copy_rects kernel walks: one DevRect per (possibly split)
// rectangle and one DevTile per <= tile_units units of work, ordered so that consecutive tiles
// pull from different source GPUs.
//
// This is the native counterpart of DirectWeightSyncDest._build_plan's cached op list
// (reference direct_weight_sync.py:221-317,334-335) and of the per-sub-request copy loop of
// the store path (transport/shared_memory.py:438-480): built once, replayed every sync.

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "tsb_internal.h"

namespace tsb {

namespace {

struct Plan {
  int device = 0;
  DevTile* d_tiles = nullptr;
  DevRect* d_rects = nullptr;
  uint32_t* d_sched = nullptr;
  uint32_t num_tiles = 0;
  uint32_t tile_units = 0;
  uint32_t kind = KIND_GENERIC;
  tsb_plan_info_t info{};
};

struct Pending {
  uint64_t plan;
  cudaEvent_t done;
  int device;
};

struct PlanRegistry {
  std::mutex mu;
  std::unordered_map<uint64_t, Plan*> plans;
  std::vector<Pending> pending;  // one-shot plans waiting for their launch to finish
  uint64_t next_id = 1;
};

PlanRegistry& R() {
  static PlanRegistry r;
  return r;
}

uint32_t dtype_size(uint32_t dt) {
  switch (dt) {
    case TSB_U8: return 1;
    case TSB_U16: case TSB_F16: case TSB_BF16: return 2;
    case TSB_U32: case TSB_F32: return 4;
    case TSB_U64: case TSB_F64: return 8;
    default: return 0;
  }
}

// (vector mode, scalar mode) for a cross-dtype pair, or false
bool cast_modes(uint32_t s, uint32_t d, uint32_t* vmode, uint32_t* smode) {
  struct Row { uint32_t s, d, v, sc; };
  static const Row table[] = {
      {TSB_F32, TSB_BF16, MODE_F32_BF16_V8, MODE_F32_BF16_S}, {TSB_F32, TSB_F16, MODE_F32_F16_V8, MODE_F32_F16_S},
      {TSB_BF16, TSB_F32, MODE_BF16_F32_V8, MODE_BF16_F32_S}, {TSB_F16, TSB_F32, MODE_F16_F32_V8, MODE_F16_F32_S},
      {TSB_BF16, TSB_F16, MODE_BF16_F16_V8, MODE_BF16_F16_S}, {TSB_F16, TSB_BF16, MODE_F16_BF16_V8, MODE_F16_BF16_S},
      {TSB_F64, TSB_F32, 0xffffffffu, MODE_F64_F32_S},        {TSB_F32, TSB_F64, 0xffffffffu, MODE_F32_F64_S},
  };
  for (const Row& r : table)
    if (r.s == s && r.d == d) {
      *vmode = r.v;
      *smode = r.sc;
      return true;
    }
  return false;
}

struct Dim {
  int64_t extent, ss, ds;
};

constexpr uint32_t kMaxRows = 1u << 30;
constexpr uint64_t kMaxUnitsPerRow = 1ull << 30;

struct Compiler {
  uint32_t tile_units;
  int plan_device;
  std::vector<DevRect> rects;
  std::vector<int32_t> rect_src_device;
  std::vector<uint32_t> rect_tiles;
  tsb_plan_info_t info{};

  // Emit one DevRect for (src, dst, outer dims, run of `units` units); splits oversize shapes.
  void emit(uint64_t src, uint64_t dst, std::vector<Dim> outer, uint64_t units, uint32_t mode,
            uint32_t sub, uint32_t dub, int32_t src_device) {
    if (units > kMaxUnitsPerRow) {
      // split the run into column chunks
      for (uint64_t u0 = 0; u0 < units; u0 += kMaxUnitsPerRow) {
        uint64_t n = std::min<uint64_t>(kMaxUnitsPerRow, units - u0);
        emit(src + u0 * sub, dst + u0 * dub, outer, n, mode, sub, dub, src_device);
      }
      return;
    }
    uint64_t rows = 1;
    for (const Dim& d : outer) rows *= static_cast<uint64_t>(d.extent);
    if (rows > kMaxRows) {
      // split the outermost dim that is > 1 in halves
      size_t k = 0;
      while (k < outer.size() && outer[k].extent == 1) ++k;
      Dim d = outer[k];
      int64_t h = d.extent / 2;
      std::vector<Dim> a = outer, b = outer;
      a[k].extent = h;
      b[k].extent = d.extent - h;
      emit(src, dst, a, units, mode, sub, dub, src_device);
      emit(src + static_cast<uint64_t>(h * d.ss), dst + static_cast<uint64_t>(h * d.ds), b, units, mode, sub, dub, src_device);
      return;
    }
    DevRect r;
    memset(&r, 0, sizeof(r));
    r.src = src;
    r.dst = dst;
    // drop unit extents created by splitting
    std::vector<Dim> o;
    for (const Dim& d : outer)
      if (d.extent != 1) o.push_back(d);
    r.n_outer = static_cast<uint32_t>(o.size());
    for (size_t i = 0; i < o.size(); ++i) {
      r.ext[i] = static_cast<uint32_t>(o[i].extent);
      r.src_stride[i] = o[i].ss;
      r.dst_stride[i] = o[i].ds;
    }
    r.rows = static_cast<uint32_t>(rows);
    r.units_per_row = static_cast<uint32_t>(units);
    r.magic = units == 1 ? 0u : static_cast<uint32_t>(((1ull << 32) + units - 1) / units);
    r.mode = mode;
    r.src_unit_bytes = sub;
    r.dst_unit_bytes = dub;
    uint64_t ntiles;
    if (units >= tile_units) {
      r.wide = 1;
      r.split = static_cast<uint32_t>((units + tile_units - 1) / tile_units);
      ntiles = rows * r.split;
    } else {
      r.wide = 0;
      r.split = std::max<uint32_t>(1u, tile_units / static_cast<uint32_t>(units));
      ntiles = (rows + r.split - 1) / r.split;
    }
    rects.push_back(r);
    rect_src_device.push_back(src_device);
    rect_tiles.push_back(static_cast<uint32_t>(ntiles));
    info.payload_bytes += rows * units * dub;
    info.src_bytes += rows * units * sub;
    if (src_device >= 0 && src_device != plan_device) info.remote_src_bytes += rows * units * sub;
    if (mode == MODE_B16 || (mode >= MODE_F32_BF16_V8 && (mode % 2) == 0 && mode <= MODE_F16_BF16_V8)) info.num_vector_rects++;
  }

  int add(const tsb_rect_t& in, uint64_t index) {
    const std::string where = "rect " + std::to_string(index) + ": ";
    if (in.ndim < 1 || in.ndim > TSB_MAX_DIMS) return fail(TSB_ERR_INVALID, where + "ndim must be in [1, " + std::to_string(TSB_MAX_DIMS) + "]");
    const uint32_t es = dtype_size(in.src_dtype), ed = dtype_size(in.dst_dtype);
    if (!es || !ed) return fail(TSB_ERR_INVALID, where + "unknown dtype");
    uint32_t vmode = 0, smode = 0;
    const bool is_cast = in.src_dtype != in.dst_dtype;
    if (is_cast && !cast_modes(in.src_dtype, in.dst_dtype, &vmode, &smode))
      return fail(TSB_ERR_UNSUPPORTED, where + "dtype cast " + std::to_string(in.src_dtype) + " -> " + std::to_string(in.dst_dtype) + " is not implemented");

    std::vector<Dim> dims;
    for (uint32_t i = 0; i < in.ndim; ++i) {
      if (in.extent[i] < 0) return fail(TSB_ERR_INVALID, where + "negative extent");
      if (in.extent[i] == 0) return TSB_OK;  // empty rect: nothing to move
      if (in.extent[i] == 1) continue;
      dims.push_back({in.extent[i], in.src_stride[i], in.dst_stride[i]});
    }
    if (in.src == 0 || in.dst == 0) return fail(TSB_ERR_INVALID, where + "NULL src/dst");

    // innermost contiguous run, in elements
    int64_t run = 1;
    if (!dims.empty() && dims.back().ss == static_cast<int64_t>(es) && dims.back().ds == static_cast<int64_t>(ed)) {
      run = dims.back().extent;
      dims.pop_back();
      while (!dims.empty() && dims.back().ss == run * static_cast<int64_t>(es) && dims.back().ds == run * static_cast<int64_t>(ed)) {
        run *= dims.back().extent;
        dims.pop_back();
      }
    }
    // merge adjacent outer dims that are jointly contiguous on both sides
    for (size_t i = dims.size(); i >= 2; --i) {
      Dim& outer = dims[i - 2];
      Dim& inner = dims[i - 1];
      if (outer.ss == inner.ss * inner.extent && outer.ds == inner.ds * inner.extent) {
        outer.extent *= inner.extent;
        outer.ss = inner.ss;
        outer.ds = inner.ds;
        dims.erase(dims.begin() + static_cast<long>(i) - 1);
      }
    }

    auto all_aligned = [&](uint64_t a) {
      if (in.src % a || in.dst % a) return false;
      for (const Dim& d : dims)
        if (static_cast<uint64_t>(d.ss < 0 ? -d.ss : d.ss) % a || static_cast<uint64_t>(d.ds < 0 ? -d.ds : d.ds) % a) return false;
      return true;
    };

    if (!is_cast) {
      const uint64_t run_bytes = static_cast<uint64_t>(run) * es;
      uint32_t vec = 16;
      while (vec > 1 && (!all_aligned(vec) || run_bytes % vec)) vec >>= 1;
      uint32_t mode = vec == 16 ? MODE_B16 : vec == 8 ? MODE_B8 : vec == 4 ? MODE_B4 : vec == 2 ? MODE_B2 : MODE_B1;
      emit(in.src, in.dst, dims, run_bytes / vec, mode, vec, vec, in.src_device);
    } else {
      if (in.src % es || in.dst % ed) return fail(TSB_ERR_INVALID, where + "src/dst not aligned to their element size");
      for (const Dim& d : dims)
        if (d.ss % static_cast<int64_t>(es) || d.ds % static_cast<int64_t>(ed)) return fail(TSB_ERR_INVALID, where + "stride not a multiple of the element size");
      if (vmode != 0xffffffffu && all_aligned(16) && run % 8 == 0) {
        emit(in.src, in.dst, dims, static_cast<uint64_t>(run) / 8, vmode, 8 * es, 8 * ed, in.src_device);
      } else {
        emit(in.src, in.dst, dims, static_cast<uint64_t>(run), smode, es, ed, in.src_device);
      }
    }
    return TSB_OK;
  }

  uint32_t kind() const {
    if (rects.empty()) return KIND_GENERIC;
    bool all_b16 = true, all_cast = true;
    for (const DevRect& r : rects) {
      all_b16 = all_b16 && r.mode == MODE_B16;
      all_cast = all_cast && r.mode == MODE_F32_BF16_V8;
    }
    return all_b16 ? KIND_B16 : all_cast ? KIND_F32_BF16 : KIND_GENERIC;
  }

  // Tile order: round-robin over source devices, starting after the plan's own device, so that
  // at any instant the resident CTAs pull from every peer (all inbound NVSwitch paths busy) and
  // all destination GPUs do not gang up on the same source.
  void order(uint32_t flags, std::vector<DevTile>* out) const {
    uint64_t total = 0;
    for (uint32_t n : rect_tiles) total += n;
    out->clear();
    out->reserve(total);
    if (flags & TSB_PLAN_NO_INTERLEAVE) {
      for (uint32_t r = 0; r < rects.size(); ++r)
        for (uint32_t t = 0; t < rect_tiles[r]; ++t) out->push_back({r, t});
      return;
    }
    std::map<int32_t, std::vector<uint32_t>> by_src;  // device -> rect ids
    for (uint32_t r = 0; r < rects.size(); ++r) by_src[rect_src_device[r]].push_back(r);
    struct Cursor {
      const std::vector<uint32_t>* rect_ids;
      size_t ri = 0;
      uint32_t ti = 0;
    };
    std::vector<Cursor> cursors;
    // rotate: first the device after ours, ..., ours last (local traffic does not need a port)
    std::vector<int32_t> keys;
    for (auto& kv : by_src) keys.push_back(kv.first);
    std::stable_sort(keys.begin(), keys.end(), [&](int32_t a, int32_t b) {
      auto rot = [&](int32_t k) -> int64_t {
        if (k < 0) return 1 << 20;
        int64_t d = static_cast<int64_t>(k) - plan_device - 1;
        if (d < 0) d += 1 << 16;
        return d;
      };
      return rot(a) < rot(b);
    });
    for (int32_t k : keys) cursors.push_back({&by_src[k]});
    // Proportional interleave: group g's j-th tile sits at fractional position (j + 0.5) / n_g of the
    // launch, so every source is drained at a constant rate for the whole kernel.  With equal groups
    // (FSDP(N)->TP(N)) this is plain round-robin; with unequal ones (mostly-local plans) it keeps the
    // small NVLink share spread out instead of bunching it at the front, where strict alternation
    // would throttle the local copies to the link rate.
    std::vector<uint64_t> group_tiles(cursors.size(), 0);
    for (size_t g = 0; g < cursors.size(); ++g)
      for (uint32_t rid : *cursors[g].rect_ids) group_tiles[g] += rect_tiles[rid];
    std::vector<uint64_t> emitted(cursors.size(), 0);
    for (uint64_t k = 0; k < total; ++k) {
      size_t best = cursors.size();
      // pick the group that is furthest behind its schedule: minimise (2*emitted+1)/(2*n)
      for (size_t g = 0; g < cursors.size(); ++g) {
        if (emitted[g] >= group_tiles[g]) continue;
        if (best == cursors.size()) { best = g; continue; }
        const unsigned __int128 lhs = static_cast<unsigned __int128>(2 * emitted[g] + 1) * group_tiles[best];
        const unsigned __int128 rhs = static_cast<unsigned __int128>(2 * emitted[best] + 1) * group_tiles[g];
        if (lhs < rhs) best = g;
      }
      Cursor& c = cursors[best];
      while (c.ti >= rect_tiles[(*c.rect_ids)[c.ri]]) {
        ++c.ri;
        c.ti = 0;
      }
      out->push_back({(*c.rect_ids)[c.ri], c.ti++});
      ++emitted[best];
    }
  }
};

uint32_t env_u32(const char* name, uint32_t dflt) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  long x = strtol(v, nullptr, 10);
  return x > 0 ? static_cast<uint32_t>(x) : dflt;
}

uint32_t default_tile_units() {
  uint32_t bytes = env_u32("TSB_TILE_BYTES", 65536);
  uint32_t units = bytes / 16;
  if (units < 64) units = 64;
  if (units > 32768) units = 32768;
  return units;
}

int compile(int device, const tsb_rect_t* rects, uint64_t n, uint32_t flags, uint32_t tile_units,
            Compiler* c, std::vector<DevTile>* tiles) {
  c->tile_units = tile_units;
  c->plan_device = device;
  for (uint64_t i = 0; i < n; ++i) {
    int st = c->add(rects[i], i);
    if (st) return st;
  }
  uint64_t total = 0;
  for (uint32_t t : c->rect_tiles) total += t;
  if (total >= (1ull << 32)) return fail(TSB_ERR_UNSUPPORTED, "plan has more than 2^32 tiles");
  c->order(flags, tiles);
  c->info.num_rects = c->rects.size();
  c->info.num_tiles = tiles->size();
  c->info.tile_bytes = tile_units * 16;
  c->info.block = 256;
  return TSB_OK;
}

void reap_pending_locked(PlanRegistry& reg, bool block) {
  for (size_t i = 0; i < reg.pending.size();) {
    Pending& p = reg.pending[i];
    cudaError_t e = block ? cudaEventSynchronize(p.done) : cudaEventQuery(p.done);
    if (e == cudaErrorNotReady) {
      cudaGetLastError();
      ++i;
      continue;
    }
    auto it = reg.plans.find(p.plan);
    if (it != reg.plans.end()) {
      Plan* pl = it->second;
      DeviceGuard guard(pl->device);
      cudaFree(pl->d_tiles);
      cudaFree(pl->d_rects);
      cudaFree(pl->d_sched);
      delete pl;
      reg.plans.erase(it);
    }
    cudaEventDestroy(p.done);
    reg.pending[i] = reg.pending.back();
    reg.pending.pop_back();
  }
}

}  // namespace

int plans_shutdown() {
  PlanRegistry& reg = R();
  std::lock_guard<std::mutex> lk(reg.mu);
  reap_pending_locked(reg, true);
  for (auto& kv : reg.plans) {
    DeviceGuard guard(kv.second->device);
    cudaFree(kv.second->d_tiles);
    cudaFree(kv.second->d_rects);
    cudaFree(kv.second->d_sched);
    delete kv.second;
  }
  reg.plans.clear();
  return TSB_OK;
}

}  // namespace tsb

using namespace tsb;

extern "C" {

int tsb_cast_supported(uint32_t src_dtype, uint32_t dst_dtype) {
  if (!dtype_size(src_dtype) || !dtype_size(dst_dtype)) return 0;
  if (src_dtype == dst_dtype) return 1;
  uint32_t v, s;
  return cast_modes(src_dtype, dst_dtype, &v, &s) ? 1 : 0;
}

// Host-only: compile rects and copy the tables out (no CUDA call).  Used by the CPU test-suite to
// check the index math of the plan compiler against the oracle without a GPU.  `out_rects` points
// at n_rect_cap records of 192 bytes (DevRect), `out_tiles` at n_tile_cap pairs of uint32.
int tsb_plan_compile_host(int device, const tsb_rect_t* rects, uint64_t n, uint32_t flags, uint32_t tile_units,
                          void* out_rects, uint64_t n_rect_cap, uint64_t* out_n_rects, void* out_tiles,
                          uint64_t n_tile_cap, uint64_t* out_n_tiles, tsb_plan_info_t* out_info) {
  if (n && !rects) return fail(TSB_ERR_INVALID, "rects is NULL");
  if (tile_units == 0) tile_units = default_tile_units();
  if (tile_units > 32768) return fail(TSB_ERR_INVALID, "tile_units must be <= 32768");
  Compiler c;
  std::vector<DevTile> tiles;
  int st = compile(device, rects, n, flags, tile_units, &c, &tiles);
  if (st) return st;
  if (out_n_rects) *out_n_rects = c.rects.size();
  if (out_n_tiles) *out_n_tiles = tiles.size();
  if (out_info) *out_info = c.info;
  if (out_rects) {
    if (c.rects.size() > n_rect_cap) return fail(TSB_ERR_NOMEM, "out_rects too small");
    memcpy(out_rects, c.rects.data(), c.rects.size() * sizeof(DevRect));
  }
  if (out_tiles) {
    if (tiles.size() > n_tile_cap) return fail(TSB_ERR_NOMEM, "out_tiles too small");
    memcpy(out_tiles, tiles.data(), tiles.size() * sizeof(DevTile));
  }
  return TSB_OK;
}

int tsb_plan_create(int device, const tsb_rect_t* rects, uint64_t n, uint32_t flags, tsb_plan_t* out) {
  if (!out || (n && !rects)) return fail(TSB_ERR_INVALID, "tsb_plan_create: NULL argument");
  int sm = 0;
  int st = device_sm_count(device, &sm);
  if (st) return st;
  PlanRegistry& reg = R();
  {
    std::lock_guard<std::mutex> lk(reg.mu);
    reap_pending_locked(reg, false);
  }
  Compiler c;
  std::vector<DevTile> tiles;
  const uint32_t tile_units = default_tile_units();
  if ((st = compile(device, rects, n, flags, tile_units, &c, &tiles))) return st;

  Plan* p = new Plan();
  p->device = device;
  p->num_tiles = static_cast<uint32_t>(tiles.size());
  p->tile_units = tile_units;
  p->info = c.info;
  p->kind = c.kind();
  uint32_t per_sm = env_u32("TSB_CTAS_PER_SM", 3);
  {
    // persistent kernel: never ask for more CTAs than can be co-resident
    DeviceGuard guard(device);
    int resident = 0;
    if (guard.ok && max_ctas_per_sm(p->kind, &resident) == TSB_OK && resident > 0 &&
        per_sm > static_cast<uint32_t>(resident))
      per_sm = static_cast<uint32_t>(resident);
  }
  uint64_t grid = static_cast<uint64_t>(sm) * per_sm;
  if (grid > tiles.size()) grid = tiles.size();
  if (grid == 0) grid = 1;
  p->info.grid = static_cast<uint32_t>(grid);
  if (!tiles.empty()) {
    DeviceGuard guard(device);
    if (!guard.ok) { delete p; return cuda_fail(guard.err, "cudaSetDevice"); }
    cudaError_t e = cudaMalloc(&p->d_tiles, tiles.size() * sizeof(DevTile));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_rects, c.rects.size() * sizeof(DevRect));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_sched, 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(p->d_sched, 0, 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpy(p->d_tiles, tiles.data(), tiles.size() * sizeof(DevTile), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_rects, c.rects.data(), c.rects.size() * sizeof(DevRect), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(p->d_tiles);
      cudaFree(p->d_rects);
      cudaFree(p->d_sched);
      delete p;
      return cuda_fail(e, "plan upload");
    }
  }
  std::lock_guard<std::mutex> lk(reg.mu);
  uint64_t id = reg.next_id++;
  reg.plans[id] = p;
  *out = id;
  return TSB_OK;
}

int tsb_plan_info(tsb_plan_t plan, tsb_plan_info_t* out) {
  if (!out) return fail(TSB_ERR_INVALID, "out is NULL");
  PlanRegistry& reg = R();
  std::lock_guard<std::mutex> lk(reg.mu);
  auto it = reg.plans.find(plan);
  if (it == reg.plans.end()) return fail(TSB_ERR_NOTFOUND, "unknown plan");
  *out = it->second->info;
  return TSB_OK;
}

int tsb_plan_run(tsb_plan_t plan, void* stream) {
  Plan* p;
  {
    PlanRegistry& reg = R();
    std::lock_guard<std::mutex> lk(reg.mu);
    auto it = reg.plans.find(plan);
    if (it == reg.plans.end()) return fail(TSB_ERR_NOTFOUND, "unknown plan");
    p = it->second;
  }
  if (p->num_tiles == 0) return TSB_OK;
  int st;
  cudaStream_t s = resolve_stream(p->device, stream, &st);
  if (st) return st;
  DeviceGuard guard(p->device);
  if (!guard.ok) return cuda_fail(guard.err, "cudaSetDevice");
  uint32_t kind = p->kind;
  if (getenv("TSB_FORCE_GENERIC")) kind = KIND_GENERIC;
  uint32_t* sched = getenv("TSB_STATIC_SCHED") ? nullptr : p->d_sched;
  LaunchParams lp{p->d_tiles, p->d_rects, p->num_tiles, p->tile_units, kind, sched};
  return launch_copy_rects(lp, p->info.grid, p->info.block, s);
}

int tsb_plan_destroy(tsb_plan_t plan) {
  PlanRegistry& reg = R();
  std::lock_guard<std::mutex> lk(reg.mu);
  auto it = reg.plans.find(plan);
  if (it == reg.plans.end()) return fail(TSB_ERR_NOTFOUND, "unknown plan");
  Plan* p = it->second;
  reg.plans.erase(it);
  DeviceGuard guard(p->device);
  // cudaFree synchronises with outstanding work that uses the buffers
  cudaFree(p->d_tiles);
  cudaFree(p->d_rects);
  cudaFree(p->d_sched);
  delete p;
  return TSB_OK;
}

int tsb_copy_rects(int device, const tsb_rect_t* rects, uint64_t n, uint32_t flags, void* stream) {
  tsb_plan_t plan;
  int st = tsb_plan_create(device, rects, n, flags, &plan);
  if (st) return st;
  st = tsb_plan_run(plan, stream);
  if (st) {
    tsb_plan_destroy(plan);
    return st;
  }
  // free the tables once the launch has drained, without blocking the caller
  cudaStream_t s = resolve_stream(device, stream, &st);
  if (st) return st;
  DeviceGuard guard(device);
  cudaEvent_t ev;
  cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(ev, s);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return tsb_plan_destroy(plan);
  }
  PlanRegistry& reg = R();
  std::lock_guard<std::mutex> lk(reg.mu);
  reg.pending.push_back({plan, ev, device});
  return TSB_OK;
}

}  // extern "C"

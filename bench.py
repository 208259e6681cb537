#!/usr/bin/env python
"""bench.py -- weight-sync GB/s of a Llama-3-8B bf16 state_dict resharded FSDP(N) -> TP(N).

    python bench.py --gpus N --steps K --warmup W            (torchrun launches it for N > 1)
    python bench.py --impl reference ...                      (the reference's CPU path, oracle port)

A "step" is ONE weight sync of the whole state_dict: every rank holds its FSDP Shard(0) source
shard and pulls its TP shard from all sources through the public API
(``ts.put_state_dict(direct_rdma=True)`` / ``ts.get_state_dict(direct_rdma=True)``), which runs one
persistent copy_rects launch per destination GPU.

  value   state_dict bytes * K / device time of K steps (CUDA events, max over ranks), inputs
          resident in HBM when the timed region starts
  e2e     the same through the same API with HOST inputs: every step copies the rank's new source
          weights from pinned host memory (H2D), refreshes/fences the source, pulls, and reads a
          result sample back (D2H); wall clock, barrier + synchronize on both sides, max over ranks
  roofline      algorithmic bytes per copy_rects launch / its average duration vs the measured peak
  cpu_baseline  the oracle's port of the reference's shm path (2 memcpy per byte) on the host cores
"""

from __future__ import annotations

import argparse
import asyncio
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import workloads  # noqa: E402

METRIC = "weight-sync GB/s (state_dict bytes / wall s)"
NVLINK_MEASURED_GBPS = 770.0  # peer copy per direction per GPU, B200_PROFILING.md
NVLINK_NOMINAL_GBPS = 900.0


# =================================================================================================
# helpers
# =================================================================================================
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML every ~5 ms; falls back to
    `nvidia-smi -lms` when pynvml is unavailable)."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.samples: list[tuple[float, int, float]] = []
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(self.device_index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None
        self._thread = threading.Thread(target=self._loop_nvml if self._nvml else self._loop_smi, daemon=True)
        self._thread.start()

    def _loop_nvml(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                clk = float(nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._handle))
                except Exception:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle))
                try:
                    power = nv.nvmlDeviceGetPowerUsage(self._handle) / 1000.0
                except Exception:
                    power = 0.0
                self.samples.append((clk, reasons, power))
            except Exception:
                pass
            time.sleep(0.004)

    def _loop_smi(self):
        fields = "clocks.sm,clocks.max.sm,power.draw"
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={fields}", "--format=csv,noheader,nounits", "-i",
                                      str(self.device_index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                self.samples.append((float(parts[0]), 0, float(parts[2])))
                self.sm_max = float(parts[1])
            except Exception:
                pass

    def stop(self) -> dict:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=6)
        clocks = [c for c, _, _ in self.samples]
        bits = 0
        for _, r, _ in self.samples:
            bits |= r
        reasons = sorted(name for name, mask in self.REASONS.items() if bits & mask)
        return {"sm_mhz": statistics.median(clocks) if clocks else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(clocks), "power_w_max": max((p for _, _, p in self.samples), default=None),
                "source": "nvml" if self._nvml else "nvidia-smi"}


_REAL_STDOUT_FD = None


def capture_stdout():
    """Route everything libraries print to stdout (e.g. NCCL's version banner) to stderr, so that
    stdout carries exactly ONE line: the JSON result written by emit()."""
    global _REAL_STDOUT_FD
    if _REAL_STDOUT_FD is None:
        sys.stdout.flush()
        _REAL_STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT_FD, line)


# =================================================================================================
# reference arm / cpu baseline: the oracle's port of the reference's shm path on the host
# =================================================================================================
def cpu_reference_run(n_ranks: int, steps: int, warmup: int, budget_s: float = 20.0, layers: int | None = None):
    """Time put_state_dict + get_state_dict of the reference's SharedMemory path, restated:
    put = every source shard copied into its shm segment (transport/shared_memory.py:373-374);
    get = every stored rectangle that intersects the wanted slice copied segment -> destination
    (shared_memory.py:473-476, client.py:284-314).  Byte movement by oracle/copy_rects_ref.c with
    all host threads; Python/RPC/pickle overhead of the real reference is NOT included (favours it).

    The workload is a bounded SAMPLE of the Llama-3-8B FSDP(n)->TP(n) sync: the first L transformer
    layers, all n source and n destination ranks emulated in one address space, L sized to the budget."""
    import numpy as np

    from oracle import c_oracle
    from torchstore_b200 import _native  # struct definition only (no GPU call)

    cores = os.cpu_count() or 1
    if layers is None:
        layers = 4 if n_ranks > 1 else 2
    layout = workloads.llama_layout(workloads.LLAMA3_8B, n_layers=layers, with_embeddings=False)
    sample_bytes = workloads.state_dict_bytes(layout)
    rng = np.random.default_rng(0)

    # source shards (rank-major), shm segments, destinations
    src, seg, dst = {}, {}, {}
    for name, (shape, tp) in layout.items():
        for r in range(n_ranks):
            off, shp = workloads.shard_box(shape, n_ranks, r, ("S", 0)) if n_ranks > 1 else ((0,) * len(shape), shape)
            a = rng.integers(0, 65536, size=shp, dtype=np.uint16)
            src[(name, r)] = a
            seg[(name, r)] = np.zeros(shp, dtype=np.uint16)
        for r in range(n_ranks):
            off, shp = workloads.shard_box(shape, n_ranks, r, tp) if n_ranks > 1 else ((0,) * len(shape), shape)
            dst[(name, r)] = np.zeros(shp, dtype=np.uint16)

    def window(arr, index):
        ptr = arr.ctypes.data
        strides = arr.strides
        shape = []
        for (a, b), st in zip(index, strides):
            ptr += a * st
            shape.append(b - a)
        return ptr, shape, strides

    def rect(r, s_ptr, d_ptr, shape, s_strides, d_strides):
        r.src, r.dst, r.ndim = s_ptr, d_ptr, max(1, len(shape))
        for i in range(_native.TSB_MAX_DIMS):
            r.extent[i], r.src_stride[i], r.dst_stride[i] = 1, 0, 0
        for i, (e, ss, ds) in enumerate(zip(shape, s_strides, d_strides)):
            r.extent[i], r.src_stride[i], r.dst_stride[i] = e, ss, ds
        r.src_dtype = r.dst_dtype = _native.TSB_U16
        r.src_device = -1

    put_list = [(src[k], seg[k]) for k in src]
    put_rects = _native.make_rect_array(len(put_list))
    for i, (a, b) in enumerate(put_list):
        full = tuple((0, e) for e in a.shape)
        sp, shape, ss = window(a, full)
        dp, _, ds = window(b, full)
        rect(put_rects[i], sp, dp, shape, ss, ds)
    get_specs = []
    for drank in range(n_ranks):
        for name, srank, s_idx, d_idx, _exact in workloads.fsdp_to_tp_rects(layout, n_ranks, drank):
            get_specs.append((seg[(name, srank)], s_idx, dst[(name, drank)], d_idx))
    get_rects = _native.make_rect_array(len(get_specs))
    for i, (a, s_idx, b, d_idx) in enumerate(get_specs):
        sp, shape, ss = window(a, s_idx)
        dp, _, ds = window(b, d_idx)
        rect(get_rects[i], sp, dp, shape, ss, ds)

    def one_sync():
        c_oracle.copy_rects(put_rects, len(put_list), 0, cores)
        c_oracle.copy_rects(get_rects, len(get_specs), 0, cores)

    for _ in range(max(1, warmup)):
        one_sync()
    # the baseline must be doing the real work: every destination rectangle equals its source slice
    for drank in range(n_ranks):
        for name, srank, s_idx, d_idx, _exact in workloads.fsdp_to_tp_rects(layout, n_ranks, drank):
            got = dst[(name, drank)][tuple(slice(a, b) for a, b in d_idx)]
            want = src[(name, srank)][tuple(slice(a, b) for a, b in s_idx)]
            if not np.array_equal(got, want):
                raise SystemExit(f"cpu reference port produced wrong bytes for {name} ({srank}->{drank})")
    t0 = time.perf_counter()
    done = 0
    while done < steps and (time.perf_counter() - t0 < budget_s or done == 0):
        one_sync()
        done += 1
    elapsed = time.perf_counter() - t0
    gbps = sample_bytes * done / elapsed / 1e9
    return {
        "value": gbps,
        "ms_per_step": elapsed / done * 1e3,
        "steps": done,
        "cores": cores,
        "kind": "port",
        "sample": (f"first {layers} transformer layers of llama3-8b ({sample_bytes} B of bf16, {len(put_list)} put copies + "
                   f"{len(get_specs)} get rectangles) FSDP({n_ranks})->TP({n_ranks}), all ranks emulated in one process, "
                   f"{done} syncs; byte movement = oracle/copy_rects_ref.c on {cores} threads; no Python/RPC/pickle cost"),
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_reference_run(args.gpus, args.steps, args.warmup)
    layout = workloads.llama_layout()
    emit({
        "impl": "reference",
        "metric": METRIC,
        "value": res["value"],
        "unit": "GB/s",
        "n_gpus": args.gpus,
        "steps": res["steps"],
        "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"],
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"llama3-8b bf16 state_dict FSDP({args.gpus})->TP({args.gpus}) put_state_dict+get_state_dict, "
                               "reference SharedMemory path restated on host memory (bounded sample)",
                   "state_dict_bytes": workloads.state_dict_bytes(layout), "tensors": len(layout),
                   "parallelism": f"fsdp{args.gpus}->tp{args.gpus}"},
        "cpu_baseline": {"value": res["value"], "unit": "GB/s", "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# =================================================================================================
# our arm
# =================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist

    import torchstore_b200 as ts
    from torchstore_b200 import _native
    from torchstore_b200.planner import StridedMem, build_rects

    import faulthandler

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # a hung collective / RPC must leave evidence: dump every thread's stack to stderr after the
    # watchdog period, repeatedly
    faulthandler.enable()
    faulthandler.dump_traceback_later(int(os.environ.get("TSB_BENCH_WATCHDOG_S", "90")), repeat=True, file=sys.stderr)
    # ... and must not hang the caller forever: hard deadline for the whole run
    deadline = float(os.environ.get("TSB_BENCH_DEADLINE_S", "420"))

    def _deadline():
        print(f"[bench r{os.environ.get('RANK', '0')}] deadline of {deadline:.0f}s exceeded; aborting", file=sys.stderr, flush=True)
        faulthandler.dump_traceback(file=sys.stderr)
        os._exit(3)

    _timer = threading.Timer(deadline, _deadline)
    _timer.daemon = True
    _timer.start()
    t_start = time.perf_counter()

    def phase(msg):
        if os.environ.get("TSB_BENCH_VERBOSE", "1") == "1":
            print(f"[bench r{rank} +{time.perf_counter() - t_start:7.2f}s] {msg}", file=sys.stderr, flush=True)

    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29517")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("LOCAL_RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("LOCAL_WORLD_SIZE", str(world))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if not os.environ.get("TSB_KEEP_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = "WARN"  # the version banner goes to stdout; keep stdout to ONE JSON line
    phase("init_process_group")
    dist.init_process_group("nccl", device_id=dev)
    _native.init()
    n = world
    phase("generating weights")

    layout = workloads.llama_layout(workloads.LLAMA3_8B)
    sd_bytes = workloads.state_dict_bytes(layout)

    # ---- synthetic weights: identical full tensors on every rank, sliced into this rank's shards ----
    src_numel = dst_numel = 0
    boxes = {}
    for name, (shape, tp) in layout.items():
        s_off, s_shape = workloads.shard_box(shape, n, rank, ("S", 0)) if n > 1 else ((0,) * len(shape), tuple(shape))
        d_off, d_shape = workloads.shard_box(shape, n, rank, tp) if n > 1 else ((0,) * len(shape), tuple(shape))
        boxes[name] = (s_off, s_shape, d_off, d_shape)
        src_numel += math.prod(s_shape) + 64  # 128-byte alignment slack per tensor
        dst_numel += math.prod(d_shape) + 64
    # flat parameter buffers (one allocation per side, like FSDP flat params): one H2D per step
    src_flat = torch.empty(src_numel, dtype=torch.bfloat16, device=dev)
    dst_flat = torch.zeros(dst_numel, dtype=torch.bfloat16, device=dev)
    src_local, dst_local, expect = {}, {}, {}
    so = do = 0
    gen = torch.Generator(device=dev)
    for idx, (name, (shape, tp)) in enumerate(layout.items()):
        s_off, s_shape, d_off, d_shape = boxes[name]
        gen.manual_seed(1234 + idx)
        full = (torch.randn(shape, generator=gen, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)
        ns, nd = math.prod(s_shape), math.prod(d_shape)
        src_local[name] = src_flat[so:so + ns].view(s_shape)
        dst_local[name] = dst_flat[do:do + nd].view(d_shape)
        src_local[name].copy_(full[tuple(slice(o, o + e) for o, e in zip(s_off, s_shape))])
        want = full[tuple(slice(o, o + e) for o, e in zip(d_off, d_shape))]
        expect[name] = int(want.contiguous().view(torch.int16).to(torch.int64).sum().item())
        so += (ns + 63) // 64 * 64
        do += (nd + 63) // 64 * 64
        del full, want
    torch.cuda.synchronize()
    phase("weights ready")

    # ---- state dicts as the API sees them: DTensors on a 1-D mesh (plain tensors at N == 1) -------
    if n > 1:
        from torch.distributed.device_mesh import init_device_mesh
        from torch.distributed.tensor import DTensor, Replicate, Shard

        mesh = init_device_mesh("cuda", (n,))

        def P(p):
            return Shard(p[1]) if p[0] == "S" else Replicate()

        src_sd = {k: DTensor.from_local(v, mesh, (Shard(0),), run_check=False, shape=torch.Size(layout[k][0]),
                                        stride=torch.empty(layout[k][0], device="meta").stride())
                  for k, v in src_local.items()}
        dst_sd = {k: DTensor.from_local(v, mesh, (P(layout[k][1]),), run_check=False, shape=torch.Size(layout[k][0]),
                                        stride=torch.empty(layout[k][0], device="meta").stride())
                  for k, v in dst_local.items()}
    else:
        src_sd, dst_sd = dict(src_local), dict(dst_local)

    def barrier():
        dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def allmax(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    KEY = "policy"

    async def main():
        if n > 1:
            # reuse the process group's c10d store as the rendezvous (no second TCP server)
            await ts.initialize_spmd(ts.LocalRankStrategy(), rendezvous=dist.distributed_c10d._get_default_store())
        else:
            await ts.initialize()
        phase("store initialized")
        # first sync: registers handles, exchanges them through the store, builds + caches the plan
        t_first = time.perf_counter()
        await ts.put_state_dict(src_sd, KEY, direct_rdma=True)
        phase("handles published")
        barrier()
        await ts.get_state_dict(KEY, user_state_dict=dst_sd, direct_rdma=True)
        phase("first pull done")
        barrier()
        first_ms = (time.perf_counter() - t_first) * 1e3

        def verify(tag):
            bad = [k for k, v in dst_local.items()
                   if int(v.contiguous().view(torch.int16).to(torch.int64).sum().item()) != expect[k]]
            if bad:
                raise SystemExit(f"[rank {rank}] PARITY FAILURE after {tag}: {bad[:5]} ({len(bad)} tensors)")

        verify("first sync")
        phase("first sync verified")
        from torchstore_b200.state_dict_utils import _get_rdma_cache

        cl = await ts.client()
        dest_sync = _get_rdma_cache(cl).dest
        info = dest_sync.plan_info()[local_rank]

        async def step():
            await ts.put_state_dict(None, KEY, direct_rdma=True)  # refresh + fence on the source side
            await ts.get_state_dict(KEY, user_state_dict=dst_sd, direct_rdma=True)

        # ---- value: HBM-resident inputs, device-timed ------------------------------------------------
        for _ in range(args.warmup):
            await step()
            barrier()
        dst_flat.zero_()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = _native.launch_count()
        kernel_ms = []
        copy_stream = _native.copy_stream(local_rank)
        ev0 = _native.Event(local_rank, timing=True)
        ev1 = _native.Event(local_rank, timing=True)
        barrier()
        if args.profiler_range:
            torch.cuda.cudart().cudaProfilerStart()
        ev0.record(copy_stream)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            # K syncs back to back on every rank; the bracket is barrier + synchronize on both sides.
            # (The cross-rank "sources are ready" dependency of a real sync is part of the e2e leg.)
            await step()
            kernel_ms.append(dest_sync.last_pull_ms[local_rank])
        ev1.record(copy_stream)
        ev1.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        if args.profiler_range:
            torch.cuda.cudart().cudaProfilerStop()
        dev_ms = ev0.elapsed_ms(ev1)
        phase(f"timed steps done: {dev_ms / args.steps:.3f} ms/step on this rank")
        barrier()
        launches = _native.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        verify("timed steps")
        dev_ms_max = allmax(dev_ms)
        wall_ms_max = allmax(wall_ms)
        kern_avg = sum(kernel_ms) / len(kernel_ms)
        kern_avg_max = allmax(kern_avg)
        total_launches = int(allsum(launches))

        e2e_ms, e2e_launches, src_bytes, d2h_bytes = None, 0, 0, 0
        if not args.no_e2e:
            # ---- e2e: host inputs, public API, wall clock ------------------------------------------------
            src_bytes = src_flat.numel() * 2
            host_src = torch.empty(src_flat.numel(), dtype=torch.bfloat16).pin_memory()
            host_src.copy_(src_flat.cpu())
            # result sample: first row (<= 4096 elements) of every destination tensor, gathered by one
            # copy_rects launch into a contiguous buffer and read back with one D2H copy
            rows = []
            for k, v in dst_local.items():
                flat = v.reshape(-1) if v.dim() == 1 else v[0]
                rows.append(flat[: min(4096, flat.numel())])
            sample_dev = torch.zeros(sum(r.numel() for r in rows), dtype=torch.bfloat16, device=dev)
            pairs, off = [], 0
            for r in rows:
                pairs.append((StridedMem.from_tensor(r), StridedMem.from_tensor(sample_dev[off:off + r.numel()])))
                off += r.numel()
            rects, nr = build_rects(pairs)
            sample_plan = _native.plan_create(local_rank, rects, nr)
            sample_host = torch.empty(sample_dev.numel(), dtype=torch.bfloat16).pin_memory()
            d2h_bytes = sample_dev.numel() * 2

            async def e2e_step():
                # (1) new weights arrive from the host
                _native.memcpy_async(local_rank, src_flat.data_ptr(), host_src.data_ptr(), src_bytes, _native.TSB_H2D, copy_stream)
                _native.stream_sync(local_rank, copy_stream)
                # (2) publish + pull through the public API
                await ts.put_state_dict(None, KEY, direct_rdma=True)
                if n > 1:
                    dist.barrier(device_ids=[local_rank])  # every source refreshed before anyone pulls
                await ts.get_state_dict(KEY, user_state_dict=dst_sd, direct_rdma=True)
                # (3) read the result sample back
                _native.plan_run(sample_plan, copy_stream)
                _native.memcpy_async(local_rank, sample_host.data_ptr(), sample_dev.data_ptr(), d2h_bytes, _native.TSB_D2H, copy_stream)
                _native.stream_sync(local_rank, copy_stream)

            for _ in range(max(1, min(args.warmup, 2))):
                await e2e_step()
            barrier()
            e_launch0 = _native.launch_count()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                await e2e_step()
            barrier()
            phase("e2e steps done")
            e2e_ms = allmax((time.perf_counter() - t0) * 1e3)
            e2e_launches = int(allsum(_native.launch_count() - e_launch0))
            assert torch.equal(sample_host.to(dev), sample_dev)
            verify("e2e steps")
            _native.plan_destroy(sample_plan)
        return dict(first_ms=first_ms, dev_ms=dev_ms_max, wall_ms=wall_ms_max, kern_avg=kern_avg_max, info=info,
                    launches=total_launches, clocks=clocks, e2e_ms=e2e_ms, e2e_launches=e2e_launches,
                    h2d=int(allsum(src_bytes)), d2h=int(allsum(d2h_bytes)))

    r = asyncio.run(main())
    faulthandler.cancel_dump_traceback_later()
    _timer.cancel()

    cpu = None
    if rank == 0 and n == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_run(1, steps=1000, warmup=1, budget_s=12.0)

    if rank == 0:
        peaks, peak_src = load_peaks()
        steps = args.steps
        value = sd_bytes * steps / (r["dev_ms"] / 1e3) / 1e9
        e2e = sd_bytes * steps / (r["e2e_ms"] / 1e3) / 1e9 if r["e2e_ms"] else None
        info = r["info"]
        # per-launch algorithmic bytes on THIS rank: bytes read + bytes written (SURVEY section 8d:
        # HBM-bound points count 2 x payload); remote reads do not touch local HBM
        local_read = info["src_bytes"] - info["remote_src_bytes"]
        algo_hbm = local_read + info["payload_bytes"]
        kern_s = r["kern_avg"] / 1e3
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(str(n))
            except Exception:
                traffic = None
        if n == 1:
            roofline = {"bound": "hbm", "achieved": algo_hbm / kern_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": algo_hbm / kern_s / 1e9 / peaks["hbm_gbs"], "traffic": traffic,
                        "peak_source": peak_src, "kernel": "copy_rects_kernel<KIND_B16>",
                        "algorithmic_bytes_per_launch": algo_hbm, "kernel_ms_avg": r["kern_avg"]}
        else:
            nvl = info["remote_src_bytes"] / kern_s / 1e9
            roofline = {"bound": "nvlink", "achieved": nvl, "peak": NVLINK_MEASURED_GBPS, "unit": "GB/s",
                        "frac": nvl / NVLINK_MEASURED_GBPS, "traffic": traffic,
                        "peak_source": "measured peer copy 770 GB/s/dir (B200_PROFILING.md); nominal 900",
                        "kernel": "copy_rects_kernel<KIND_B16>",
                        "algorithmic_bytes_per_launch": info["remote_src_bytes"], "kernel_ms_avg": r["kern_avg"],
                        "hbm": {"achieved": algo_hbm / kern_s / 1e9, "peak": peaks["hbm_gbs"]}}
        out = {
            "metric": METRIC,
            "value": value,
            "unit": "GB/s",
            "n_gpus": n,
            "steps": steps,
            "warmup": args.warmup,
            "ms_per_step": r["dev_ms"] / steps,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "bf16",
            "data": "synthetic",
            "config": {
                "workload": f"llama3-8b bf16 state_dict ({len(layout)} tensors) FSDP({n}) Shard(0) -> TP({n}) "
                            "direct_weight_sync via ts.put_state_dict/get_state_dict(direct_rdma=True), colocated ranks",
                "state_dict_bytes": sd_bytes,
                "parallelism": f"fsdp{n}->tp{n}",
                "rects_per_dest_rank": info["num_rects"],
                "tiles_per_launch": info["num_tiles"],
                "l2": "inputs larger than L2 (>= 2 GB moved per GPU per step; no flush)",
                "first_sync_ms_incl_registration_and_planning": r["first_ms"],
                "wall_ms_per_step": r["wall_ms"] / steps,
            },
            "clocks": r["clocks"],
            "e2e": {"value": e2e, "unit": "GB/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                    "ms_per_step": (r["e2e_ms"] / steps) if r["e2e_ms"] else None, "gpu_launches": r["e2e_launches"]},
            "gpu_launches": r["launches"],
            "roofline": roofline,
            "nvlink_fraction": (info["remote_src_bytes"] / (r["dev_ms"] / steps / 1e3) / 1e9 / NVLINK_NOMINAL_GBPS) if n > 1 else 0.0,
            "cpu_baseline": None if cpu is None else {"value": cpu["value"], "unit": "GB/s", "cores": cpu["cores"],
                                                      "kind": cpu["kind"], "sample": cpu["sample"]},
        }
        emit(out)

    async def fin():
        await ts.shutdown()

    try:
        asyncio.run(fin())
    except Exception as e:  # noqa: BLE001
        print(f"[rank {rank}] shutdown: {e}", file=sys.stderr)
    dist.barrier(device_ids=[local_rank])
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed steps (for ncu --profile-from-start off)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-input leg (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    capture_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

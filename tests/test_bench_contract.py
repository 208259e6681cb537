"""The JSON line bench.py prints is a contract with the driver: every required key must be there for the
HBM-bound (N=1) and the NVLink-bound (N>1) form, with traffic figures coming from the committed ncu table."""

import importlib
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    return importlib.import_module("bench")


INFO = {"num_rects": 1194, "num_tiles": 22665, "payload_bytes": 2008031232, "src_bytes": 2008031232,
        "remote_src_bytes": 528948224, "num_link_tiles": 143815, "link_bytes": 528948224, "grid": 444, "block": 288,
        "tile_bytes": 65536, "num_vector_rects": 1194, "link_tile_bytes": 4096, "link_stages": 6}
REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"]


def test_roofline_blocks_have_the_contract_keys_and_live_traffic():
    b = _bench()
    hbm = b.roofline_for(False, dict(INFO, remote_src_bytes=0), 4.7, 1)
    nvl = b.roofline_for(True, INFO, 0.81, 8)
    for r in (hbm, nvl):
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in r
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert hbm["bound"] == "hbm" and nvl["bound"] == "nvlink"
    table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert sorted(table) == ["1", "2", "4", "8"]
    assert hbm["traffic"] == table["1"]["dram_bytes"] and nvl["traffic"] == table["8"]["dram_bytes"]
    assert nvl["nvlink_traffic"] == table["8"]["nvlrx_bytes"]
    # the capture of the shipped build moves exactly the algorithmic bytes over the link, and no more than them through DRAM
    assert table["8"]["nvlrx_user_bytes"] == INFO["remote_src_bytes"]
    algo_hbm_n1 = 2 * 16060522496
    assert 0.99 < table["1"]["dram_bytes"] / algo_hbm_n1 < 1.01
    assert nvl["frac_of_bidirectional_peak"] > nvl["frac"]  # the two-way link rate is the tighter bound
    assert nvl["hbm"]["algorithmic_bytes_per_launch_incl_serving_peers"] == 2 * INFO["payload_bytes"]


def test_base_line_carries_every_contract_key():
    b = _bench()
    ctx = types.SimpleNamespace(world=8, args=types.SimpleNamespace(steps=30, warmup=3, config="4"), numa={"bound": True})
    timed = {"clocks": {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []}, "launches": 240}
    line = b.base_line(ctx, 18000.0, 0.85, "workload", {"state_dict_bytes": 16060522496}, timed,
                       b.roofline_for(True, INFO, 0.81, 8),
                       {"value": 414.0, "unit": "GB/s", "h2d_bytes_per_step": 16060522496, "d2h_bytes_per_step": 1}, None, 0.7)
    for k in REQUIRED:
        assert k in line, k
    assert line["config"]["workload"] == "workload" and "model" not in line["config"]
    assert line["higher_is_better"] is True and line["n_gpus"] == 8 and line["gpu_launches"] == 240
    json.dumps(line)  # serialisable

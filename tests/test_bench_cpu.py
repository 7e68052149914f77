"""bench.py's host side, checked without a GPU: the config table is BASELINE.json's, the e2e destination is what image.cpp allocates,
the reference arm (`--impl reference`, the one leg the driver runs on the host cores) prints a line with the contract's keys, and
profiles/kernel_facts.json has what the roofline block reads."""
import json
import pathlib
import subprocess
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402


def test_config_table_is_baselines():
    c = bench.CONFIGS
    assert sorted(c) == ["c1", "c2", "c3", "c4", "c5"]
    assert (c["c1"].width, c["c1"].height, c["c1"].spp, c["c1"].depth, c["c1"].mode) == (800, 600, 30, 10, "mg")        # main.cpp:153, scene.hpp:10-11
    assert (c["c2"].width, c["c2"].height, c["c2"].spp, c["c2"].depth, c["c2"].mode) == (1920, 1080, 64, 50, "sm")
    assert (c["c3"].width, c["c3"].height, c["c3"].spp) == (1920, 1080, 256) and c["c3"].scene == "rtiow"
    assert (c["c4"].width, c["c4"].height, c["c4"].spp) == (3840, 2160, 64) and c["c4"].scene == "grid"
    assert (c["c5"].width, c["c5"].height, c["c5"].spp) == (3840, 2160, 4096) and c["c5"].scene == "rtiow"
    assert len(bench.load_scene(c["c3"]).spheres) == 484 and len(bench.load_scene(c["c1"]).spheres) == 3
    for cfg in c.values():
        assert cfg.key in cfg.label.lower() and str(cfg.spp) + "spp" in cfg.label
        assert cfg.cpu_spp <= cfg.spp and cfg.cpu_renderer in ("mg_ray_tracer", "sm_ray_tracer")


def test_e2e_destination_is_pageable_and_64_byte_aligned():
    img = bench.aligned_pageable(1080, 1920)
    assert img.shape == (1080, 1920) and img.dtype == np.uint32 and img.ctypes.data % 64 == 0 and img.flags.c_contiguous
    img[...] = 0x000000FF
    assert int(img[-1, -1]) == 0xFF


def test_reference_arm_prints_the_contracts_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--config", "c1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 0 and line["gpu_launches"] == 0
    assert line["config"]["workload"] == bench.CONFIGS["c1"].label
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1 and "rows 0::" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 of a torchrun launch exit without work
    import os
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True, timeout=60, cwd=ROOT,
                       env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_kernel_facts_name_their_capture():
    facts = json.loads((ROOT / "profiles" / "kernel_facts.json").read_text())
    for key in ("c1", "c2", "c3", "c4", "c5"):
        f = facts[key]
        assert f["dram_bytes_per_launch"] == f["dram_read_bytes"] + f["dram_write_bytes"] > 0
        assert 1.0 <= f["active_lanes"] <= 32.0 and "git" in f["source"] and f["kernel"].startswith("k_render_")
        assert bench.kernel_facts(key) == f


def test_only_the_json_line_reaches_stdout():
    """Libraries write to fd 1 behind Python's back (NCCL's version banner under NCCL_DEBUG=VERSION): bench.py moves fd 1 to stderr
    and writes its one line to the original stdout."""
    code = "import os, bench; bench.claim_stdout(); os.write(1, b'NCCL version x.y\\n'); print('python noise'); bench.emit({'a': 1})"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"a": 1}\n'
    assert "NCCL version x.y" in r.stderr and "python noise" in r.stderr

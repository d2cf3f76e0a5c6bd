"""bench.py's reference arm runs anywhere (it is the CPU oracle timed on the host cores): check the JSON line it
prints against the contract -- one line, the base keys, `impl`, `cpu_baseline` and an `e2e` that repeats the line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--games", "32", *extra], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_line_mcts():
    d = _run()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "mcts_simulations_per_sec" and d["unit"] == "simulations/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_line_playout():
    d = _run("--workload", "playout")
    assert d["impl"] == "reference" and d["metric"] == "playout_plies_per_sec" and d["unit"] == "plies/s" and d["value"] > 0


def test_bench_names_the_rollout_kernel_the_library_picks():
    """bench.rollout_kernel_name mirrors launch_lane_job (lane_kernels.cu): the thresholds there must be the ones here"""
    import re
    sys.path.insert(0, ROOT)
    import bench
    src = open(os.path.join(ROOT, "die_e_b200", "csrc", "lane_kernels.cu")).read()
    m = re.search(r"MODE == LANE_PLAYOUT \? (\d+) : (\d+)\);", src)
    assert m and (int(m.group(1)), int(m.group(2))) == (400, 640)
    assert "job.n_items <= 2ll * sms * PK_S ? 2 : 3" in src
    hdr = open(os.path.join(ROOT, "die_e_b200", "csrc", "lane_pack.cuh")).read()
    assert re.search(r"#define DIEE_PK_S 512", hdr)
    old = os.environ.pop("DIEE_LANE_PACK", None)
    try:
        assert bench.rollout_kernel_name(512 * 100).startswith("lane_run_kernel")            # 51,200 < 148 * 640
        assert "one wave" in bench.rollout_kernel_name(1024 * 100)                           # the headline batch
        assert "refilled" in bench.rollout_kernel_name(8192 * 100)
        assert bench.rollout_kernel_name(32768, playout=True).startswith("lane_run_kernel")  # < 148 * 400
        assert "one wave" in bench.rollout_kernel_name(65536, playout=True)
        os.environ["DIEE_LANE_PACK"] = "0"
        assert bench.rollout_kernel_name(8192 * 100).startswith("lane_run_kernel")
    finally:
        os.environ.pop("DIEE_LANE_PACK", None)
        if old is not None:
            os.environ["DIEE_LANE_PACK"] = old

"""The lane-per-game engine (die_e_b200/csrc/bg_lane.cuh) compiled for the HOST and checked against the
oracle: legal-move count, every k-th play, successor states and whole Philox playouts (tests/lane_harness.cpp).
This is the CPU-side pin of the rollout / playout kernels' arithmetic; the GPU tests pin the kernels themselves."""
import os
import subprocess

import orc

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _build():
    orc.build()
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "lane_harness")
    src = os.path.join(HERE, "lane_harness.cpp")
    hdr = os.path.join(ROOT, "die_e_b200", "csrc", "bg_lane.cuh")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src, "-L" + orc.ORACLE_DIR, "-lorc",
                        "-Wl,-rpath," + orc.ORACLE_DIR], check=True)
    return exe


def test_lane_engine_matches_oracle_on_played_and_synthetic_positions():
    exe = _build()
    for seed in (1, 2):
        r = subprocess.run([exe, "300", "400000", str(seed)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "lane engine == oracle" in r.stdout

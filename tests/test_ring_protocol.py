"""Host model of the queue protocol of the packed lane kernels (tests/ring_model.cpp, std::atomic threads in place of
warps): reserved-then-written rings, compare-and-swap claims, EMPTY sentinels.  Checks that no resident game is ever held
by two workers, lost or duplicated, and that a ring never wraps onto a live entry -- under real concurrency on the host."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _build():
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "ring_model")
    src = os.path.join(HERE, "ring_model.cpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, src], check=True)
    return exe


@pytest.mark.parametrize("workers,slots,kinds,plies", [(8, 512, 6, 3000), (16, 96, 6, 10000), (3, 33, 2, 30000), (8, 512, 1, 2000)])
def test_ring_protocol_keeps_every_game_exactly_once(workers, slots, kinds, plies):
    exe = _build()
    for seed in (1, 2, 3):
        r = subprocess.run([exe, str(workers), str(slots), str(kinds), str(plies), str(seed)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-1500:]
        assert "ring protocol ok" in r.stdout

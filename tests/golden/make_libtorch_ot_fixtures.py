"""Builds tests/golden/ot_writer.cpp against the libtorch that ships inside the torch wheel (headers + libtorch.so; g++)
and writes the committed fixtures tests/golden/libtorch_{ps,states,outcomes}.ot -- three `Tensor::save` files exactly as
tch writes them (torch::save -> OutputArchive, key "0"), holding the first three MemoryFragments of a seeded pattern.

    python tests/golden/make_libtorch_ot_fixtures.py

tests/test_ot_fixture.py reads the committed files with die_e_b200.alphazero.load_training_data, and -- where g++ and
the wheel's headers are present -- rebuilds the writer to pin nnet.load_ot on a whole VarStore archive too."""
import os
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def build_writer(out_dir):
    import torch
    tdir = os.path.dirname(torch.__file__)
    exe = os.path.join(out_dir, "ot_writer")
    src = os.path.join(HERE, "ot_writer.cpp")
    if os.path.exists(exe) and os.path.getmtime(exe) >= os.path.getmtime(src):
        return exe
    os.makedirs(out_dir, exist_ok=True)
    abi = int(torch.compiled_with_cxx11_abi())
    cmd = ["g++", "-O1", "-std=c++17", f"-I{tdir}/include", f"-I{tdir}/include/torch/csrc/api/include",
           f"-D_GLIBCXX_USE_CXX11_ABI={abi}", src, "-o", exe, f"-L{tdir}/lib", "-ltorch", "-ltorch_cpu", "-lc10",
           f"-Wl,-rpath,{tdir}/lib"]
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


def _spec(named):
    out = [struct.pack("<I", len(named))]
    for name, a in named:
        a = np.ascontiguousarray(a)
        dt = {np.dtype("float32"): 0, np.dtype("int8"): 1}[a.dtype]
        nb = name.encode()
        out += [struct.pack("<I", len(nb)), nb, struct.pack("<II", dt, a.ndim), struct.pack(f"<{a.ndim}q", *a.shape), a.tobytes()]
    return b"".join(out)


def write_multi(exe, named, path, tmp):
    """VarStore::save: named tensors -> one archive"""
    spec = os.path.join(tmp, "spec_multi.bin")
    open(spec, "wb").write(_spec(named))
    subprocess.run([exe, "multi", spec, path], check=True)


def write_single(exe, array, path, tmp):
    """Tensor::save: one tensor, key "0" """
    spec = os.path.join(tmp, "spec_single.bin")
    open(spec, "wb").write(_spec([("0", array)]))
    subprocess.run([exe, "single", spec, path], check=True)


def pattern():
    """three MemoryFragments (alphazero.rs:69-73): ps [3,1352] f32, states [3,6,4,6] f32, outcomes [3] i8"""
    rng = np.random.default_rng(0xD1EE)
    ps = np.zeros((3, 1352), np.float32)
    for i in range(3):
        idx = rng.choice(1352, 9, replace=False)
        ps[i, idx] = rng.random(9).astype(np.float32)
    states = rng.integers(-5, 6, (3, 6, 4, 6)).astype(np.float32)
    outcomes = np.array([1, -1, 0], np.int8)
    return ps, states, outcomes


if __name__ == "__main__":
    import tempfile
    exe = build_writer(os.path.join(os.path.dirname(HERE), "_build"))
    ps, states, outcomes = pattern()
    with tempfile.TemporaryDirectory() as tmp:
        write_single(exe, ps, os.path.join(HERE, "libtorch_ps.ot"), tmp)
        write_single(exe, states, os.path.join(HERE, "libtorch_states.ot"), tmp)
        write_single(exe, outcomes, os.path.join(HERE, "libtorch_outcomes.ot"), tmp)
    for f in ("libtorch_ps.ot", "libtorch_states.ot", "libtorch_outcomes.ot"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")

"""quick device-side timing of the net forward (development helper)"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from die_e_b200 import _ffi, nnet

ctx = _ffi.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
tens = nnet.synthetic_tensors(seed=1, filters=256, blocks=19)
net = _ffi.Net(ctx, tens)
for n in [int(a) for a in sys.argv[1:]] or (1024, 4096, 16384):
    s = np.zeros(n, dtype=_ffi.BG_STATE)
    s["pts"][:] = [2, 0, 0, 0, 0, -5, 0, -3, 0, 0, 0, 5, -5, 0, 0, 0, 3, 0, 5, 0, 0, 0, 0, -2]
    s["player"] = -1
    s["roll"][:] = (3, 1)
    d_s = torch.from_numpy(s.view(np.uint8).reshape(n, 32)).cuda()
    d_p = torch.empty(n, 1352, device="cuda")
    d_v = torch.empty(n, device="cuda")
    for _ in range(3):
        net.forward_dev(d_s.data_ptr(), n, d_p.data_ptr(), d_v.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    reps = 10
    for _ in range(reps):
        net.forward_dev(d_s.data_ptr(), n, d_p.data_ptr(), d_v.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"n={n}: {ms:.3f} ms/forward, {n / ms * 1e3:.0f} evals/s, {1.0825e9 * n / ms / 1e9:.1f} TFLOP/s")

"""forward latency of the 256x19 net per batch size and precision (CUDA events, device-resident)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from die_e_b200 import _ffi as ffi, nnet
ctx = ffi.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
net = ffi.Net(ctx, nnet.synthetic_tensors(seed=1, filters=256, blocks=19, bn_stats="identity"))
h = bench.midgame_states(ctx, ffi, 0, 1024)
ds = torch.from_numpy(h.view(np.uint8).reshape(-1, 32)).to(dev)
pol = torch.zeros(1024, 1352, device=dev); val = torch.zeros(1024, device=dev)
for prec, name in ((ffi.NET_BF16, "bf16"), (ffi.NET_SPLIT3, "split3")):
    net.set_precision(prec)
    row = []
    for n in (16, 64, 128, 256, 512, 1024):
        for _ in range(3): net.forward_dev(ds.data_ptr(), n, pol.data_ptr(), val.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10): net.forward_dev(ds.data_ptr(), n, pol.data_ptr(), val.data_ptr())
        e1.record(stream); e1.synchronize()
        row.append(f"{n}: {e0.elapsed_time(e1) / 10:.3f}")
    print(name, "ms per forward ->", "  ".join(row), flush=True)

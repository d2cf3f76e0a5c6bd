"""PyTorch oracle of the policy/value ResNet (src/alphazero/nnet.rs), CPU, fp64 by default.

`forward(tensors, x)` evaluates the reference architecture with the raw (unfolded) parameters.
`forward_bf16_emulated` applies the product's numeric recipe (BatchNorm folded in fp64, weights and
inter-layer activations rounded to bf16, wide accumulation) so the tensor-core kernels can be checked
tightly; the distance between the two is the precision cost of the bf16 path."""
import numpy as np
import torch
import torch.nn.functional as F


def _split(tensors, blocks):
    t = [torch.from_numpy(np.asarray(a)).double() for a in tensors]
    i = 0

    def take(n):
        nonlocal i
        r = t[i:i + n]
        i += n
        return r
    init = take(2) + take(4)
    blks = []
    for _ in range(blocks):
        c1, c2, b1, b2 = take(2), take(2), take(4), take(4)
        blks.append((c1, c2, b1, b2))
    pol = (take(2), take(4), take(2))
    val = (take(2), take(4), take(2))
    assert i == len(t)
    return init, blks, pol, val


def _bn(x, p):
    g, b, m, v = p
    return F.batch_norm(x, m, v, g, b, training=False, eps=1e-5)


def forward(tensors, x, blocks, dtype=torch.float64):
    """x: [N,6,4,6] (as_tensor).  -> policy [N,1352] softmaxed, value [N]"""
    init, blks, pol, val = _split(tensors, blocks)
    cast = lambda ts: [a.to(dtype) for a in ts]
    x = torch.from_numpy(np.asarray(x)).to(dtype)
    w, b = cast(init[:2])
    x = F.relu(_bn(F.conv2d(x, w, b, padding=1), cast(init[2:])))
    for c1, c2, b1, b2 in blks:
        y = F.relu(_bn(F.conv2d(x, *cast(c1), padding=1), cast(b1)))
        y = _bn(F.conv2d(y, *cast(c2), padding=1), cast(b2))
        x = F.relu(y + x)
    p = F.relu(_bn(F.conv2d(x, *cast(pol[0]), padding=1), cast(pol[1]))).flatten(1)
    p = F.softmax(F.linear(p, *cast(pol[2])), dim=1)
    v = F.relu(_bn(F.conv2d(x, *cast(val[0]), padding=1), cast(val[1]))).flatten(1)
    v = torch.tanh(F.linear(v, *cast(val[2]))).reshape(-1)
    return p.double().numpy(), v.double().numpy()


def _fold(conv, bn):
    w, b = conv
    g, beta, m, v = bn
    scale = g / torch.sqrt(v + 1e-5)
    return w * scale.view(-1, 1, 1, 1), (b - m) * scale + beta


def _bf16(x):
    return x.float().bfloat16().double()


def forward_bf16_emulated(tensors, x, blocks):
    init, blks, pol, val = _split(tensors, blocks)
    x = torch.from_numpy(np.asarray(x)).double()
    w, b = _fold(init[:2], init[2:])
    x = _bf16(F.relu(F.conv2d(_bf16(x), _bf16(w), b.float().double(), padding=1)))
    for c1, c2, b1, b2 in blks:
        w1, bb1 = _fold(c1, b1)
        w2, bb2 = _fold(c2, b2)
        y = _bf16(F.relu(F.conv2d(x, _bf16(w1), bb1.float().double(), padding=1)))
        x = _bf16(F.relu(F.conv2d(y, _bf16(w2), bb2.float().double(), padding=1) + x))
    wp, bp = _fold(pol[0], pol[1])
    p = _bf16(F.relu(F.conv2d(x, _bf16(wp), bp.float().double(), padding=1))).flatten(1)
    p = F.softmax(F.linear(p, _bf16(pol[2][0]), pol[2][1].float().double()), dim=1)
    wv, bv = _fold(val[0], val[1])
    v = F.relu(F.conv2d(x, _bf16(wv), bv.float().double(), padding=1)).float().double().flatten(1)
    v = torch.tanh(F.linear(v, val[2][0].float().double(), val[2][1].float().double())).reshape(-1)
    return p.numpy(), v.numpy()

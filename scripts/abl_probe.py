"""Timing experiments for the training MLP kernels under NB_TC_ABLATE (set in the environment BEFORE the process starts):
   0   as shipped
   16  forward: no stash copy-out
   64  forward stash / dgrad dY written to a 64-tile window (stays in L2); wgrad reads dY from that window
   128 wgrad reads X from a 64-tile window too
Prints one JSON line per kernel: fwd<train>, dgrad (stage 1), wgrad (stage 2), plus pure HBM write / read rates (torch fill / sum).
Results under ablation are NOT valid gradients -- timing only."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_pytorch_paeng_b200.engine import get_engine  # noqa: E402
from nerf_pytorch_paeng_b200.model import NeRF  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device('cuda', 0)
    eng = get_engine(dev)
    abl = int(os.environ.get('NB_TC_ABLATE', '0'))
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev)
    net.set_precision('bf16')
    m = net.model_fine
    flat, pk = m.flat_params(), m.packed_weights()
    n = 4096
    rays = torch.cat([torch.zeros(n, 3, device=dev), torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)], -1)
    rays[:, 2] = 4.0
    for S in (64, 192):
        P = n * S
        z = torch.sort(torch.rand(n, S, device=dev) * 4 + 2, -1)[0]
        t_inf = timeit(lambda: eng.mlp_forward(m.desc, flat, pk, m.precision, rays=rays, z=z, save=False), iters=20)
        t_fwd = timeit(lambda: eng.mlp_forward(m.desc, flat, pk, m.precision, rays=rays, z=z, save=True), iters=20)
        raw, act = eng.mlp_forward(m.desc, flat, pk, m.precision, rays=rays, z=z, save=True)
        d_raw = torch.randn_like(raw) * 1e-3
        grad = torch.zeros_like(flat)
        t_dg = timeit(lambda: eng.mlp_backward(m.desc, flat, pk, m.precision, P, act, d_raw, grad, stage=1), iters=20)
        t_wg = timeit(lambda: eng.mlp_backward(m.desc, flat, pk, m.precision, P, act, d_raw, grad, stage=2), iters=20)
        print(json.dumps({'abl': abl, 'points': P, 'fwd_infer_ms': t_inf, 'fwd_train_ms': t_fwd, 'dgrad_ms': t_dg, 'wgrad_ms': t_wg,
                          'fwd_train_tflops': 1186816 * P / t_fwd / 1e9, 'dgrad_tflops': 1115392 * P / t_dg / 1e9,
                          'wgrad_tflops': 1186816 * P / t_wg / 1e9}), flush=True)
        del act
    if abl == 0:
        big = torch.empty(1 << 30, dtype=torch.float32, device=dev)      # 4 GiB
        t_w = timeit(lambda: big.fill_(1.0), iters=5)
        t_r = timeit(lambda: big.sum(), iters=5)
        b2 = torch.empty_like(big)
        t_c = timeit(lambda: b2.copy_(big), iters=5)
        print(json.dumps({'hbm_write_GBps': big.numel() * 4 / t_w / 1e6, 'hbm_read_GBps': big.numel() * 4 / t_r / 1e6,
                          'hbm_copy_GBps': 2 * big.numel() * 4 / t_c / 1e6}), flush=True)


if __name__ == '__main__':
    main()

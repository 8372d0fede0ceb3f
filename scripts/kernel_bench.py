"""Per-kernel timings on the B200 (CUDA events on the launching stream, inputs larger than L2 where it
matters).  Usage: python scripts/kernel_bench.py [--what fwd,render,...]  -> JSON lines."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_pytorch_paeng_b200.engine import get_engine  # noqa: E402
from nerf_pytorch_paeng_b200.model import NeRF  # noqa: E402

PEAK_TF, PEAK_HBM = 1373.4, 6549.8
try:
    _p = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
    PEAK_TF, PEAK_HBM = _p['bf16_tflops_sustained'], _p['hbm_gbs']
    PEAK_TF_BURST = _p['bf16_tflops']
except Exception:
    PEAK_TF_BURST = 1644.5


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
    ap = argparse.ArgumentParser()
    ap.add_argument('--what', default='fwd,hbm')
    ap.add_argument('--rays', type=int, default=4096)
    args = ap.parse_args()
    what = args.what.split(',')
    dev = torch.device('cuda', 0)
    eng = get_engine(dev)
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev)
    n = args.rays
    rays = torch.cat([torch.zeros(n, 3, device=dev), torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)], -1)
    rays[:, 2] = 4.0
    out = []
    if 'fwd' in what:
        for prec in ('bf16', 'fp32'):
            net.set_precision(prec)
            m = net.model_fine
            flat = m.flat_params()
            pk = m.packed_weights()
            for S in (64, 192):
                z = torch.sort(torch.rand(n, S, device=dev) * 4 + 2, -1)[0]
                for save in (False, True):
                    if prec == 'fp32' and save:
                        continue
                    ms = timeit(lambda: eng.mlp_forward(m.desc, flat, pk, m.precision, rays=rays, z=z, save=save), iters=5 if prec == 'fp32' else 20)
                    fl = 1186816 * n * S
                    out.append({'kernel': f'mlp_forward_{prec}', 'points': n * S, 'save': save, 'ms': ms, 'tflops': fl / ms / 1e9,
                                'frac_of_sustained_bf16': fl / ms / 1e9 / PEAK_TF, 'frac_of_burst_bf16': fl / ms / 1e9 / PEAK_TF_BURST})
    if 'bwd' in what:
        net.set_precision('bf16')
        m = net.model_fine
        flat = m.flat_params()
        pk = m.packed_weights()
        for S in (64, 192):
            z = torch.sort(torch.rand(n, S, device=dev) * 4 + 2, -1)[0]
            raw, act = eng.mlp_forward(m.desc, flat, pk, m.precision, rays=rays, z=z, save=True)
            d_raw = torch.randn_like(raw) * 1e-3
            grad = torch.zeros_like(flat)
            ms = timeit(lambda: eng.mlp_backward(m.desc, flat, pk, m.precision, n * S, act, d_raw, grad), iters=10)
            fl = 2302208 * n * S
            out.append({'kernel': 'mlp_backward_bf16', 'points': n * S, 'ms': ms, 'tflops': fl / ms / 1e9,
                        'frac_of_sustained_bf16': fl / ms / 1e9 / PEAK_TF})
    if 'hbm' in what:
        big = 1 << 20   # rays: inputs far larger than L2
        S = 192
        raw = torch.randn(big, S, 4, device=dev)
        z = torch.sort(torch.rand(big, S, device=dev) * 4 + 2, -1)[0]
        d = torch.randn(big, 3, device=dev)
        ms = timeit(lambda: eng.composite_forward(raw, z, d), iters=5)
        by = big * (S * 20 + S * 4 + 24 + 12)
        out.append({'kernel': 'composite_forward', 'rays': big, 'S': S, 'ms': ms, 'GBps': by / ms / 1e6, 'frac_hbm': by / ms / 1e6 / PEAK_HBM})
        g = torch.randn(big, 3, device=dev)
        ms = timeit(lambda: eng.composite_backward(raw, z, d, g), iters=5)
        by = big * (S * 20 + S * 16 + 24)
        out.append({'kernel': 'composite_backward', 'rays': big, 'S': S, 'ms': ms, 'GBps': by / ms / 1e6, 'frac_hbm': by / ms / 1e6 / PEAK_HBM})
        del raw
        zc = torch.sort(torch.rand(big, 64, device=dev) * 4 + 2, -1)[0]
        w = torch.rand(big, 64, device=dev)
        ms = timeit(lambda: eng.sample_pdf(zc, w, 128, u=None, seed=1), iters=5)
        by = big * (256 + 256 + 768)
        out.append({'kernel': 'sample_pdf(philox u)', 'rays': big, 'ms': ms, 'GBps': by / ms / 1e6, 'frac_hbm': by / ms / 1e6 / PEAK_HBM})
        lower = torch.linspace(2, 6, 64, device=dev)
        span = torch.full((64,), 0.06, device=dev)
        big2 = 1 << 23
        ms = timeit(lambda: eng.stratified(big2, lower, span, None, 1, 0), iters=5)
        by = big2 * 256
        out.append({'kernel': 'stratified(philox)', 'rays': big2, 'ms': ms, 'GBps': by / ms / 1e6, 'frac_hbm': by / ms / 1e6 / PEAK_HBM})
        K = np.array([[1111.111, 0, 2000.], [0, 1111.111, 2000.], [0, 0, 1.]])
        pose = torch.eye(4, device=dev)[:3]
        ms = timeit(lambda: eng.raygen(4000, 4000, K, pose), iters=5)
        by = 16000000 * 24
        out.append({'kernel': 'raygen_pinhole', 'rays': 16000000, 'ms': ms, 'GBps': by / ms / 1e6, 'frac_hbm': by / ms / 1e6 / PEAK_HBM})
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == '__main__':
    main()

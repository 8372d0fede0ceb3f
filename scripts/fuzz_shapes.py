"""One-off robustness sweep: random (rays, S_c, S_f) shapes through the fused train step, bf16 vs fp32 path and fused vs
stage-by-stage enqueueing; checks finiteness, identical renders between the two enqueue routes and bf16/fp32 agreement."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerf_pytorch_paeng_b200 import nerf_process as NP, trainer  # noqa: E402
from nerf_pytorch_paeng_b200.model import NeRF  # noqa: E402

dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev)
rs = np.random.RandomState(int(os.environ.get('FUZZ_SEED', '0')))
bad = 0
for trial in range(int(os.environ.get('FUZZ_N', '40'))):
    n = int(rs.choice([1, 2, 3, 31, 33, 127, 128, 129, 255, 500, 1000, 1531, 2048, 4099]))
    sc = int(rs.choice([4, 8, 16, 32, 64, 96, 128]))
    sf = int(rs.choice([0, 4, 32, 64, 128, 160, 256]))
    perturb = float(rs.choice([0., 1.]))
    rays = torch.cat([torch.randn(n, 3, device=dev) * .2 + torch.tensor([0., 0., 4.], device=dev),
                      torch.nn.functional.normalize(torch.randn(n, 3, device=dev) * .3 + torch.tensor([0., 0., -1.], device=dev), dim=-1)], -1)
    tgt = torch.rand(n, 3, device=dev)
    res = {}
    for prec in ('fp32', 'bf16'):
        model.set_precision(prec)
        for fused in (True, False):
            opts = bench.make_opts(rank_dev=0, seed=trial, N_samples_c=sc, N_samples_f=sf, perturb=perturb, fused_driver=fused)
            NP._counter[0] = 0
            for net in (model.model_coarse, model.model_fine):
                net.bind_flat_grad().zero_()
            out = trainer.render_losses_and_grads(model, rays, tgt, opts)
            torch.cuda.synchronize()
            g = torch.cat([model.model_coarse.flat_grad, model.model_fine.flat_grad]).clone()
            res[(prec, fused)] = (out, g)
            for k, v in out.items():
                if not torch.isfinite(v).all():
                    print('NONFINITE', trial, n, sc, sf, prec, fused, k); bad += 1
            if not torch.isfinite(g).all():
                print('NONFINITE grad', trial, n, sc, sf, prec, fused); bad += 1
    for prec in ('fp32', 'bf16'):
        a, b = res[(prec, True)], res[(prec, False)]
        for k in a[0]:
            if k != 'loss_buf' and not torch.equal(a[0][k], b[0][k]):
                print('FUSED!=STEPS', trial, n, sc, sf, prec, k, float((a[0][k] - b[0][k]).abs().max())); bad += 1
        gn = float(b[1].norm())
        if gn > 0 and float((a[1] - b[1]).norm()) / gn > 1e-3:
            print('FUSED grad != STEPS', trial, n, sc, sf, prec, float((a[1] - b[1]).norm()) / gn); bad += 1
    key = 'rgb_f' if sf > 0 else 'rgb_c'
    # A ray whose LAST sample has sigma ~ 0 is a step function of sign(sigma): the reference gives that sample a 1e10-long
    # interval (nerf_process.py:98), so alpha is 0 or 1.  With random-init weights (|sigma| ~ 1e-3) the ~4e-3 bf16 noise flips a
    # few such rays by up to 0.5; everything else must agree closely.
    dr = (res[('bf16', True)][0][key] - res[('fp32', True)][0][key]).abs().max(-1)[0]
    frac = float((dr > 3e-2).float().mean())
    med = float(dr.median())
    if (frac > 0.02 and int((dr > 3e-2).sum()) > 8) or (med > 1e-2 and n >= 16):      # (a median over a handful of rays is one ray's tail)
        print('BF16 vs FP32 render', trial, n, sc, sf, perturb, frac, med); bad += 1
    print('trial', trial, 'n', n, 'S', sc, sf, 'perturb', perturb, 'bf16-fp32 median', f'{med:.2e}', 'rays off by > 3e-2:', f'{frac:.4f}', flush=True)
print('BAD', bad)

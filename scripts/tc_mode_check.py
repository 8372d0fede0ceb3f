"""Forward + backward of the bf16 MLP under the cluster mode given by NB_TC_CLUSTER, against the fp32 path.
Run as a subprocess (the mode is read once per process).  Prints 'OK <rel errors>' or raises."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_pytorch_paeng_b200._lib import NB_BF16, NB_FP32  # noqa: E402
from nerf_pytorch_paeng_b200.engine import get_engine  # noqa: E402
from nerf_pytorch_paeng_b200.model import NeRF  # noqa: E402

dev = torch.device('cuda', 0)
eng = get_engine(dev)
torch.manual_seed(0)
net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev)
m = net.model_coarse
n, S = 700, 77                      # 421 tiles + ragged tail; odd tile count -> ghost tile in cluster modes
rays = torch.cat([torch.zeros(n, 3, device=dev), torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)], -1)
rays[:, 2] = 4.0
z = torch.sort(torch.rand(n, S, device=dev) * 4 + 2, -1)[0]
d_raw = torch.randn(n * S, 4, device=dev) * 1e-2
res = {}
for prec in (NB_FP32, NB_BF16):
    m.precision = prec
    flat = m.flat_params()
    raw, act = eng.mlp_forward(m.desc, flat, m.packed_weights(), prec, rays=rays, z=z, save=True)
    grad = torch.empty_like(flat)
    eng.mlp_backward(m.desc, flat, m.packed_weights(), prec, n * S, act, d_raw, grad)
    torch.cuda.synchronize()
    res[prec] = (raw.clone(), grad.clone())
e_raw = float((res[NB_BF16][0] - res[NB_FP32][0]).norm() / res[NB_FP32][0].norm())
e_grad = float((res[NB_BF16][1] - res[NB_FP32][1]).norm() / res[NB_FP32][1].norm())
assert np.isfinite(e_raw) and e_raw < 2e-2, e_raw
assert np.isfinite(e_grad) and e_grad < 0.15, e_grad        # adversarial i.i.d. d_raw (see test_tc_backward_vs_fp32)
print(f'OK mode={os.environ.get("NB_TC_CLUSTER", "default")} raw_rel={e_raw:.4f} grad_rel={e_grad:.4f}')

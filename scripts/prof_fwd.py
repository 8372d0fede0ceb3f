"""Tiny driver for ncu: a few launches of the bf16 forward chain (inference and training variants)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_pytorch_paeng_b200.engine import get_engine
from nerf_pytorch_paeng_b200.model import NeRF
dev = torch.device('cuda', 0)
eng = get_engine(dev)
torch.manual_seed(0)
net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev).set_precision('bf16')
m = net.model_fine
flat = m.flat_params(); pk = m.packed_weights()
n, S = 4096, 192
rays = torch.cat([torch.zeros(n, 3, device=dev), torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)], -1)
z = torch.sort(torch.rand(n, S, device=dev) * 4 + 2, -1)[0]
for save in (False, False, False, True, True):
    raw, act = eng.mlp_forward(m.desc, flat, pk, m.precision, rays=rays, z=z, save=save)
d_raw = torch.randn_like(raw) * 1e-3
grad = torch.zeros_like(flat)
for _ in range(2):
    eng.mlp_backward(m.desc, flat, pk, m.precision, n * S, act, d_raw, grad)
torch.cuda.synchronize()
print('done')

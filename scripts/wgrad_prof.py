"""Per-CTA time stamps of mlp_wgrad_kernel (NB_TC_PROF diagnostic in nb_mlp_tc_bwd.cu): where does the kernel's time go, and which
CTAs finish last?  Prints the library's per-CTA table (stderr) for the fine pass of a 4096-ray step; timing only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_pytorch_paeng_b200.engine import get_engine  # noqa: E402
from nerf_pytorch_paeng_b200.model import NeRF  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    eng = get_engine(dev)
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev)
    net.set_precision('bf16')
    m = net.model_fine
    flat, pk = m.flat_params(), m.packed_weights()
    n, S = 4096, int(os.environ.get('S', '192'))
    rays = torch.cat([torch.zeros(n, 3, device=dev), torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)], -1)
    rays[:, 2] = 4.0
    z = torch.sort(torch.rand(n, S, device=dev) * 4 + 2, -1)[0]
    raw, act = eng.mlp_forward(m.desc, flat, pk, m.precision, rays=rays, z=z, save=True)
    d_raw = torch.randn_like(raw) * 1e-3
    grad = torch.zeros_like(flat)
    eng.mlp_backward(m.desc, flat, pk, m.precision, n * S, act, d_raw, grad, stage=1)
    torch.cuda.synchronize()
    os.environ['NB_TC_PROF'] = '1'
    for _ in range(3):
        eng.mlp_backward(m.desc, flat, pk, m.precision, n * S, act, d_raw, grad, stage=2)
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()

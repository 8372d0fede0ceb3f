"""MMA-warp cycle counters of the forward chain kernel per CTA (NB_TC_PROF diagnostic in nb_mlp_tc.cu), inference and training."""
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
    for save in (False, True):
        for _ in range(3):
            eng.mlp_forward(m.desc, flat, pk, m.precision, rays=rays, z=z, save=save)
        torch.cuda.synchronize()
    os.environ['NB_TC_PROF'] = '1'
    for save in (False, True):
        for _ in range(2):
            eng.mlp_forward(m.desc, flat, pk, m.precision, rays=rays, z=z, save=save)
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()

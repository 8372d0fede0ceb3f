"""Two 800x800 coarse+fine frames through trainer.render_frame (the path bench.py's `render` key times); for ncu."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerf_pytorch_paeng_b200 import trainer  # noqa: E402
from nerf_pytorch_paeng_b200.model import NeRF  # noqa: E402

dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev).set_precision('bf16')
opts = bench.make_opts(rank_dev=0, seed=1)
K = np.array([[bench.FOCAL, 0, 400.], [0, bench.FOCAL, 400.], [0, 0, 1.]])
poses = torch.from_numpy(bench.synthetic_poses(2, seed=0)).to(dev)
for i in range(2):
    rgb, disp = trainer.render_frame(model, 800, 800, K, poses[i, :3, :4], opts)
torch.cuda.synchronize()
print('ok', float(rgb.mean()))

"""BASELINE config 5: sweep of samples per ray (64+128 .. 256+512) and batch size (4k .. 64k rays): fused train step
(render + loss + backward + Adam) timings on one B200.  JSON lines -> profiles/."""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_pytorch_paeng_b200 import trainer  # noqa: E402
from nerf_pytorch_paeng_b200.engine import get_engine  # noqa: E402
from nerf_pytorch_paeng_b200.model import NeRF  # noqa: E402

dev = torch.device('cuda', 0)
eng = get_engine(dev)
torch.manual_seed(0)
net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev).set_precision('bf16')
opt = trainer.FlatAdam(net, lr=5e-4)
K = np.array([[1111.111, 0, 400.], [0, 1111.111, 400.], [0, 0, 1.]])
pose = torch.tensor([[1., 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 4.]], device=dev)
peak = 1373.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['bf16_tflops_sustained']
except Exception:
    pass
for (sc, sf) in ((64, 128), (128, 256), (256, 512)):
    for n in (4096, 8192, 16384, 32768, 65536):
        opts = SimpleNamespace(near=2., far=6., N_samples_c=sc, N_samples_f=sf, perturb=1., data_type='blender', gpu_ids=[0], rank=0,
                               chunk_rays=n, chunk_pts=524288, N_rays=n, seed=1)
        pix = eng.select_pixels(n, 800, 800, seed=n + sc)
        o, d = eng.raygen(800, 800, K, pose, pix_idx=pix)
        rays = torch.cat((o, d), -1)
        target = torch.rand(n, 3, device=dev)
        iters = 3 if n * (sc + sf) > 8e6 else 10
        for _ in range(2):
            trainer.train_step(net, opt, rays, target, opts)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            trainer.train_step(net, opt, rays, target, opts)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        pts = n * (2 * sc + sf)
        fl = 3489024 * pts
        print(json.dumps({'S_c': sc, 'S_f': sf, 'rays': n, 'points_per_step': pts, 'ms_per_step': ms, 'rays_per_s': n / ms * 1e3,
                          'tflops': fl / ms / 1e9, 'frac_of_sustained_bf16': fl / ms / 1e9 / peak,
                          'passes': int(np.ceil(n * (sc + sf) / trainer.MAX_POINTS_PER_PASS))}), flush=True)
        del rays, target
        torch.cuda.empty_cache()

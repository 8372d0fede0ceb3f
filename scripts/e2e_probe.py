"""Where does the end-to-end step go?  CPU enqueue time of train.train vs device time per step (diagnostic)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerf_pytorch_paeng_b200 import train as train_mod, trainer  # noqa: E402
from nerf_pytorch_paeng_b200.engine import get_engine  # noqa: E402
from nerf_pytorch_paeng_b200.model import NeRF, get_positional_encoder  # noqa: E402

dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
eng = get_engine(dev)
H = W = 800
torch.manual_seed(0)
model = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev)
model.set_precision('bf16')
opts = bench.make_opts(rank_dev=0, seed=1000, device_select=True)
K = np.array([[bench.FOCAL, 0, 400.], [0, bench.FOCAL, 400.], [0, 0, 1.]])
poses = bench.synthetic_poses(16, seed=0)
optimizer = trainer.FlatAdam(model, lr=5e-4)
posenc = [get_positional_encoder(10)[0], get_positional_encoder(4)[0]]
n_img = 4
images = [torch.rand(H, W, 3).pin_memory() for _ in range(n_img)]
gt_cam = (K, poses[:n_img])
crit = torch.nn.MSELoss()
host_loss = torch.zeros(1).pin_memory()


def one(i, sync):
    t0 = time.perf_counter()
    loss = train_mod.train(i + 1, list(range(n_img)), images, gt_cam, (H, W), model, crit, posenc, optimizer, None, None, opts)
    t1 = time.perf_counter()
    if sync:
        host_loss.copy_(loss.reshape(1), non_blocking=False)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


for i in range(5):
    one(i, True)
torch.cuda.synchronize()
for sync in (True, False):
    enq, wait = [], []
    t0 = time.perf_counter()
    for i in range(30):
        a, b = one(i, sync)
        enq.append(a)
        wait.append(b)
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) / 30
    print(f'sync_each_step={sync}: step {tot*1e3:.3f} ms, cpu enqueue {np.median(enq)*1e3:.3f} ms, wait {np.median(wait)*1e3:.3f} ms')
# device-only reference: resident rays
rays = torch.rand(4096, 6, device=dev)
rays[:, 2] = 4
tgt = torch.rand(4096, 3, device=dev)
for i in range(3):
    trainer.train_step(model, optimizer, rays, tgt, opts)
torch.cuda.synchronize()
t0 = time.perf_counter()
enq = []
for i in range(30):
    a = time.perf_counter()
    trainer.train_step(model, optimizer, rays, tgt, opts)
    enq.append(time.perf_counter() - a)
torch.cuda.synchronize()
print(f'resident train_step: {(time.perf_counter()-t0)/30*1e3:.3f} ms/step, cpu enqueue {np.median(enq)*1e3:.3f} ms')
if os.environ.get('E2E_CPROFILE'):
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for i in range(50):
        one(i, False)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(45)

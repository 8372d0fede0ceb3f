"""End-to-end training demonstration on a 3-D-consistent synthetic scene (no dataset needed).

Ground truth: soft coloured Gaussian blobs, volume-rendered analytically (256 uniform samples per ray through the library's
own compositing kernel) from poses on the radius-4 sphere.  A NeRF (8x256, coarse+fine, 64+128 samples) is then trained
with the fused step (bf16 tcgen05 path, optionally the fp32 parity path for the same number of steps) and evaluated on
held-out poses.  JSON lines: step, loss, held-out PSNR, wall time.   usage: python scripts/train_demo.py [--steps 3000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (synthetic_poses, make_opts)
from nerf_pytorch_paeng_b200 import trainer  # noqa: E402
from nerf_pytorch_paeng_b200.engine import get_engine  # noqa: E402
from nerf_pytorch_paeng_b200.model import NeRF  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=3000)
ap.add_argument('--fp32-steps', dest='fp32_steps', type=int, default=300)
ap.add_argument('--res', type=int, default=200)
args = ap.parse_args()

dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
eng = get_engine(dev)
H = W = args.res
focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
K = np.array([[focal, 0, .5 * W], [0, focal, .5 * H], [0, 0, 1.]])
poses = torch.from_numpy(bench.synthetic_poses(28, seed=1)).to(dev)
train_ids, test_ids = list(range(24)), [24, 25, 26, 27]

rs = np.random.RandomState(0)
centres = torch.tensor(rs.uniform(-0.8, 0.8, (6, 3)), dtype=torch.float32, device=dev)
radii = torch.tensor(rs.uniform(0.25, 0.45, 6), dtype=torch.float32, device=dev)
colours = torch.tensor(rs.uniform(0.1, 0.9, (6, 3)), dtype=torch.float32, device=dev)


def field(x):
    """x [P,3] -> raw [P,4] = [logit(rgb), sigma] of the blob scene."""
    d2 = ((x[:, None, :] - centres[None]) ** 2).sum(-1)
    w = torch.exp(-d2 / (2 * radii[None] ** 2))
    sigma = 25. * w.sum(-1)
    rgb = (w @ colours) / (w.sum(-1, keepdim=True) + 1e-6)
    rgb = rgb.clamp(0.02, 0.98)
    return torch.cat([torch.log(rgb / (1 - rgb)), sigma[:, None]], -1)


@torch.no_grad()
def gt_image(pose):
    out = torch.empty(H * W, 3, device=dev)
    zs = torch.linspace(2., 6., 256, device=dev)
    for s in range(0, H * W, 8192):
        pix = torch.arange(s, min(H * W, s + 8192), device=dev)
        o, d = eng.raygen(H, W, K, pose, pix_idx=pix)
        z = zs[None].expand(o.shape[0], -1).contiguous()
        pts = o[:, None] + d[:, None] * z[..., None]
        raw = field(pts.reshape(-1, 3)).view(o.shape[0], 256, 4).contiguous()
        out[s:s + o.shape[0]] = eng.composite_forward(raw, z, d.contiguous())[0]
    return out


images = torch.stack([gt_image(poses[i, :3, :4]) for i in range(poses.shape[0])])       # [28, H*W, 3]


def psnr_heldout(model, opts):
    vals = []
    for i in test_ids:
        rgb, _ = trainer.render_frame(model, H, W, K, poses[i, :3, :4], opts)
        vals.append(float(-10. * torch.log10(((rgb - images[i]) ** 2).mean())))
    return float(np.mean(vals))


def run(precision, steps, log_every):
    torch.manual_seed(0)
    model = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev)
    model.set_precision(precision)
    opt = trainer.FlatAdam(model, lr=5e-4)
    opts = bench.make_opts(rank_dev=0, seed=7)
    gen = np.random.RandomState(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    warm = max(steps // 10, 1)
    for it in range(1, steps + 1):
        # the reference's schedule shape (scheduler.py: linear warm-up 5e-5 -> 5e-4, then cosine back to 5e-5), scaled to `steps`
        if it <= warm:
            opt.param_groups[0]['lr'] = 5e-5 + (5e-4 - 5e-5) * it / warm
        else:
            opt.param_groups[0]['lr'] = 5e-5 + 0.5 * (5e-4 - 5e-5) * (1 + np.cos(np.pi * (it - warm) / max(steps - warm, 1)))
        i = train_ids[gen.randint(len(train_ids))]
        pix = eng.select_pixels(4096, H, W, seed=11 + i, offset=it * 4096)
        o, d = eng.raygen(H, W, K, poses[i, :3, :4], pix_idx=pix)
        loss = trainer.train_step(model, opt, torch.cat((o, d), -1), eng.gather_rows(images[i], pix), opts)
        if it % log_every == 0 or it == steps:
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
            print(json.dumps({'precision': precision, 'step': it, 'loss_c': float(loss[0]), 'loss_f': float(loss[1]),
                              'train_psnr_f': float(-10 * np.log10(float(loss[1]))), 'heldout_psnr': psnr_heldout(model, opts),
                              'train_seconds': t, 'rays_per_s': it * 4096 / t}), flush=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter() - t          # exclude the evaluation from the training clock


run('bf16', args.steps, max(args.steps // 6, 1))
if args.fp32_steps > 0:
    run('bf16', args.fp32_steps, args.fp32_steps)
    run('fp32', args.fp32_steps, args.fp32_steps)

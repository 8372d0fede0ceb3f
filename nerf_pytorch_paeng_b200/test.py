"""Drop-in for the hot-path part of the reference's test.py (file:line refs into the reference):
``test`` (test.py:17-108) and ``render`` (test.py:111-174) loop over poses and call the frame
renderer (make_o_d -> batchify, test.py:38-40 / 143-145).  PNG/GIF/MP4 writing and SSIM/LPIPS
(third-party nets) are out of scope: frames are returned as uint8 arrays (and written only when
imageio is importable); PSNR is computed on the device."""
import os

import numpy as np
import torch

from . import trainer
from .engine import get_engine
from .nerf_process import apply_precision
from .config import LOG_DIR
from .utils import img2mse, mse2psnr, to8b


def _load_checkpoint(model, opts, idx):
    """test.py:20-21 / 128-130: the reference torch.load()s unconditionally, so a wrong idx / exp_name raises; so does this.
    opts.allow_missing_checkpoint=True evaluates the weights already in memory instead (benchmarks, tests)."""
    path = os.path.join(LOG_DIR, opts.exp_name, opts.exp_name + '_{}.pth.tar'.format(idx))
    if not os.path.exists(path):
        if getattr(opts, 'allow_missing_checkpoint', False):
            return
        raise FileNotFoundError(f'checkpoint {path} not found (set opts.allow_missing_checkpoint=True to evaluate the weights in memory)')
    model.load_state_dict(torch.load(path, map_location='cpu')['model_state_dict'])


def _frames(model, poses, K, hw, opts, dist_ctx):
    img_h, img_w = hw
    device = torch.device(f'cuda:{opts.gpu_ids[opts.rank]}')
    for pose in poses:
        pose = torch.as_tensor(np.asarray(pose.cpu() if isinstance(pose, torch.Tensor) else pose), dtype=torch.float32).to(device)
        rgb, disp = trainer.render_frame(model, img_h, img_w, K, pose[:3, :4], opts, dist_ctx=dist_ctx)
        yield rgb.view(img_h, img_w, 3), disp.view(img_h, img_w)


def test(idx, i_test, posenc, model, test_imgs, gt_intrinsic, gt_extrinsic, hw, opts, dist_ctx=None, save=True):
    model.eval()
    _load_checkpoint(model, opts, idx)
    apply_precision(model, opts)
    out_dir = os.path.join(LOG_DIR, opts.exp_name, opts.exp_name + '_{}'.format(idx), 'test_result')
    psnrs, frames = [], []
    for i, (rgb, disp) in enumerate(_frames(model, gt_extrinsic, gt_intrinsic, hw, opts, dist_ctx)):
        gt = torch.as_tensor(test_imgs[i], dtype=torch.float32).to(rgb.device)
        psnr = mse2psnr(img2mse(rgb, gt).reshape(1))
        psnrs.append(float(psnr))
        rgb8_d, disp8_d = get_engine(rgb.device).frame_to8b(rgb, disp)      # to8b + disp/nanmax(disp) on the device
        rgb8 = rgb8_d.cpu().numpy()
        frames.append(rgb8)
        if save:
            _imwrite(out_dir, '{:03d}.png'.format(i), rgb8)
            _imwrite(out_dir, '{:03d}_disp.png'.format(i), disp8_d.cpu().numpy())
        print('idx:{} | PSNR:{}'.format(i, psnrs[-1]))
    return {'psnr': psnrs, 'frames': frames}


def render(idx, posenc, model, gt_intrinsic, render_pose, hw, opts, dist_ctx=None, save=True):
    model.eval()
    _load_checkpoint(model, opts, idx)
    apply_precision(model, opts)
    out_dir = os.path.join(LOG_DIR, opts.exp_name, opts.exp_name + '_{}'.format(idx), 'render_result')
    rgbs, disps = [], []
    for i, (rgb, disp) in enumerate(_frames(model, render_pose, gt_intrinsic, hw, opts, dist_ctx)):
        rgb8_d, disp8_d = get_engine(rgb.device).frame_to8b(rgb, disp)
        rgbs.append(rgb8_d.cpu().numpy())
        disps.append(disp8_d.cpu().numpy())
        if save:
            _imwrite(out_dir, f'{i}_rgb.png', rgbs[-1])
            _imwrite(out_dir, f'{i}_disp.png', disps[-1])
    return np.stack(rgbs, 0), np.stack(disps, 0)


def _imwrite(out_dir, name, arr):
    try:
        import imageio
    except Exception:
        return
    os.makedirs(out_dir, exist_ok=True)
    imageio.imwrite(os.path.join(out_dir, name), arr)

"""Drop-in for the reference's train.py:12-119 (file:line refs into the reference).

Same signature.  What changes underneath: the selected rays are generated directly by the ray-gen
kernel (only N_rays of the H*W pixels), render + loss + backward run as the fused
no-autograd step of trainer.py when `criterion` is an MSELoss, and gradients are exposed to the
caller's optimizer through p.grad views of the flat gradient buffers (so torch.optim.Adam keeps
working) -- or updated by trainer.FlatAdam in one launch per network.  Visdom plotting and the
matplotlib camera plot (train.py:73-102,117-119) are out of scope; the print line and the
checkpoint dict/cadence (train.py:105-114) are kept.
"""
import os

import numpy as np
import torch

from . import trainer
from .config import LOG_DIR
from .nerf_process import _seed, apply_precision, batchify_rays_and_render_by_chunk, ndc_rays
from .rays import make_o_d_selected
from .utils import mse2psnr

_cache = {}
_streams = {}


def _side_stream(device):
    key = str(device)
    if key not in _streams:
        _streams[key] = torch.cuda.Stream(device=device)
    return _streams[key]


def _device_copy(arr, device, key):
    """K / poses: the reference re-uploads them every step (train.py:18-21); here once per array."""
    k = (key, id(arr), str(device))
    if k not in _cache:
        t = arr if isinstance(arr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(arr))
        _cache[k] = (arr, t.to(device))            # keeps `arr` alive so that its id cannot be recycled
    return _cache[k][1]


_images = {}


def _resident_image(images, i_img, device):
    """Device copy [H*W,3] fp32 of images[i_img], uploaded once.  Keyed by the container's identity (a reference to it is kept
    so the id cannot be recycled) and the image index."""
    key = (id(images), str(device))
    ent = _images.get(key)
    if ent is None or ent[0] is not images:
        if len(_images) > 8:
            _images.clear()
        ent = _images[key] = (images, {})
    dev_imgs = ent[1]
    t = dev_imgs.get(int(i_img))
    if t is None:
        img = images[i_img]
        img = img if isinstance(img, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(img))
        t = dev_imgs[int(i_img)] = img.to(device=device, dtype=torch.float32).reshape(-1, 3).contiguous()
    return t


def load_checkpoint(path, model, optimizer=None, map_location='cpu'):
    """main.py:111-117 (resume): model + optimizer state of a checkpoint written by train() below OR by the reference
    (train.py:105-114; its torch.optim.Adam state loads into trainer.FlatAdam); restores the sampling stream position when the
    checkpoint carries it.  Returns the checkpoint's idx."""
    from . import nerf_process
    ck = torch.load(path, map_location=map_location)
    model.load_state_dict(ck['model_state_dict'])
    if optimizer is not None and 'optimizer_state_dict' in ck:
        optimizer.load_state_dict(ck['optimizer_state_dict'])
    if 'nb_rng_counter' in ck:
        nerf_process.set_rng_state(ck['nb_rng_counter'])
    return ck.get('idx', 0)


def select_pixels(i, img_h, img_w, opts):
    """rays.py:40-54: indices of N_rays distinct pixels (centre crop while i < precrop_iters); same
    host RNG call as the reference so seeded runs pick the same pixels."""
    if i < opts.precrop_iters:
        dH = int(img_h // 2 * opts.precrop_frac)
        dW = int(img_w // 2 * opts.precrop_frac)
        rows = np.arange(img_h // 2 - dH, img_h // 2 + dH)
        cols = np.arange(img_w // 2 - dW, img_w // 2 + dW)
    else:
        rows, cols = np.arange(img_h), np.arange(img_w)
    sel = np.random.choice(a=rows.size * cols.size, size=opts.N_rays, replace=False)
    return (rows[sel // cols.size] * img_w + cols[sel % cols.size]).astype(np.int64)


def train(idx, i_train, images, gt_cam_param, hw, model, criterion, posenc, optimizer, global_batch_idx, vis, opts,
          dist_ctx=None):
    if not model.training:                      # train.py:14 (a full module-tree walk; skipped when already in training mode)
        model.train()
    apply_precision(model, opts)
    device = torch.device(f'cuda:{opts.gpu_ids[opts.rank]}')
    img_h, img_w = hw
    gt_intrinsic, gt_extrinsic = gt_cam_param
    llff = opts.data_type == 'llff'

    if global_batch_idx is not None and opts.global_batch:                       # train.py:25-32
        if hasattr(global_batch_idx, 'next_batch'):                              # device-native cursor (utils.GetterRayBatchIdx)
            rays_o, rays_d, target = global_batch_idx.next_batch(opts.N_rays)
        else:                                                                    # the reference's own getter object
            i_batch, rays_rgb, epoch = global_batch_idx(opts.N_rays)
            batch = rays_rgb[i_batch - opts.N_rays:i_batch]
            rays_o, rays_d, target = batch[:, 0].contiguous(), batch[:, 1].contiguous(), batch[:, 2].contiguous()
        if llff:
            rays_o, rays_d = ndc_rays(img_h, img_w, float(gt_intrinsic[0][0]), 1., rays_o, rays_d)
    else:                                                                         # train.py:35-45
        i_img = np.random.choice(i_train)
        if getattr(opts, 'cache_poses', True):
            pose = _device_copy(gt_extrinsic, device, 'poses')[i_img, :3, :4]
        else:       # per-step upload of the selected camera (the reference uploads ALL poses every step, train.py:20-21)
            ext = gt_extrinsic if isinstance(gt_extrinsic, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(gt_extrinsic))
            pose = ext[i_img, :3, :4].to(device=device, dtype=torch.float32, non_blocking=True)
        if getattr(opts, 'device_select', False):
            # selection on the device (keyed bijection, distinct pixels, no host permutation)
            region = None
            if idx < opts.precrop_iters:
                dH, dW = int(img_h // 2 * opts.precrop_frac), int(img_w // 2 * opts.precrop_frac)
                region = (img_h // 2 - dH, img_w // 2 - dW, 2 * dH, 2 * dW)
            from .engine import get_engine as _ge
            pix = _ge(device).select_pixels(opts.N_rays, img_h, img_w, region, seed=_seed(opts) + 7919 * i_img,
                                            offset=idx * opts.N_rays)
        else:
            pix = torch.from_numpy(select_pixels(idx, img_h, img_w, opts)).to(device, non_blocking=True)
        from .engine import get_engine
        if getattr(opts, 'cache_images', True):
            # SURVEY 8(f)-1, second half: the training images live on the device (uploaded the first time each one is used: 7.68 MB
            # per 800x800 image, 768 MB for 100 views); a step only gathers its N_rays target pixels from the resident stack.
            # (The reference re-uploads the whole image every step, train.py:37-38; opts.cache_images=False restores that, e.g.
            # for callers that modify `images` in place between steps.)
            target = get_engine(device).gather_rows(_resident_image(images, i_img, device), pix)
        else:
            img = images[i_img]
            img = img if isinstance(img, torch.Tensor) else torch.from_numpy(img)
            # H2D of the target image (train.py:37-38) and the gather of the selected pixels (rays.py:62) on a side stream:
            # the target is only needed at the first loss, so both overlap ray generation and the coarse forward
            main = torch.cuda.current_stream(device)
            side = _side_stream(device)
            pix_ready = torch.cuda.Event()
            pix_ready.record(main)
            with torch.cuda.stream(side):
                img_dev = img.to(device=device, dtype=torch.float32, non_blocking=True)
                side.wait_event(pix_ready)
                tgt = get_engine(device).gather_rows(img_dev.reshape(-1, 3), pix)
                copied = torch.cuda.Event()
                copied.record(side)
            tgt.record_stream(main)
            pix.record_stream(side)
            target = (tgt, copied)
        rays_o, rays_d = make_o_d_selected(img_w, img_h, gt_intrinsic, pose, pix, ndc=llff, near=1.)
    rays = torch.cat((rays_o, rays_d), dim=-1)

    fused = isinstance(criterion, torch.nn.MSELoss) or criterion is None
    if not fused and isinstance(target, tuple):
        torch.cuda.current_stream(device).wait_event(target[1])
        target = target[0]
    if fused:
        for net in (model.model_coarse, model.model_fine):
            net.bind_flat_grad()
        if not isinstance(optimizer, trainer.FlatAdam):
            optimizer.zero_grad(set_to_none=False)
        loss_buf = trainer.train_step(model, optimizer, rays, target, opts, dist_ctx=dist_ctx)
        loss_c, loss_f = loss_buf[0], loss_buf[1]
        loss = loss_c + loss_f
    else:                                                                         # generic criterion: autograd path
        saved = opts.data_type
        opts.data_type = 'blender'                                               # rays are already warped
        try:
            rgb_c, _, rgb_f, _ = batchify_rays_and_render_by_chunk(rays_o, rays_d, model, posenc, img_h, img_w, gt_intrinsic, opts)
        finally:
            opts.data_type = saved
        optimizer.zero_grad()
        loss_c = criterion(rgb_c, target)
        loss_f = criterion(rgb_f, target) if opts.N_samples_f > 0 else torch.zeros_like(loss_c)
        loss = loss_c + loss_f
        loss.backward()
        optimizer.step()

    if idx % opts.idx_print == 0:                                                # the only sync point (train.py:73-77)
        if opts.N_samples_f > 0:
            print('i : {} , Loss_C : {} , Loss_F : {} , Total_Loss : {} , PSNR_C : {} , PSNR_F : {}'.format(
                idx, float(loss_c), float(loss_f), float(loss), float(mse2psnr(loss_c.reshape(1))), float(mse2psnr(loss_f.reshape(1)))))
        else:
            print('i : {} , LOSS : {} , PSNR : {}'.format(idx, float(loss), float(mse2psnr(loss.reshape(1)))))

    if opts.idx_save and idx % opts.idx_save == 0 and idx > 0 and (dist_ctx is None or dist_ctx.rank == 0):   # train.py:105-114
        save_path = os.path.join(LOG_DIR, opts.exp_name)
        os.makedirs(save_path, exist_ok=True)
        from . import nerf_process
        checkpoint = {'idx': idx, 'model_state_dict': model.state_dict(), 'optimizer_state_dict': optimizer.state_dict(),
                      'nb_rng_counter': nerf_process.rng_state()}       # extra key (ignored by the reference's main.py:111-117)
        torch.save(checkpoint, os.path.join(save_path, opts.exp_name + '_{}.pth.tar'.format(idx)))
    return loss

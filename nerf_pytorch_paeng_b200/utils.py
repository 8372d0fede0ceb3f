"""Drop-in for the hot-path-adjacent helpers of the reference's utils.py (file:line refs into the
reference).  SSIM / LPIPS (utils.py:22-34) wrap third-party networks and are out of scope."""
import numpy as np
import torch


def to8b(x):
    """utils.py:11."""
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def img2mse(x, y):
    """utils.py:14."""
    return torch.mean((x - y) ** 2)


def mse2psnr(x):
    """utils.py:17-19: -10 log10(mse)."""
    return -10. * torch.log(x) / torch.log(torch.tensor([10.], device=x.device))


def put_epsilon(map):
    """utils.py:37."""
    return torch.max(1e-10 * torch.ones_like(map), map)


class GetterRayBatchIdx(object):
    """utils.py:41-58: cursor over the globally shuffled [N,3,3] ray/rgb table; reshuffles
    (torch.randperm) when an epoch is exhausted."""

    def __init__(self, rays_rgb):
        self.rays_rgb = rays_rgb
        self.epoch = 0
        self.i_batch = 0

    def shuffle_ray_idx(self, batch_size):
        print("Shuffle data after an epoch!")
        rand_idx = torch.randperm(self.rays_rgb.shape[0], device=self.rays_rgb.device)
        self.rays_rgb = self.rays_rgb[rand_idx]
        self.i_batch = batch_size
        self.epoch += 1

    def __call__(self, batch_size):
        self.i_batch += batch_size
        if self.i_batch >= self.rays_rgb.shape[0]:
            self.shuffle_ray_idx(batch_size)
        return self.i_batch, self.rays_rgb, self.epoch

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
    """Batch cursor over the global [N,3,3] (rays_o, rays_d, rgb) table of main.py:95-106 -- the reference's utils.py:41-58
    protocol (``i_batch, rays_rgb, epoch = getter(batch_size)``; the caller slices ``rays_rgb[i_batch-batch_size:i_batch]``; a
    torch.randperm reshuffle when an epoch is exhausted) plus a device-native fast path used by train.train:

    ``next_batch(batch_size)`` returns the same sequence of batches WITHOUT rewriting the table.  The reference re-materialises the
    whole table in shuffled order at every epoch boundary (``rays_rgb[rand_idx]``: a 2.3 GB gather for 100 views of 800x800);
    here the table stays put and only a composed int64 permutation is kept (perm_e = perm_{e-1}[randperm]), the batch rows being
    gathered by the nb_gather_rows kernel.  Both paths consume torch.randperm identically, so they can be mixed."""

    def __init__(self, rays_rgb):
        self.rays_rgb = rays_rgb
        self.epoch = 0
        self.i_batch = 0
        self._perm = None                   # None: identity (epoch 0 uses the table's own order, main.py:100 shuffled it on the host)
        self._materialised = True           # whether self.rays_rgb is physically in the current epoch's order

    def _advance(self, batch_size):
        """Cursor logic of utils.py:53-58; returns True when this call started a new epoch."""
        self.i_batch += batch_size
        if self.i_batch < self.rays_rgb.shape[0]:
            return False
        print("Shuffle data after an epoch!")
        n = self.rays_rgb.shape[0]
        rand_idx = torch.randperm(n, device=self.rays_rgb.device)
        self._perm = rand_idx if self._perm is None else self._perm[rand_idx]
        self._materialised = False
        self.i_batch = batch_size
        self.epoch += 1
        return True

    def shuffle_ray_idx(self, batch_size):
        """utils.py:47-51 as a public call: start a new epoch now."""
        self.i_batch = self.rays_rgb.shape[0] - batch_size
        self._advance(batch_size)

    def __call__(self, batch_size):
        """Reference protocol: the returned table is physically in epoch order (re-gathered lazily after a reshuffle)."""
        self._advance(batch_size)
        if not self._materialised:
            self.rays_rgb = self.rays_rgb[self._perm]
            self._perm = None
            self._materialised = True
        return self.i_batch, self.rays_rgb, self.epoch

    def next_batch(self, batch_size):
        """-> (rays_o, rays_d, target), each [batch_size, 3]: the rows the reference protocol would slice, gathered on the device."""
        self._advance(batch_size)
        lo, hi = self.i_batch - batch_size, self.i_batch
        if self._perm is None:
            rows = self.rays_rgb[lo:hi]
        else:
            from .engine import get_engine
            rows = get_engine(self.rays_rgb.device).gather_rows(self.rays_rgb.reshape(-1, 9), self._perm[lo:hi]).view(-1, 3, 3)
        return rows[:, 0].contiguous(), rows[:, 1].contiguous(), rows[:, 2].contiguous()

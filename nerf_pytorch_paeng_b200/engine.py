"""Per-GPU engine: owns the nb_handle, a growable workspace, and thin tensor->pointer wrappers.

PyTorch is used for device memory, streams and (elsewhere) torch.distributed only; every
computation below is a call into libnerf_b200.so on the current CUDA stream.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import NB_BF16, NB_FP32, MlpDesc, NBError, RenderCfg

_engines = {}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk32(t, name):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise NBError(f'{name}: expected a CUDA float32 tensor, got {t.dtype} on {t.device}')
    return t.contiguous()


class Engine:
    def __init__(self, device):
        if not torch.cuda.is_available():
            raise NBError('nerf_pytorch_paeng_b200 needs a CUDA (sm_100) device: there is no CPU fallback')
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise NBError(f'engine device must be cuda, got {self.device}')
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device('cuda', idx)
        h = C.c_void_p()
        rc = self.lib.nb_create(C.byref(h), idx, 0)
        if rc != 0:
            raise NBError(f'nb_create(device={idx}) failed with {rc} (needs an sm_100 GPU)')
        self.h = h
        self._ws = None
        self._fws = None
        info = (C.c_int32 * 4)()
        self.lib.nb_device_info(self.h, C.byref(info))
        self.sm_count = int(info[0])

    # ------------------------------------------------------------------ plumbing
    def _call(self, name, *args):
        rc = getattr(self.lib, name)(self.h, *args)
        if rc != 0:
            raise NBError(f'{name} failed ({rc}): {self.lib.nb_last_error(self.h).decode()}')

    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def launch_count(self):
        return int(self.lib.nb_launch_count(self.h))

    def workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.1) + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    # ------------------------------------------------------------------ K1
    def raygen(self, H, W, K, pose, pix_idx=None, ndc=False, ndc_focal=0., ndc_near=1.):
        """rays.py:20-34 (+ gather rays.py:59-60, + NDC nerf_process.py:8-28).  K: 3x3 (any host/torch
        array, used as doubles); pose: CUDA fp32 [3|4,4].  Returns rays_o, rays_d [N,3]."""
        fx, fy, cx, cy = float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2])
        pose = pose if (pose.dtype == torch.float32 and pose.stride(-1) == 1) else pose.float().contiguous()
        if pix_idx is None:
            n = H * W
        else:
            pix_idx = pix_idx.to(device=self.device, dtype=torch.int64).contiguous()
            n = pix_idx.numel()
        o = self.empty(n, 3)
        d = self.empty(n, 3)
        self._call('nb_raygen_pinhole', H, W, fx, fy, cx, cy, _ptr(pose), pose.stride(0), _ptr(pix_idx), n,
                   _ptr(o), _ptr(d), _lib.NB_RAYGEN_NDC if ndc else 0, float(ndc_focal), float(ndc_near), self.stream)
        return o, d

    def raygen_f64(self, H, W, K, pose):
        """rays.py:7-17 as NumPy >= 2 computes it (float64 throughout): rays_d [H*W,3] float64."""
        fx, fy, cx, cy = float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2])
        pose = pose if (pose.dtype == torch.float32 and pose.stride(-1) == 1) else pose.float().contiguous()
        d = self.empty(H * W, 3, dtype=torch.float64)
        self._call('nb_raygen_pinhole_f64', H, W, fx, fy, cx, cy, _ptr(pose), pose.stride(0), _ptr(d), self.stream)
        return d

    def ndc_rays(self, H, W, focal, near, rays_o, rays_d):
        o = _chk32(rays_o.reshape(-1, 3), 'rays_o')
        d = _chk32(rays_d.reshape(-1, 3), 'rays_d')
        oo, do = torch.empty_like(o), torch.empty_like(d)
        self._call('nb_ndc_rays', o.shape[0], H, W, float(focal), float(near), _ptr(o), _ptr(d), _ptr(oo), _ptr(do), self.stream)
        return oo, do

    def gather_rows(self, src, idx):
        src = _chk32(src, 'src')
        c = src.shape[-1]
        src2 = src.reshape(-1, c)
        idx = idx.to(device=self.device, dtype=torch.int64).contiguous()
        out = self.empty(idx.numel(), c)
        self._call('nb_gather_rows', idx.numel(), c, _ptr(idx), _ptr(src2), _ptr(out), self.stream)
        return out

    def select_pixels(self, n, H, W, region=None, seed=0, offset=0):
        """n distinct random flat pixel indices (int64, device) from region=(r0, c0, nr, nc) or the whole image."""
        r0, c0, nr, nc = region if region is not None else (0, 0, H, W)
        out = self.empty(n, dtype=torch.int64)
        self._call('nb_select_pixels', n, H, W, r0, c0, nr, nc, seed, offset, _ptr(out), self.stream)
        return out

    # ------------------------------------------------------------------ K2
    def stratified(self, n_rays, lower, span, t_rand=None, seed=0, offset=0):
        s_c = lower.numel()
        if t_rand is not None:
            t_rand = _chk32(t_rand, 't_rand')
            assert t_rand.shape == (n_rays, s_c)
        z = self.empty(n_rays, s_c)
        self._call('nb_stratified', n_rays, s_c, _ptr(lower), _ptr(span), _ptr(t_rand), seed, offset, None, _ptr(z), self.stream)
        return z

    def sample_pdf(self, z_c, weights_c, n_fine, u=None, seed=0, offset=0, cdf_in=None, bins_in=None, want_fine=True,
                   want_samples=False, want_inds=False, want_cdf=False, cdf_rows=0):
        """weights_c is the full [N,S_c] coarse weight tensor (the kernel applies the [...,1:-1] slice).
        cdf_rows: summation order of the pdf/cdf (nb_sample_pdf): 0 = torch's CUDA order for this call's N rows, k > 0 = for a
        call of k rows (the reference's chunk), -1 = fp64 accumulation (torch's CPU order)."""
        if z_c is not None:
            z_c = _chk32(z_c, 'z_vals')
            n, s_c = z_c.shape
        else:
            bins_in = _chk32(bins_in, 'bins')
            n, s_c = bins_in.shape[0], bins_in.shape[1] + 1
            want_fine = False
        if weights_c is not None:
            weights_c = _chk32(weights_c, 'weights')
            assert weights_c.shape == (n, s_c)
        if cdf_in is not None:
            cdf_in = _chk32(cdf_in, 'cdf')
        if bins_in is not None:
            bins_in = _chk32(bins_in, 'bins')
        if u is None:
            mode = 2
        else:
            u = _chk32(u, 'u')
            mode = 0 if u.dim() == 1 else 1
            assert u.shape[-1] == n_fine and (mode == 0 or u.shape[0] == n)
        z_f = self.empty(n, s_c + n_fine) if want_fine else None
        zs = self.empty(n, n_fine) if want_samples else None
        inds = self.empty(n, n_fine, dtype=torch.int64) if want_inds else None
        cdf = self.empty(n, s_c - 1) if want_cdf else None
        self._call('nb_sample_pdf', n, s_c, n_fine, _ptr(z_c), _ptr(weights_c), _ptr(u), mode, seed, offset, _ptr(cdf_in),
                   _ptr(bins_in), _ptr(z_f), _ptr(zs), _ptr(inds), _ptr(cdf), int(cdf_rows), None, self.stream)
        return z_f, zs, inds, cdf

    # ------------------------------------------------------------------ K3
    def posenc(self, x, L):
        x = _chk32(x, 'x')
        lead = x.shape[:-1]
        x2 = x.reshape(-1, 3)
        out = self.empty(x2.shape[0], 3 + 6 * L)
        self._call('nb_posenc', x2.shape[0], L, _ptr(x2), _ptr(out), self.stream)
        return out.reshape(*lead, 3 + 6 * L)

    def embed_points(self, rays, z, L_x, L_d):
        rays = _chk32(rays, 'rays')
        z = _chk32(z, 'z_vals')
        n, s = z.shape
        w = 6 + 6 * L_x + 6 * L_d
        out = self.empty(n * s, w)
        self._call('nb_embed_points', n, s, L_x, L_d, _ptr(rays), _ptr(z), _ptr(out), w, self.stream)
        return out

    # ------------------------------------------------------------------ K4
    def mlp_bytes(self, desc, n_pts, precision):
        act, wsf, wsb = C.c_size_t(), C.c_size_t(), C.c_size_t()
        self._call('nb_mlp_act_bytes', C.byref(desc), n_pts, precision, C.byref(act))
        self._call('nb_mlp_workspace_bytes', C.byref(desc), n_pts, precision, 0, C.byref(wsf))
        self._call('nb_mlp_workspace_bytes', C.byref(desc), n_pts, precision, 1, C.byref(wsb))
        return act.value, wsf.value, wsb.value

    def mlp_packed_bytes(self, desc):
        out = C.c_size_t()
        self._call('nb_mlp_packed_bytes', C.byref(desc), C.byref(out))
        return out.value

    def mlp_pack(self, desc, params, packed):
        self._call('nb_mlp_pack', C.byref(desc), _ptr(params), _ptr(packed), self.stream)

    def mlp_forward(self, desc, params, packed, precision, *, x=None, rays=None, z=None, save=False):
        """Returns (raw [P,4], act_save or None)."""
        if x is not None:
            x = _chk32(x, 'x')
            n_pts = x.shape[0]
        else:
            rays = _chk32(rays, 'rays')
            z = _chk32(z, 'z_vals')
            n_pts = z.numel()
        act_b, ws_f, _ = self.mlp_bytes(desc, n_pts, precision)
        act = torch.empty(act_b, dtype=torch.uint8, device=self.device) if save else None
        ws = self.workspace(ws_f)
        raw = self.empty(n_pts, 4)
        if x is not None:
            self._call('nb_mlp_forward_emb', C.byref(desc), _ptr(params), _ptr(packed), n_pts, _ptr(x), x.stride(0),
                       _ptr(raw), _ptr(act), precision, _ptr(ws), ws.numel(), self.stream)
        else:
            self._call('nb_mlp_forward_rays', C.byref(desc), _ptr(params), _ptr(packed), z.shape[0], z.shape[1],
                       _ptr(rays), _ptr(z), _ptr(raw), _ptr(act), precision, _ptr(ws), ws.numel(), self.stream)
        return raw, act

    def mlp_backward(self, desc, params, packed, precision, n_pts, act, d_raw, grad, accumulate=False, stage=None):
        """stage None: whole backward; 1 / 2: the two halves of nb_mlp_backward_stage (dgrad chain / wgrad)."""
        d_raw = _chk32(d_raw, 'd_raw')
        _, _, ws_b = self.mlp_bytes(desc, n_pts, precision)
        ws = self.workspace(ws_b)
        if stage is None:
            self._call('nb_mlp_backward', C.byref(desc), _ptr(params), _ptr(packed), n_pts, _ptr(act), _ptr(d_raw), _ptr(grad),
                       1 if accumulate else 0, precision, _ptr(ws), ws.numel(), self.stream)
        else:
            self._call('nb_mlp_backward_stage', C.byref(desc), _ptr(params), _ptr(packed), n_pts, _ptr(act), _ptr(d_raw), _ptr(grad),
                       1 if accumulate else 0, precision, _ptr(ws), ws.numel(), int(stage), self.stream)

    def mlp_tc_probe(self, desc, params, packed, rays, z, step):
        """Diagnostic: fp32 TMEM accumulators of chain step `step` ([P,256]) and raw [P,4]."""
        rays = _chk32(rays, 'rays')
        z = _chk32(z, 'z_vals')
        acc = torch.zeros(z.numel(), 256, device=self.device)
        raw = self.empty(z.numel(), 4)
        self._call('nb_mlp_tc_probe', C.byref(desc), _ptr(params), _ptr(packed), z.shape[0], z.shape[1], _ptr(rays), _ptr(z),
                   int(step), _ptr(acc), _ptr(raw), self.stream)
        return acc, raw

    # ------------------------------------------------------------------ K5
    def composite_forward(self, raw, z, rays_d, want_all=True):
        raw = _chk32(raw, 'outputs')
        z = _chk32(z, 'z_vals')
        rays_d = _chk32(rays_d, 'rays_d')
        n, s = z.shape
        rgb = self.empty(n, 3)
        disp = self.empty(n)
        acc = self.empty(n) if want_all else None
        w = self.empty(n, s) if want_all else None
        depth = self.empty(n) if want_all else None
        self._call('nb_composite_forward', n, s, _ptr(raw), _ptr(z), _ptr(rays_d), _ptr(rgb), _ptr(disp), _ptr(acc),
                   _ptr(w), _ptr(depth), self.stream)
        return rgb, disp, acc, w, depth

    def composite_backward(self, raw, z, rays_d, d_rgb):
        raw = _chk32(raw, 'outputs')
        d_rgb = _chk32(d_rgb, 'd_rgb')
        n, s = z.shape
        d_raw = torch.empty_like(raw)
        self._call('nb_composite_backward', n, s, _ptr(raw), _ptr(z), _ptr(rays_d), _ptr(d_rgb), _ptr(d_raw), self.stream)
        return d_raw

    # ------------------------------------------------------------------ fused drivers
    def _fused_ws(self, desc, n, cfg, train):
        need = C.c_size_t()
        self._call('nb_render_workspace_bytes', C.byref(desc), n, C.byref(cfg), 1 if train else 0, C.byref(need))
        if self._fws is None or self._fws.numel() < need.value:
            self._fws = None
            self._fws = torch.empty(need.value + 4096, dtype=torch.uint8, device=self.device)
        return self._fws

    @staticmethod
    def _u_mode(u):
        return 2 if u is None else (0 if u.dim() == 1 else 1)

    def render_rays(self, desc, nets, rays, lower, span, n_fine, precision, t_rand=None, u=None, seed=0, offset_c=0, offset_f=0, cdf_rows=0, ctr=None, exact_last=False):
        """nerf_process.py:185-216 as one nb_render_rays call.  nets = ((flat_c, packed_c), (flat_f, packed_f)).
        Returns (rgb_c, disp_c, rgb_f, disp_f); the fine pair is None when n_fine == 0."""
        rays = _chk32(rays, 'rays')
        n = rays.shape[0]
        cfg = RenderCfg(lower.numel(), max(int(n_fine), 0), precision, self._u_mode(u), seed, offset_c, offset_f, int(cdf_rows),
                        None if ctr is None else ctr.data_ptr(), 1 if exact_last else 0, 0)
        ws = self._fused_ws(desc, n, cfg, False)
        (pc, kc), (pf, kf) = nets
        rgb_c, disp_c = self.empty(n, 3), self.empty(n)
        rgb_f, disp_f = (self.empty(n, 3), self.empty(n)) if cfg.S_f > 0 else (None, None)
        self._call('nb_render_rays', C.byref(desc), C.byref(cfg), _ptr(pc), _ptr(kc), _ptr(pf), _ptr(kf), n, _ptr(rays), _ptr(lower),
                   _ptr(span), _ptr(t_rand), _ptr(u), _ptr(rgb_c), _ptr(disp_c), _ptr(rgb_f), _ptr(disp_f), _ptr(ws), ws.numel(),
                   self.stream)
        return rgb_c, disp_c, rgb_f, disp_f

    def train_rays(self, desc, nets, grads, rays, target, n_global, lower, span, n_fine, precision, loss_buf, out, which=3,
                   t_rand=None, u=None, seed=0, offset_c=0, offset_f=0, accumulate=False, target_ready=None, cdf_rows=0, ctr=None, exact_last=False):
        """train.py:53-69 minus the optimizer as nb_train_rays.  `out` is a dict that receives / supplies the
        rgb_c, disp_c, rgb_f, disp_f tensors (so a coarse call and a fine call can share it); which = nets bit mask."""
        rays = _chk32(rays, 'rays')
        target = _chk32(target, 'target')
        n = rays.shape[0]
        cfg = RenderCfg(lower.numel(), max(int(n_fine), 0), precision, self._u_mode(u), seed, offset_c, offset_f, int(cdf_rows),
                        None if ctr is None else ctr.data_ptr(), 1 if exact_last else 0, 0)
        ws = self._fused_ws(desc, n, cfg, True)
        (pc, kc), (pf, kf) = nets
        for tag, bit in (('c', 1), ('f', 2)):
            if which & bit and 'rgb_' + tag not in out:
                out['rgb_' + tag], out['disp_' + tag] = self.empty(n, 3), self.empty(n)
        ev = None if target_ready is None else C.c_void_p(target_ready.cuda_event)
        self._call('nb_train_rays', C.byref(desc), C.byref(cfg), _ptr(pc), _ptr(kc), _ptr(pf), _ptr(kf), n, _ptr(rays), _ptr(target),
                   ev, int(n_global), _ptr(lower), _ptr(span), _ptr(t_rand), _ptr(u), _ptr(grads[0]), _ptr(grads[1]),
                   1 if accumulate else 0, _ptr(loss_buf), _ptr(out.get('rgb_c')), _ptr(out.get('disp_c')), _ptr(out.get('rgb_f')),
                   _ptr(out.get('disp_f')), int(which), _ptr(ws), ws.numel(), self.stream)
        return out

    def frame_to8b(self, rgb, disp=None):
        """test.py:50-61: uint8 frame (and disparity normalised by its nanmax) on the device."""
        rgb = _chk32(rgb, 'rgb')
        n = rgb.shape[0] if rgb.dim() == 2 else rgb.numel() // 3
        rgb8 = torch.empty(rgb.shape, dtype=torch.uint8, device=self.device)
        disp8 = scratch = None
        if disp is not None:
            disp = _chk32(disp, 'disp')
            disp8 = torch.empty(disp.shape, dtype=torch.uint8, device=self.device)
            scratch = self.empty(1)
        self._call('nb_frame_to8b', n, _ptr(rgb), _ptr(disp), _ptr(scratch), _ptr(rgb8), _ptr(disp8), self.stream)
        return rgb8, disp8

    # ------------------------------------------------------------------ loss / optimiser
    def mse_grad(self, rgb, target, scale, loss_scale=0., loss_out=None, want_grad=True):
        rgb = _chk32(rgb, 'rgb')
        target = _chk32(target, 'target')
        d = torch.empty_like(rgb) if want_grad else None
        self._call('nb_mse_grad', rgb.shape[0], _ptr(rgb), _ptr(target), float(scale), float(loss_scale), _ptr(d),
                   _ptr(loss_out), self.stream)
        return d

    def adam_step_sum(self, p, g, srcs, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
        """Adam on g := sum(srcs) (rank-ordered list of equally shaped device tensors / pointers)."""
        arr = (C.c_void_p * len(srcs))(*[t if isinstance(t, int) else t.data_ptr() for t in srcs])
        self._call('nb_adam_step_sum', p.numel(), _ptr(p), _ptr(g), arr, len(srcs), _ptr(m), _ptr(v), float(lr), beta1, beta2, eps,
                   int(step), self.stream)

    def counter_add(self, ctr, delta):
        """ctr (device int64/uint64 scalar tensor) += delta, in stream order (graph-capturable)."""
        self._call('nb_counter_add', _ptr(ctr), int(delta), self.stream)

    def adam_step(self, p, g, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
        self._call('nb_adam_step', p.numel(), _ptr(p), _ptr(g), _ptr(m), _ptr(v), float(lr), beta1, beta2, eps, int(step),
                   self.stream)


def get_engine(device=None):
    """One engine (one nb_handle) per GPU per process."""
    if device is None:
        if not torch.cuda.is_available():
            raise NBError('nerf_pytorch_paeng_b200 needs a CUDA (sm_100) device: there is no CPU fallback')
        device = torch.device('cuda', torch.cuda.current_device())
    device = torch.device(device)
    if device.type != 'cuda':
        raise NBError(f'nerf_pytorch_paeng_b200 runs on CUDA only (got {device}); there is no CPU fallback')
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _engines:
        _engines[idx] = Engine(torch.device('cuda', idx))
    return _engines[idx]


__all__ = ['Engine', 'get_engine', 'NBError', 'MlpDesc', 'NB_FP32', 'NB_BF16']

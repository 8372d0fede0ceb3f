"""Fused train step / frame render drivers (the path bench.py measures).

``train_step`` is train.py:53-70 of the reference (render coarse+fine, MSE_c + MSE_f, backward,
Adam) without autograd: every stage is one call into libnerf_b200.so, gradients land in one flat
fp32 buffer per network (so data-parallel training is a single NCCL all-reduce over it), and Adam
runs over the flat parameter buffer.  ``render_frame`` is test.py:38-40 (make_o_d -> batchify).
"""
import os

import torch

from . import nerf_process as NP
from .engine import get_engine


class FlatAdam:
    """torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8) (main.py:79-80) over the model's two flat
    parameter buffers, one nb_adam_step launch per network."""

    def __init__(self, model, lr=5e-4, betas=(0.9, 0.999), eps=1e-8):
        self.model = model
        self.lr, self.betas, self.eps = lr, betas, eps
        self.step_count = 0
        self.state = {}
        self.param_groups = [{'lr': lr}]          # so reference-style LR schedulers can poke ['lr']

    def _nets(self):
        return (self.model.model_coarse, self.model.model_fine)

    def zero_grad(self, set_to_none=False):
        for net in self._nets():
            if net.flat_grad is not None:
                net.flat_grad.zero_()

    def step(self):
        self.step_count += 1
        lr = self.param_groups[0]['lr']
        for net in self._nets():
            flat = net.flat_params()
            grad = net.bind_flat_grad()
            eng = get_engine(flat.device)
            st = self.state.get(id(net))
            if st is None or st[0].shape != flat.shape or st[0].device != flat.device:
                st = (torch.zeros_like(flat), torch.zeros_like(flat))
                self.state[id(net)] = st
            eng.adam_step(flat, grad, st[0], st[1], lr, self.step_count, self.betas[0], self.betas[1], self.eps)
            net.mark_weights_changed()

    def state_dict(self):
        return {'step': self.step_count, 'lr': self.param_groups[0]['lr'],
                'state': [tuple(t.clone() for t in self.state[id(n)]) if id(n) in self.state else None for n in self._nets()]}

    def load_state_dict(self, sd):
        self.step_count = sd['step']
        self.param_groups[0]['lr'] = sd['lr']
        for n, st in zip(self._nets(), sd['state']):
            if st is not None:
                self.state[id(n)] = tuple(t.clone() for t in st)


MAX_POINTS_PER_PASS = 6 * 1024 * 1024      # activation stash + dY workspace ~ 10 KB/point (bf16) -> ~60 GB per pass at most


def render_losses_and_grads(model, rays, target, opts, n_global=None, want_grads=True, loss_buf=None, on_net_done=None):
    """Forward coarse+fine, loss, and (optionally) parameter gradients into the flat buffers.

    rays [N,6] (already NDC-warped for llff), target [N,3] -- or a callable returning it, evaluated right before the first
    loss so that e.g. a host->device copy of the target image can overlap the coarse forward.  n_global: total rays over all
    ranks (the MSE mean is over the GLOBAL batch so that summed rank gradients equal the single-GPU gradient).
    Batches whose point count would not fit the activation stash (BASELINE config 5: up to 64k rays x 256+512 samples) are
    processed in ray chunks with gradient accumulation.  Returns dict(rgb_c, rgb_f, disp_c, disp_f, loss_buf[2] device tensor).
    """
    eng = get_engine(rays.device)
    n = rays.shape[0]
    n_global = n if n_global is None else n_global
    rays = rays.contiguous()
    scale = 2.0 / (3.0 * n_global)
    if loss_buf is None:
        loss_buf = torch.zeros(2, device=rays.device)
    out = {'loss_buf': loss_buf}
    s_max = opts.N_samples_c + max(opts.N_samples_f, 0)
    max_pts = int(getattr(opts, 'max_points_per_pass', MAX_POINTS_PER_PASS))
    rays_per_pass = n if (not want_grads or n * s_max <= max_pts) else max(128, (max_pts // s_max) // 128 * 128)
    if (want_grads and rays_per_pass == n and n > 0 and getattr(opts, 'fused_driver', True)
            and model.model_coarse.precision == model.model_fine.precision):
        return _fused_losses_and_grads(eng, model, rays, target, opts, n_global, loss_buf, on_net_done, out)
    if isinstance(target, tuple):                 # (tensor, ready event) form of train.train
        torch.cuda.current_stream(rays.device).wait_event(target[1])
        target = target[0]
    tgt = None
    rng = getattr(opts, 'rng', None)
    parts = {k: [] for k in ('rgb_c', 'disp_c', 'rgb_f', 'disp_f')}
    nets_done = [False, False]
    for r0 in range(0, n, rays_per_pass):
        r1 = min(n, r0 + rays_per_pass)
        rays_k = rays if (r0 == 0 and r1 == n) else rays[r0:r1].contiguous()
        rays_d = rays_k[:, 3:].contiguous()
        if rng is not None and not (r0 == 0 and r1 == n):
            opts.rng = {k: (v[r0:r1] if v.dim() == 2 else v) for k, v in rng.items()}
        nk = r1 - r0
        z_prev = w_prev = None
        for i, fine in enumerate((False, True)):
            if fine and opts.N_samples_f <= 0:
                break
            net = model.model_fine if fine else model.model_coarse
            flat = net.flat_params()
            z = NP._fine_z(rays_k, opts, z_prev, w_prev) if fine else NP._coarse_z(rays_k, opts)
            raw, act = eng.mlp_forward(net.desc, flat, net.packed_weights(), net.precision, rays=rays_k, z=z, save=want_grads)
            raw3 = raw.view(nk, z.shape[1], 4)
            rgb, disp, acc, w, depth = eng.composite_forward(raw3, z, rays_d, want_all=not fine)
            tag = 'f' if fine else 'c'
            parts['rgb_' + tag].append(rgb)
            parts['disp_' + tag].append(disp)
            if tgt is None:
                tgt = target() if callable(target) else target
            tgt_k = tgt if (r0 == 0 and r1 == n) else tgt[r0:r1]
            d_rgb = eng.mse_grad(rgb, tgt_k, scale, 1.0 / (3.0 * n_global), loss_buf[i:i + 1], want_grad=want_grads)
            if want_grads:
                d_raw = eng.composite_backward(raw3, z, rays_d, d_rgb)
                grad = net.bind_flat_grad()
                eng.mlp_backward(net.desc, flat, net.packed_weights(), net.precision, nk * z.shape[1], act, d_raw.view(-1, 4), grad,
                                 accumulate=r0 > 0)
                del act
                if on_net_done is not None and r1 == n:
                    on_net_done(net)      # e.g. start this network's gradient all-reduce while the other network runs
            z_prev, w_prev = z, w
    if rng is not None:
        opts.rng = rng
    for k, v in parts.items():
        if v:
            out[k] = v[0] if len(v) == 1 else torch.cat(v, 0)
    return out


def _fused_losses_and_grads(eng, model, rays, target, opts, n_global, loss_buf, on_net_done, out):
    """Single-pass case of render_losses_and_grads as one nb_train_rays call (two when a hook wants to run between the
    coarse and the fine network, e.g. to start the coarse gradient all-reduce)."""
    n = rays.shape[0]
    ready = None
    if isinstance(target, tuple):
        target, ready = target
    elif callable(target):
        target = target()
    nc, nf = model.model_coarse, model.model_fine
    use_fine = opts.N_samples_f > 0
    nets = ((nc.flat_params(), nc.packed_weights()), (nf.flat_params(), nf.packed_weights()) if use_fine else (None, None))
    grads = (nc.bind_flat_grad(), nf.bind_flat_grad() if use_fine else None)
    args = NP._fused_sampling_args(n, opts, rays.device)
    passes = (3 if use_fine else 1,) if on_net_done is None or not use_fine else (1, 2)
    for which in passes:
        eng.train_rays(nc.desc, nets, grads, rays, target, n_global, precision=nc.precision, loss_buf=loss_buf, out=out,
                       which=which, target_ready=ready, **args)
        ready = None
        if on_net_done is not None:
            for net, bit in ((nc, 1), (nf, 2)):
                if which & bit:
                    on_net_done(net)
    return out


def train_step(model, optimizer, rays, target, opts, dist_ctx=None):
    """One optimisation step on a ray batch; returns the device tensor [loss_c, loss_f] (global means).

    Data parallel (dist_ctx): every rank normalises by the global ray count, the two flat gradient buffers and the two
    losses live in one joint buffer and are summed by a single all-reduce.  NB_DP_MODE=overlap restores the earlier
    schedule (coarse all-reduce launched before the fine pass, three collectives per step) for comparison."""
    n_global = rays.shape[0] * (dist_ctx.world_size if dist_ctx is not None else 1)
    if dist_ctx is None:
        out = render_losses_and_grads(model, rays, target, opts, n_global=n_global)
        optimizer.step()
        return out['loss_buf']
    mode = os.environ.get('NB_DP_MODE', 'joint')
    if mode == 'overlap':
        works = []
        out = render_losses_and_grads(model, rays, target, opts, n_global=n_global,
                                      on_net_done=lambda net: works.append(dist_ctx.allreduce_grad_async(net)))
        dist_ctx.allreduce_(out['loss_buf'])
        for w in works:
            w.wait()
        optimizer.step()
        return out['loss_buf']
    whole, loss = dist_ctx.joint_grad_buffer(model)
    loss.zero_()
    render_losses_and_grads(model, rays, target, opts, n_global=n_global, loss_buf=loss)
    if mode != 'none':               # 'none': no collective at all (timing experiments only; ranks diverge)
        dist_ctx.allreduce_(whole)
    optimizer.step()
    return loss


@torch.no_grad()
def render_frame(model, H, W, K, pose, opts, dist_ctx=None, chunk=None):
    """test.py:38-40: all H*W rays of one pose, coarse+fine.  Rays are generated on the device
    (ray-gen [+NDC] kernel), rendered in chunks of opts.chunk_rays; with a dist_ctx each rank renders
    a contiguous band of pixels and the bands are all-gathered.  Returns rgb [H*W,3], disp [H*W]."""
    eng = get_engine(pose.device)
    n = H * W
    lo, hi = (0, n) if dist_ctx is None else dist_ctx.shard_range(n)
    # chunk_rays is a memory knob of the reference (nerf_process.py:236); the fused kernels hold no [n_pts,90]
    # tensor, so frames are rendered in larger chunks unless the caller pins `chunk`
    chunk = chunk or max(int(opts.chunk_rays), 65536)
    rgb = eng.empty(hi - lo, 3)
    disp = eng.empty(hi - lo)
    ndc = opts.data_type == 'llff'
    use_fine = opts.N_samples_f > 0
    for s in range(lo, hi, chunk):
        e = min(hi, s + chunk)
        pix = torch.arange(s, e, device=pose.device, dtype=torch.int64)
        o, d = eng.raygen(H, W, K, pose, pix_idx=pix, ndc=ndc, ndc_focal=float(K[0][0]), ndc_near=1.)
        rays = torch.cat((o, d), dim=-1)
        out = render_rays_fused(model, rays, opts)
        rgb[s - lo:e - lo] = out['rgb_f' if use_fine else 'rgb_c']
        disp[s - lo:e - lo] = out['disp_f' if use_fine else 'disp_c']
    if dist_ctx is not None:
        rgb, disp = dist_ctx.gather_rows(rgb, n), dist_ctx.gather_rows(disp, n)
    return rgb, disp


def render_losses_and_grads_free(model, rays, opts):
    """Inference-only coarse+fine render of a ray chunk (no loss, no stash)."""
    eng = get_engine(rays.device)
    n = rays.shape[0]
    rays = rays.contiguous()
    rays_d = rays[:, 3:].contiguous()
    out = {}
    z_prev = w_prev = None
    for fine in (False, True):
        if fine and opts.N_samples_f <= 0:
            break
        net = model.model_fine if fine else model.model_coarse
        flat = net.flat_params()
        z = NP._fine_z(rays, opts, z_prev, w_prev) if fine else NP._coarse_z(rays, opts)
        raw, _ = eng.mlp_forward(net.desc, flat, net.packed_weights(), net.precision, rays=rays, z=z, save=False)
        rgb, disp, acc, w, depth = eng.composite_forward(raw.view(n, z.shape[1], 4), z, rays_d, want_all=not fine)
        tag = 'f' if fine else 'c'
        out['rgb_' + tag], out['disp_' + tag] = rgb, disp
        z_prev, w_prev = z, w
    return out


def render_rays_fused(model, rays, opts):
    """The same render as one nb_render_rays call (the route render_frame takes)."""
    eng = get_engine(rays.device)
    nc, nf = model.model_coarse, model.model_fine
    use_fine = opts.N_samples_f > 0
    nets = ((nc.flat_params(), nc.packed_weights()), (nf.flat_params(), nf.packed_weights()) if use_fine else (None, None))
    rgb_c, disp_c, rgb_f, disp_f = eng.render_rays(nc.desc, nets, rays, precision=nc.precision,
                                                   **NP._fused_sampling_args(rays.shape[0], opts, rays.device))
    out = {'rgb_c': rgb_c, 'disp_c': disp_c}
    if use_fine:
        out['rgb_f'], out['disp_f'] = rgb_f, disp_f
    return out

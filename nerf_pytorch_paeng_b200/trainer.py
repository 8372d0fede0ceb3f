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


class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam(model.parameters(), lr, betas=(0.9,0.999), eps=1e-8) (main.py:79-80) as ONE nb_adam_step launch per
    network over the model's two flat parameter / gradient buffers.

    A real torch.optim.Optimizer: `param_groups` holds model.parameters() (so the reference's scheduler.py:6
    CosineAnnealingWarmupRestarts(_LRScheduler) and main.py:82-90,161 work unchanged), and `state_dict()` / `load_state_dict()`
    speak torch.optim.Adam's format (per-parameter 'step', 'exp_avg', 'exp_avg_sq', parameters indexed in model.parameters()
    order) -- a checkpoint written by the reference (train.py:105-114) resumes here and vice versa (main.py:111-115).  The
    per-parameter moments are views of two flat buffers per network."""

    def __init__(self, model, lr=5e-4, betas=(0.9, 0.999), eps=1e-8):
        self.model = model
        self.step_count = 0
        self._flat_state = {}                     # id(net) -> (m, v) flat fp32 buffers
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None, capturable=False,
                        differentiable=False, fused=None)
        super().__init__(list(model.parameters()), defaults)

    # ---- convenience mirrors of the single param group
    @property
    def lr(self):
        return self.param_groups[0]['lr']

    @property
    def betas(self):
        return self.param_groups[0]['betas']

    @property
    def eps(self):
        return self.param_groups[0]['eps']

    def _nets(self):
        return (self.model.model_coarse, self.model.model_fine)

    def _moments(self, net):
        """Flat first/second-moment buffers of one network; self.state[p] holds views of them."""
        flat = net.flat_params()
        st = self._flat_state.get(id(net))
        if st is None or st[0].shape != flat.shape or st[0].device != flat.device:
            old = st
            st = (torch.zeros_like(flat), torch.zeros_like(flat))
            if old is not None and old[0].shape == flat.shape:          # the model moved to another device: keep the moments
                st[0].copy_(old[0])
                st[1].copy_(old[1])
            self._flat_state[id(net)] = st
            for p, (o, n, shape) in zip(net._plist, net.slices):
                ps = self.state[p]
                ps['exp_avg'] = st[0][o:o + n].view(shape)
                ps['exp_avg_sq'] = st[1][o:o + n].view(shape)
                ps.setdefault('step', torch.tensor(float(self.step_count)))
        return st

    @torch.no_grad()
    def step_summed(self, srcs, offsets):
        """Data-parallel step with the gradient sum over ranks folded into the update: srcs = rank-ordered joint gradient buffers
        [grad_coarse | grad_fine | losses] (distributed.PeerGradExchange.finish()), offsets = start of each network inside them."""
        self.step_count += 1
        g = self.param_groups[0]
        for net, off in zip(self._nets(), offsets):
            flat = net.flat_params()
            grad = net.bind_flat_grad()
            m, v = self._moments(net)
            ptrs = [t.data_ptr() + 4 * off for t in srcs]
            get_engine(flat.device).adam_step_sum(flat, grad, ptrs, m, v, g['lr'], self.step_count, g['betas'][0], g['betas'][1], g['eps'])
            net.mark_weights_changed()

    def zero_grad(self, set_to_none=False):
        """Zeroes the flat gradient buffers (p.grad stay views of them; set_to_none is ignored on purpose)."""
        for net in self._nets():
            if net.flat_grad is not None:
                net.flat_grad.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.step_count += 1
        g = self.param_groups[0]
        for net in self._nets():
            flat = net.flat_params()
            grad = net.bind_flat_grad()
            m, v = self._moments(net)
            get_engine(flat.device).adam_step(flat, grad, m, v, g['lr'], self.step_count, g['betas'][0], g['betas'][1], g['eps'])
            net.mark_weights_changed()
        return loss

    def state_dict(self):
        """torch.optim.Adam's layout: {'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]}."""
        if self.step_count > 0:
            for net in self._nets():
                self._moments(net)
            for ps in self.state.values():
                ps['step'] = torch.tensor(float(self.step_count))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """Accepts a torch.optim.Adam state_dict over model.parameters() (the reference's 'optimizer_state_dict',
        main.py:115) or one written by this class."""
        if 'param_groups' not in state_dict and 'step' in state_dict and 'state' in state_dict:     # round-1 format of this class
            self.step_count = int(state_dict['step'])
            self.param_groups[0]['lr'] = state_dict['lr']
            for net, st in zip(self._nets(), state_dict['state']):
                if st is not None:
                    m, v = self._moments(net)
                    m.copy_(st[0])
                    v.copy_(st[1])
            return
        super().load_state_dict(state_dict)               # per-parameter tensors, cast to the parameters' device / dtype
        steps = [float(ps['step']) for ps in self.state.values() if 'step' in ps]
        self.step_count = int(max(steps)) if steps else 0
        loaded = {p: dict(ps) for p, ps in self.state.items()}
        self._flat_state = {}
        for net in self._nets():
            net.flat_params()
            m, v = self._moments(net)                     # rebinds self.state[p] to views of fresh flat buffers
            for p, (o, n, shape) in zip(net._plist, net.slices):
                src = loaded.get(p)
                if src is not None and 'exp_avg' in src:
                    m[o:o + n].copy_(src['exp_avg'].reshape(-1))
                    v[o:o + n].copy_(src['exp_avg_sq'].reshape(-1))


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
    schedule (coarse all-reduce launched before the fine pass, three collectives per step) for comparison.  Default ('auto'): the
    copy-engine exchange of distributed.PeerGradExchange ('peer') for world sizes up to NB_DP_PEER_MAX_WORLD (2), else 'joint'."""
    n_global = rays.shape[0] * (dist_ctx.world_size if dist_ctx is not None else 1)
    if dist_ctx is None:
        out = render_losses_and_grads(model, rays, target, opts, n_global=n_global)
        optimizer.step()
        return out['loss_buf']
    mode = os.environ.get('NB_DP_MODE', 'auto')
    if mode == 'auto':        # copy-engine exchange when the optimizer can fold the sum (FlatAdam) and symmetric memory is available
        mode = 'joint'
        # measured (profiles/r02_scaling_*): at 2 GPUs the peer exchange wins (5.61 vs 5.69 ms/step); at 8 GPUs its 7 pushes per part
        # cost more than they hide (4.98 vs 4.92 ms/step with one NCCL all-reduce), so larger worlds keep the all-reduce
        if (isinstance(optimizer, FlatAdam) and rays.is_cuda and getattr(dist_ctx, '_peer_ok', True)
                and dist_ctx.world_size <= int(os.environ.get('NB_DP_PEER_MAX_WORLD', '2'))):
            try:
                if getattr(dist_ctx, '_peer_exchange', None) is None:
                    from .distributed import PeerGradExchange
                    dist_ctx._peer_exchange = PeerGradExchange(dist_ctx, model)
                mode = 'peer'
            except Exception as e:                       # no symmetric memory (e.g. gloo, no P2P): NCCL all-reduce of the joint buffer
                dist_ctx._peer_ok = False
                import warnings
                warnings.warn(f'peer gradient exchange unavailable ({type(e).__name__}: {e}); using one all-reduce of the joint buffer')
    if mode == 'overlap':
        works = []
        out = render_losses_and_grads(model, rays, target, opts, n_global=n_global,
                                      on_net_done=lambda net: works.append(dist_ctx.allreduce_grad_async(net)))
        dist_ctx.allreduce_(out['loss_buf'])
        for w in works:
            w.wait()
        optimizer.step()
        return out['loss_buf']
    if mode == 'peer' and isinstance(optimizer, FlatAdam):
        # copy-engine pushes under the kernels + sum folded into Adam (distributed.PeerGradExchange)
        px = getattr(dist_ctx, '_peer_exchange', None)
        if px is None:
            from .distributed import PeerGradExchange
            px = dist_ctx._peer_exchange = PeerGradExchange(dist_ctx, model)
        nc_sz = px.sizes[0]
        px.loss.zero_()

        def net_done(net):
            if net is model.model_coarse:
                px.push(0, nc_sz)                       # travels during the fine pass
            else:
                px.push(nc_sz, px.G)                    # fine gradient + the two losses
        render_losses_and_grads(model, rays, target, opts, n_global=n_global, loss_buf=px.loss, on_net_done=net_done)
        srcs = px.finish()
        loss_local = px.loss.clone()
        optimizer.step_summed(srcs, (0, nc_sz))
        total = loss_local
        for r, t in enumerate(srcs):
            if r != px.rank:
                total = total + t[px.G - 2:px.G]
        return total
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
    if model.model_fine.precision != NP.NB_BF16:
        # the fp32 parity path materialises every layer's [points, W] activations (~2.4 KB/point of workspace): keep a pass at <= 4M points
        chunk = min(chunk, max(int(opts.chunk_rays), (4 << 20) // (opts.N_samples_c + max(opts.N_samples_f, 0))))
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


class GraphedTrainStep:
    """render + MSE_c + MSE_f + backward of both networks (the nb_train_rays call: 14 kernel launches, 2 memsets,
    3 small copies) captured ONCE as a CUDA graph and replayed per step.  Nothing the graph needs changes on the host between
    replays: the ray batch and targets are copied into static buffers, the loss lands in a static buffer, and the Philox counters of
    the stratified / inverse-CDF draws are read from a device-resident counter that a captured nb_counter_add advances.  The Adam
    update (whose learning rate and bias corrections are host scalars owned by the caller's scheduler) and the bf16 weight re-pack
    stay ordinary launches after the replay.  With a dist_ctx the joint-buffer gradient all-reduce is part of the graph."""

    def __init__(self, model, opts, n_rays, device, dist_ctx=None):
        self.model, self.opts, self.n = model, opts, int(n_rays)
        self.eng = get_engine(device)
        dev = self.eng.device
        self.dist_ctx = dist_ctx
        self.n_global = self.n * (dist_ctx.world_size if dist_ctx is not None else 1)
        self.rays = torch.zeros(self.n, 6, device=dev)
        self.target = torch.zeros(self.n, 3, device=dev)
        self.whole = None
        if dist_ctx is not None:       # data parallel: gradients + losses in one joint buffer, ONE captured all-reduce (NCCL is graph-capturable)
            self.whole, self.loss = dist_ctx.joint_grad_buffer(model)
        else:
            self.loss = torch.zeros(2, device=dev)
        self.ctr = torch.zeros(1, dtype=torch.int64, device=dev)
        self.out = {}
        self.graph = None
        nc, nf = model.model_coarse, model.model_fine
        if nc.precision != nf.precision or opts.N_samples_f <= 0 or getattr(opts, 'rng', None) is not None:
            raise ValueError('GraphedTrainStep needs both networks in one precision, N_samples_f > 0 and in-kernel Philox draws')
        self._per_step = self.n * opts.N_samples_c // 4 + 1 + self.n * opts.N_samples_f // 4 + 1

    def _enqueue(self):
        nc, nf = self.model.model_coarse, self.model.model_fine
        nets = ((nc.flat, nc._packed), (nf.flat, nf._packed))
        grads = (nc.flat_grad, nf.flat_grad)
        args = NP._fused_sampling_args(self.n, self.opts, self.rays.device)
        args['offset_c'], args['offset_f'] = 0, self.n * self.opts.N_samples_c // 4 + 1     # relative to the device counter
        self.loss.zero_()
        self.eng.train_rays(nc.desc, nets, grads, self.rays, self.target, self.n_global, precision=nc.precision, loss_buf=self.loss, out=self.out,
                            which=3, ctr=self.ctr, **args)
        self.eng.counter_add(self.ctr, self._per_step)
        if self.dist_ctx is not None:
            self.dist_ctx.allreduce_(self.whole)

    def capture(self):
        nc, nf = self.model.model_coarse, self.model.model_fine
        for net in (nc, nf):
            net.flat_params()
            net.packed_weights()
            net.bind_flat_grad()
        side = torch.cuda.Stream(device=self.rays.device)
        side.wait_stream(torch.cuda.current_stream(self.rays.device))
        with torch.cuda.stream(side):           # warm-up outside capture: workspace growth, function attributes, output tensors
            for _ in range(2):
                self._enqueue()
        torch.cuda.current_stream(self.rays.device).wait_stream(side)
        torch.cuda.synchronize(self.rays.device)
        self._ptrs = (nc.flat.data_ptr(), nf.flat.data_ptr(), nc.flat_grad.data_ptr(), nf.flat_grad.data_ptr(), nc._packed.data_ptr(),
                      nf._packed.data_ptr(), self.eng._fws.data_ptr())
        self.graph = torch.cuda.CUDAGraph()
        k0 = self.eng.launch_count()
        with torch.cuda.graph(self.graph):
            self._enqueue()
        self.kernels_per_replay = self.eng.launch_count() - k0      # kernels of this library inside the graph
        return self

    def _still_valid(self):
        nc, nf = self.model.model_coarse, self.model.model_fine
        return (nc.flat is not None and nc.flat_grad is not None and nc._packed is not None and self.eng._fws is not None
                and self._ptrs == (nc.flat.data_ptr(), nf.flat.data_ptr(), nc.flat_grad.data_ptr(), nf.flat_grad.data_ptr(),
                                   nc._packed.data_ptr(), nf._packed.data_ptr(), self.eng._fws.data_ptr()))

    def __call__(self, optimizer, rays, target):
        """One optimisation step; returns the static device tensor [loss_c, loss_f] (overwritten by the next call)."""
        if self.graph is None or not self._still_valid():       # buffers were re-allocated (e.g. .to(), a larger workspace): re-capture
            self.capture()
        for net in (self.model.model_coarse, self.model.model_fine):
            net.packed_weights()                                 # re-pack after the previous optimizer step (ordinary launch)
        self.rays.copy_(rays, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        self.graph.replay()
        optimizer.step()
        return self.loss

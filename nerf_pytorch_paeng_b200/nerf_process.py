"""Drop-in for the reference's nerf_process.py (file:line refs are into the reference).

Same function names, argument orders and return conventions; bodies call libnerf_b200.so.
``render_rays`` / ``batchify_rays_and_render_by_chunk`` take the fused route (the [n_pts, 90]
embedding is only materialised by ``pre_process`` when that function is called directly).
RNG: the reference draws torch.rand on the device (nerf_process.py:55,162).  Here the draws come
from ``opts.rng`` when present -- a dict with injected tensors {'t_rand': [N,S_c], 'u': [N,S_f]}
(parity tests) -- else from the in-kernel Philox stream keyed by (opts.seed, call counter).
"""
import torch

from .engine import NB_BF16, get_engine
from .model.NeRF import NeRF

_zlin_cache = {}
_u_det_cache = {}
_counter = [0]


def _dev(opts, like=None):
    if like is not None and like.is_cuda:
        return like.device
    return torch.device(f'cuda:{opts.gpu_ids[opts.rank]}')


def _coarse_bins(opts, device):
    """lower / span of nerf_process.py:51-57, ray independent: computed once per (near, far, S_c)
    with the same torch ops as the reference (so t_vals is torch.linspace's, SURVEY B-6)."""
    key = (float(opts.near), float(opts.far), int(opts.N_samples_c), str(device))
    if key not in _zlin_cache:
        near = opts.near * torch.ones([1, 1], device=device)
        far = opts.far * torch.ones([1, 1], device=device)
        t_vals = torch.linspace(0., 1., steps=opts.N_samples_c, device=device)
        z_vals = near * (1. - t_vals) + far * (t_vals)
        mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat([mids, z_vals[..., -1:]], -1)
        lower = torch.cat([z_vals[..., :1], mids], -1)
        _zlin_cache[key] = (lower.reshape(-1).contiguous(), (upper - lower).reshape(-1).contiguous())
    return _zlin_cache[key]


def _next_offset(n):
    off = _counter[0]
    _counter[0] += int(n)
    return off


def rng_state():
    """Philox counter of this process' sampling stream (saved in checkpoints by train.train so that a resumed run does not
    replay the draws of the first steps)."""
    return int(_counter[0])


def set_rng_state(v):
    _counter[0] = int(v)


def _seed(opts):
    """Philox key: opts.seed decorrelated per data-parallel rank (every rank of train() gets the same opts.seed, and ray i of
    every rank would otherwise see identical stratified / pdf draws)."""
    seed = int(getattr(opts, 'seed', 0))
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        seed += 0x9E3779B1 * torch.distributed.get_rank()
    return seed & 0xFFFFFFFFFFFFFFFF


def apply_precision(model, opts):
    """opts.precision ('bf16' | 'fp32', config.py --precision, default bf16) selects the MLP path of the drop-in entry points;
    hand-built opts without the attribute leave the module's own setting alone."""
    prec = getattr(opts, 'precision', None)
    if prec is not None and isinstance(model, NeRF):
        model.set_precision(prec)


def _cdf_rows(opts, n):
    """Summation order of the fine pdf/cdf (nb_sample_pdf's cdf_rows).  Default: torch's CUDA order for the call the REFERENCE
    would make -- n rows, or opts.chunk_rays when this engine is handed a larger chunk than the reference's batchify loop uses
    (nerf_process.py:236) -- so bin indices are bit-identical to the reference on the same device.  opts.cdf_order = 'fp64'
    selects fp64 accumulation (torch's CPU cumsum; what the CPU-generated fixtures hold)."""
    if getattr(opts, 'cdf_order', 'cuda') == 'fp64':
        return -1
    chunk = int(getattr(opts, 'chunk_rays', 0) or 0)
    return min(int(n), chunk) if chunk > 0 else int(n)


def _injected(opts, name):
    rng = getattr(opts, 'rng', None)
    return None if rng is None else rng.get(name)


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """nerf_process.py:8-28."""
    return get_engine(rays_o.device).ndc_rays(H, W, float(focal), float(near), rays_o, rays_d)


def _coarse_z(rays, opts):
    eng = get_engine(rays.device)
    lower, span = _coarse_bins(opts, rays.device)
    n = rays.shape[0]
    t_rand = _injected(opts, 't_rand')
    return eng.stratified(n, lower, span, t_rand=t_rand, seed=_seed(opts),
                          offset=_next_offset(n * opts.N_samples_c // 4 + 1))


def _fine_z(rays, opts, z_vals, weights):
    eng = get_engine(rays.device)
    n = rays.shape[0]
    if opts.perturb == 0.:                                            # det=True, nerf_process.py:65,158-161
        key = (int(opts.N_samples_f), str(rays.device))
        if key not in _u_det_cache:
            _u_det_cache[key] = torch.linspace(0., 1., steps=opts.N_samples_f, device=rays.device)
        u = _u_det_cache[key]
    else:
        u = _injected(opts, 'u')
    z_fine, _, _, _ = eng.sample_pdf(z_vals, weights.detach(), opts.N_samples_f, u=u, seed=_seed(opts),
                                     offset=_next_offset(n * opts.N_samples_f // 4 + 1), cdf_rows=_cdf_rows(opts, n))
    return z_fine


def _fused_sampling_args(n, opts, device):
    """Arguments of the fused drivers (nb_render_rays / nb_train_rays) that reproduce _coarse_z followed by _fine_z:
    same injected draws, same Philox counters in the same order."""
    lower, span = _coarse_bins(opts, device)
    n_fine = max(int(opts.N_samples_f), 0)
    off_c = _next_offset(n * opts.N_samples_c // 4 + 1)
    u = None
    off_f = 0
    if n_fine > 0:
        if opts.perturb == 0.:
            key = (n_fine, str(device))
            if key not in _u_det_cache:
                _u_det_cache[key] = torch.linspace(0., 1., steps=n_fine, device=device)
            u = _u_det_cache[key]
        else:
            u = _injected(opts, 'u')
        off_f = _next_offset(n * n_fine // 4 + 1)
    return dict(lower=lower, span=span, n_fine=n_fine, t_rand=_injected(opts, 't_rand'), u=u, seed=_seed(opts),
                offset_c=off_c, offset_f=off_f, cdf_rows=_cdf_rows(opts, n), exact_last=bool(getattr(opts, 'exact_last_sample', False)))


def pre_process(rays, posenc, opts, z_vals=None, weights=None, isFine=False):
    """nerf_process.py:32-85.  Returns (embedded [n_pts, 90], z_vals, rays_d)."""
    fn_posenc, fn_posenc_d = posenc
    rays = rays.contiguous()
    z_vals = _fine_z(rays, opts, z_vals, weights) if isFine else _coarse_z(rays, opts)
    L_x = getattr(fn_posenc, 'L', None)
    L_d = getattr(fn_posenc_d, 'L', None)
    if L_x is not None and L_d is not None:
        embedded = get_engine(rays.device).embed_points(rays, z_vals, L_x, L_d)
    else:  # foreign encoders: compose exactly as the reference does (nerf_process.py:69-84)
        rays_o, rays_d = rays[:, :3], rays[:, 3:]
        viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
        pts = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * z_vals.unsqueeze(-1)
        embedded = torch.cat([fn_posenc(pts.view(-1, 3)),
                              fn_posenc_d(viewdirs.unsqueeze(1).expand(pts.size()).reshape(-1, 3))], -1)
    return embedded, z_vals, rays[:, 3:]


class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, z_vals, rays_d):
        eng = get_engine(outputs.device)
        outputs = outputs.contiguous()
        z_vals = z_vals.contiguous()
        rays_d = rays_d.contiguous()
        rgb, disp, acc, w, depth = eng.composite_forward(outputs, z_vals, rays_d)
        ctx.save_for_backward(outputs, z_vals, rays_d)
        ctx.mark_non_differentiable(disp, acc, w, depth)   # only rgb_map feeds the loss (train.py:60-66)
        return rgb, disp, acc, w, depth

    @staticmethod
    def backward(ctx, d_rgb, *unused):
        outputs, z_vals, rays_d = ctx.saved_tensors
        d_raw = get_engine(outputs.device).composite_backward(outputs, z_vals, rays_d, d_rgb.contiguous())
        return d_raw, None, None


def post_process(outputs, z_vals, rays_d):
    """nerf_process.py:89-140.  outputs [N,S,4], z_vals [N,S], rays_d [N,3] ->
    (rgb_map, disp_map, acc_map, weights, depth_map).  Differentiable wrt outputs through rgb_map."""
    return _Composite.apply(outputs, z_vals, rays_d)


def sample_pdf(bins, weights, N_samples, det=False, opts=None):
    """nerf_process.py:144-182.  bins [N,M], weights [N,M-1] -> samples [N,N_samples] (unsorted).
    Same kernel as the fused fine path, fed the bin positions directly (bins_in)."""
    assert opts is not None
    eng = get_engine(bins.device)
    n, m = bins.shape
    w_pad = torch.zeros(n, m + 1, device=bins.device)      # the kernel slices [...,1:-1] itself
    w_pad[:, 1:-1] = weights
    if det:
        u = torch.linspace(0., 1., steps=N_samples, device=bins.device)
    else:
        u = _injected(opts, 'u')
    _, samples, _, _ = eng.sample_pdf(None, w_pad, N_samples, u=u, bins_in=bins, want_samples=True,
                                      seed=_seed(opts), offset=_next_offset(n * N_samples // 4 + 1),
                                      cdf_rows=-1 if getattr(opts, 'cdf_order', 'cuda') == 'fp64' else n)
    return samples


def _net(model, fine):
    return model.model_fine if fine else model.model_coarse


def render_rays(rays, model, posenc, opts):
    """nerf_process.py:185-216.  rays [N,6] -> dict(rgb_c, disp_c[, rgb_f, disp_f])."""
    if not isinstance(model, NeRF):
        raise TypeError('render_rays needs the nerf_pytorch_paeng_b200.model.NeRF module (CUDA path only)')
    apply_precision(model, opts)
    rays = rays.contiguous()
    rays_d = rays[:, 3:].contiguous()
    # 1-a/2-a) coarse depths, fused points+PE+MLP (chunk_pts is a memory knob of the reference; the
    # fused kernels do not need it)
    z_vals = _coarse_z(rays, opts)
    raw = _net(model, False).forward_rays(rays, z_vals).view(z_vals.shape[0], z_vals.shape[1], 4)
    rgb_map, disp_map, acc_map, weights, depth_map = post_process(raw, z_vals, rays_d)
    if opts.N_samples_f > 0:
        z_fine = _fine_z(rays, opts, z_vals, weights)
        raw_f = _net(model, True).forward_rays(rays, z_fine).view(z_fine.shape[0], z_fine.shape[1], 4)
        rgb_f, disp_f, _, _, _ = post_process(raw_f, z_fine, rays_d)
        return {'rgb_c': rgb_map, 'disp_c': disp_map, 'rgb_f': rgb_f, 'disp_f': disp_f}
    return {'rgb_c': rgb_map, 'disp_c': disp_map}


def batchify_rays_and_render_by_chunk(ray_o, ray_d, model, posenc, H, W, K, opts):
    """nerf_process.py:220-252.  Accepts the stride-0 expanded rays_o that make_o_d returns."""
    flat_ray_o, flat_ray_d = ray_o.reshape(-1, 3), ray_d.reshape(-1, 3)
    if opts.data_type == 'llff':
        flat_ray_o, flat_ray_d = ndc_rays(H, W, float(K[0][0]), 1., flat_ray_o.contiguous(), flat_ray_d.contiguous())
    N_rays = flat_ray_o.size(0)
    rays = torch.cat((flat_ray_o, flat_ray_d), dim=-1)
    rng = getattr(opts, 'rng', None)
    outs = {'rgb_c': [], 'disp_c': [], 'rgb_f': [], 'disp_f': []}
    for i in range(0, N_rays, opts.chunk_rays):
        if rng is not None:   # slice injected draws per chunk
            opts.rng = {k: (v[i:i + opts.chunk_rays] if v.dim() == 2 else v) for k, v in rng.items()}
        d = render_rays(rays[i:i + opts.chunk_rays], model, posenc, opts)
        for k, v in d.items():
            outs[k].append(v)
    if rng is not None:
        opts.rng = rng
    cat = {k: (torch.cat(v, dim=0) if len(v) > 1 else (v[0] if v else None)) for k, v in outs.items()}
    if opts.N_samples_f > 0:
        return cat['rgb_c'], cat['disp_c'], cat['rgb_f'], cat['disp_f']
    return cat['rgb_c'], cat['disp_c'], None, None

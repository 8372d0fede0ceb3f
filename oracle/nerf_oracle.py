"""CPU oracle for the NeRF ray-batch hot path.  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference algorithm
(nuggy875/NeRF_pytorch_paeng).  It is NOT part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import it, and there only as the checker or as
the timed CPU baseline.  The product path (``nerf_pytorch_paeng_b200``) never
imports it and has no CPU fallback.

Parity status: PINNED.  The reference ships no tests/golden vectors of its own
(SURVEY.md section 4), so the pin is manufactured: ``oracle/make_golden.py``
imports the unmodified reference from ``/root/reference`` (CPU, with the
``IQA_pytorch`` stub and the two device patches described in SURVEY.md 8(c)),
runs every hot-path function on seeded inputs with injected RNG, and stores
inputs+outputs under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks
every function below against those vectors.

Rounding model: every fp32 operation of the reference is one PyTorch eager op,
i.e. individually rounded, no FMA contraction.  numpy float32 arithmetic has
the same property.  The two places where the reference's CPU result is an FMA
chain (the K=3 matmul of ``make_o_d``) are emulated in float64 (products of two
fp32 numbers are exact in fp64).

Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

F32 = np.float32


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def _fma32(a, b, c):
    """fp32 fma(a, b, c) emulated in fp64: a*b is exact, one add, round to fp32.

    (Double rounding can differ from a true fma with probability ~2^-29 per op.)
    """
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# K1: ray generation
# ----------------------------------------------------------------------------------------------
def make_o_d(img_w, img_h, img_k, pose):
    """rays.py:20-34.  Pinhole rays for a full image (torch variant).

    i = column, j = row (linspace with step 1, meshgrid('ij') on (W,H) then .t()).
    dirs = [(i-cx)/fx, -(j-cy)/fy, -1] with K demoted from f64 to fp32 scalars.
    rays_d = dirs @ R^T: on CPU/MKL this is the k-ordered FMA chain
    fma(z,R[k,2], fma(y,R[k,1], x*R[k,0])).
    rays_o = pose[:3,3] broadcast.
    Returns (rays_o[H,W,3], rays_d[H,W,3]) float32.
    """
    pose = _f32(pose)
    cx, cy = F32(img_k[0][2]), F32(img_k[1][2])
    fx, fy = F32(img_k[0][0]), F32(img_k[1][1])
    i = np.broadcast_to(np.arange(img_w, dtype=np.float32)[None, :], (img_h, img_w))
    j = np.broadcast_to(np.arange(img_h, dtype=np.float32)[:, None], (img_h, img_w))
    x = (i - cx) / fx
    y = -((j - cy) / fy)
    z = -np.ones_like(x)
    R = pose[:3, :3]
    d = np.empty((img_h, img_w, 3), dtype=np.float32)
    for k in range(3):
        acc = x * R[k, 0]
        acc = _fma32(y, np.broadcast_to(R[k, 1], y.shape), acc)
        acc = _fma32(z, np.broadcast_to(R[k, 2], z.shape), acc)
        d[..., k] = acc
    o = np.broadcast_to(pose[:3, 3], d.shape).copy()
    return o, d


def get_rays_np(H, W, K, c2w):
    """rays.py:7-17.  NumPy variant used for the global-batch precompute.

    Under NumPy >= 2 (NEP 50) K's float64 entries promote dirs / rays_d to
    float64; products then a left-to-right np.sum over 3, no FMA.
    The reference casts the stacked result to fp32 at main.py:101; we return
    the same dtypes the reference returns (rays_d float64, rays_o c2w's dtype).
    """
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    dirs = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], -1)
    prod = dirs[..., np.newaxis, :] * c2w[:3, :3]
    rays_d = (prod[..., 0] + prod[..., 1]) + prod[..., 2]
    rays_o = np.broadcast_to(c2w[:3, -1], np.shape(rays_d))
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """nerf_process.py:8-28.  LLFF NDC warp, fp32, c_w/c_h evaluated in double."""
    o = _f32(rays_o)
    d = _f32(rays_d)
    near32 = F32(near)
    t = -(near32 + o[..., 2]) / d[..., 2]
    o = o + t[..., None] * d
    c_w = F32(-1. / (W / (2. * float(focal))))
    c_h = F32(-1. / (H / (2. * float(focal))))
    two_near = F32(2. * near)
    o0 = c_w * o[..., 0] / o[..., 2]
    o1 = c_h * o[..., 1] / o[..., 2]
    o2 = F32(1.) + two_near / o[..., 2]
    d0 = c_w * (d[..., 0] / d[..., 2] - o[..., 0] / o[..., 2])
    d1 = c_h * (d[..., 1] / d[..., 2] - o[..., 1] / o[..., 2])
    d2 = F32(-2. * near) / o[..., 2]
    return np.stack([o0, o1, o2], -1).astype(np.float32), np.stack([d0, d1, d2], -1).astype(np.float32)


def select_rays(rays_o, rays_d, target_img, selected_idx, img_w):
    """rays.py:54-62 with the random index vector injected (no precrop):
    coords = (row, col) = divmod(idx, W); gather o, d, rgb."""
    r = selected_idx // img_w
    c = selected_idx % img_w
    return rays_o[r, c], rays_d[r, c], target_img[r, c]


# ----------------------------------------------------------------------------------------------
# K3: positional encoding
# ----------------------------------------------------------------------------------------------
def positional_encoding(x, L):
    """model/PositionalEncoding.py:7-30.  [x, sin(2^k x), cos(2^k x)]_{k<L}; blocks are 3 wide."""
    x = _f32(x)
    outs = [x]
    for k in range(L):
        f = F32(2.0 ** k)
        xf = x * f
        outs.append(np.sin(xf))
        outs.append(np.cos(xf))
    return np.concatenate(outs, -1).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# K2: sampling
# ----------------------------------------------------------------------------------------------
def torch_linspace01(steps):
    """torch.linspace(0., 1., steps) fp32.

    ATen (RangeFactories) computes step = (end-start)/(steps-1) in fp32 and fills
    symmetrically: i < steps/2: start + step*i, else end - step*(steps-1-i), the
    latter contracted to an FMA by the compiler (verified bit-exact against
    torch 2.11 CPU for steps in {5,33,64,128,192}).  NOT s/(S-1) (SURVEY B-6).
    """
    step = F32(1.0) / F32(steps - 1)
    idx = np.arange(steps)
    lo = (np.float64(step) * idx).astype(np.float32)
    hi = (1.0 - np.float64(step) * (steps - 1 - idx)).astype(np.float32)
    return np.where(idx < steps // 2, lo, hi).astype(np.float32)


def stratified_z(near, far, n_samples, t_rand, t_vals=None):
    """nerf_process.py:43-60.  Coarse depths; jitter is unconditional (SURVEY B-4).

    t_rand: [N, S_c] injected uniform numbers (the reference draws torch.rand).
    """
    t = torch_linspace01(n_samples) if t_vals is None else _f32(t_vals)
    z_lin = F32(near) * (F32(1.) - t) + F32(far) * t
    mids = F32(.5) * (z_lin[1:] + z_lin[:-1])
    upper = np.concatenate([mids, z_lin[-1:]])
    lower = np.concatenate([z_lin[:1], mids])
    t_rand = _f32(t_rand)
    return (lower[None, :] + (upper - lower)[None, :] * t_rand).astype(np.float32)


def aten_cuda_scan_threads_x(num_rows, row_size):
    """ATen ScanUtils.cuh get_log_num_threads_x_inner_scan<uint32_t> (torch 2.x): threads per row of the innermost-dim scan
    kernel, INCLUDING its unsigned wrap-around when num_rows >> row_size."""
    lx = ly = 0
    while (1 << lx) < row_size:
        lx += 1
    while (1 << ly) < num_rows:
        ly += 1
    diff = (lx - ly) & 0xFFFFFFFF
    lx = ((9 + diff) & 0xFFFFFFFF) // 2
    return 1 << min(max(4, lx), 9)


def aten_cuda_sum_lastdim(w):
    """torch.sum(w, -1, keepdim=True) of a contiguous fp32 [N, n] tensor on CUDA for n < 128 (ATen Reduce.cuh, no input
    vectorisation): block_width bw = min(2^floor(log2 n), 32) lanes; lane x owns elements x, x+bw, x+2bw, x+3bw in four
    accumulators combined as ((v0+v1)+v2)+v3; then block_x_reduce: p[x] += p[x+off] for off = bw/2 .. 1."""
    w = _f32(w)
    n = w.shape[-1]
    assert 1 <= n < 128
    bw = 1
    while bw * 2 <= n and bw < 32:
        bw *= 2
    pad = np.zeros(w.shape[:-1] + (4 * bw,), np.float32)
    pad[..., :n] = w
    v = pad.reshape(w.shape[:-1] + (4, bw))
    p = (((v[..., 0, :] + v[..., 1, :]).astype(np.float32) + v[..., 2, :]).astype(np.float32) + v[..., 3, :]).astype(np.float32)
    off = bw // 2
    while off > 0:
        p = p.copy()
        p[..., :off] = (p[..., :off] + p[..., off:2 * off]).astype(np.float32)
        off //= 2
    return p[..., :1]


def aten_cuda_cumsum_lastdim(x, num_rows=None):
    """torch.cumsum(x, -1) of a contiguous fp32 [N, n] tensor on CUDA (ATen ScanUtils.cuh tensor_kernel_scan_innermost_dim):
    Sklansky scan over blocks of 2*ntx elements, the running total added to the first element of the next block."""
    x = _f32(x)
    N, n = x.shape
    ntx = aten_cuda_scan_threads_x(N if num_rows is None else num_rows, n)
    B = 2 * ntx
    out = np.empty_like(x)
    total = np.zeros(N, np.float32)
    lg = ntx.bit_length() - 1
    for c0 in range(0, n, B):
        m_len = min(B, n - c0)
        width = 1
        while width < m_len:
            width *= 2
        buf = np.zeros((N, max(width, 2)), np.float32)          # elements past the block end never feed kept outputs
        buf[:, :m_len] = x[:, c0:c0 + m_len]
        buf[:, 0] = (buf[:, 0] + total).astype(np.float32)
        for mm in range(lg + 1):
            sft = 1 << mm
            if sft >= buf.shape[1]:
                break
            new = buf.copy()
            for t in range(min(ntx, buf.shape[1] // 2)):
                a = ((t >> mm) << (mm + 1)) | sft
                ti = a + (t % sft)
                if ti < buf.shape[1]:
                    new[:, ti] = (buf[:, ti] + buf[:, a - 1]).astype(np.float32)
            buf = new
        out[:, c0:c0 + m_len] = buf[:, :m_len]
        if m_len == B:
            total = buf[:, B - 1].copy()
    return out


def pdf_to_cdf(weights, order='cpu', rows=None):
    """nerf_process.py:150-155.  weights[N,M] (already sliced [...,1:-1]) -> cdf[N,M+1].

    Summation order is device specific in the reference (SURVEY B-5):
      order='cpu'  : row sum accumulated in float64 then rounded to fp32; cumsum accumulated in float64, each output rounded
                     to fp32 (the latter is exactly ATen's CPU cumsum) -- the order the CPU-generated fixtures are closest to;
      order='cuda' : the exact fp32 order of torch.sum / torch.cumsum on CUDA for a call with `rows` rows (default N), pinned
                     bit for bit against the reference running on a B200 (tests/golden/sample_pdf_cuda.npz).
    """
    w = (_f32(weights) + F32(1e-5)).astype(np.float32)
    if order == 'cuda':
        s = aten_cuda_sum_lastdim(w)
        pdf = (w / s).astype(np.float32)
        cdf = aten_cuda_cumsum_lastdim(pdf, rows)
        return np.concatenate([np.zeros_like(cdf[..., :1]), cdf], -1)
    s = w.astype(np.float64).sum(-1, keepdims=True).astype(np.float32)
    pdf = (w / s).astype(np.float32)
    cdf = np.cumsum(pdf.astype(np.float64), -1).astype(np.float32)
    return np.concatenate([np.zeros_like(cdf[..., :1]), cdf], -1)


def invert_cdf(bins, cdf, u):
    """nerf_process.py:166-182.  u[N,S_f] -> (samples[N,S_f], inds[N,S_f] int64).

    inds = searchsorted(cdf, u, right=True) = first index with cdf > u.
    """
    bins = _f32(bins)
    cdf = _f32(cdf)
    u = _f32(u)
    n_knots = cdf.shape[-1]
    inds = (cdf[:, None, :] <= u[:, :, None]).sum(-1).astype(np.int64)
    below = np.maximum(0, inds - 1)
    above = np.minimum(n_knots - 1, inds)
    cdf_b = np.take_along_axis(cdf, below, -1)
    cdf_a = np.take_along_axis(cdf, above, -1)
    bins_b = np.take_along_axis(bins, below, -1)
    bins_a = np.take_along_axis(bins, above, -1)
    denom = cdf_a - cdf_b
    denom = np.where(denom < F32(1e-5), F32(1.), denom)
    t = (u - cdf_b) / denom
    samples = bins_b + t * (bins_a - bins_b)
    return samples.astype(np.float32), inds


def sample_pdf(bins, weights, u, order='cpu', rows=None):
    """nerf_process.py:144-182 with u injected (det: torch.linspace(0,1,S_f) expanded)."""
    cdf = pdf_to_cdf(weights, order, rows)
    u = np.broadcast_to(_f32(u), (cdf.shape[0], np.shape(u)[-1]))
    return invert_cdf(bins, cdf, u)


def fine_z(z_vals, weights, u, order='cpu', rows=None):
    """nerf_process.py:62-67.  mids -> sample_pdf(weights[...,1:-1]) -> sort(cat)."""
    z_vals = _f32(z_vals)
    mids = F32(.5) * (z_vals[..., 1:] + z_vals[..., :-1])
    z_samples, inds = sample_pdf(mids, _f32(weights)[..., 1:-1], u, order, rows)
    z_fine = np.sort(np.concatenate([z_vals, z_samples], -1), -1)
    return z_fine.astype(np.float32), z_samples, inds


def embed_points(rays, z_vals, L_x=10, L_d=4):
    """nerf_process.py:34-39,69-85.  rays[N,6], z[N,S] -> embedded[N*S, (3+6L_x)+(3+6L_d)].

    viewdirs = d/||d|| (post-NDC d, SURVEY B-9); pts = o + d*z (un-normalised d).
    """
    rays = _f32(rays)
    o, d = rays[:, :3], rays[:, 3:]
    norm = np.sqrt((d * d).sum(-1, keepdims=True, dtype=np.float32)).astype(np.float32)
    viewdirs = d / norm
    pts = o[:, None, :] + d[:, None, :] * _f32(z_vals)[..., None]
    emb_x = positional_encoding(pts.reshape(-1, 3), L_x)
    dirs = np.broadcast_to(viewdirs[:, None, :], pts.shape).reshape(-1, 3)
    emb_d = positional_encoding(dirs, L_d)
    return np.concatenate([emb_x, emb_d], -1)


# ----------------------------------------------------------------------------------------------
# K4: MLP
# ----------------------------------------------------------------------------------------------
def mlp_param_names(D=8):
    """state_dict order of one NeRFModule (model/NeRF.py:24-30)."""
    names = []
    for i in range(D):
        names += [f'linear_x.{i}.weight', f'linear_x.{i}.bias']
    for m in ('linear_d', 'linear_feat', 'linear_density', 'linear_color'):
        names += [f'{m}.weight', f'{m}.bias']
    return names


def mlp_forward(params, x, D=8, skips=(4,), input_ch=63, return_acts=False):
    """model/NeRF.py:33-52.  params: dict name->ndarray of one NeRFModule; x[n, 63+27] -> [n,4].

    h = relu(W_i h + b_i); after layer i in skips h = [x63, h]; sigma = W_s h (no
    activation); f = W_f h; g = relu(W_d [f, d27]); rgb = W_c g; out = [rgb, sigma].
    """
    x = _f32(x)
    xin, din = x[:, :input_ch], x[:, input_ch:]
    acts = {}
    h = xin
    for i in range(D):
        acts[f'in{i}'] = h
        h = h @ params[f'linear_x.{i}.weight'].T + params[f'linear_x.{i}.bias']
        h = np.maximum(h, F32(0))
        if i in skips:
            h = np.concatenate([xin, h], -1)
    acts['trunk'] = h
    density = h @ params['linear_density.weight'].T + params['linear_density.bias']
    feat = h @ params['linear_feat.weight'].T + params['linear_feat.bias']
    hd = np.concatenate([feat, din], -1)
    acts['view_in'] = hd
    g = np.maximum(hd @ params['linear_d.weight'].T + params['linear_d.bias'], F32(0))
    acts['g'] = g
    rgb = g @ params['linear_color.weight'].T + params['linear_color.bias']
    out = np.concatenate([rgb, density], -1).astype(np.float32)
    if return_acts:
        return out, acts
    return out


def mlp_backward(params, x, d_out, D=8, skips=(4,), input_ch=63, acts=None):
    """Autograd of mlp_forward wrt the parameters (what loss.backward() gives,
    train.py:69).  d_out[n,4] -> dict name -> grad.  Input grads are not needed
    (sample positions are data, nerf_process.py:66).  acts: the forward's saved
    activations (autograd keeps them; recomputed here only if not supplied)."""
    if acts is None:
        _, acts = mlp_forward(params, x, D, skips, input_ch, return_acts=True)
    d_out = _f32(d_out)
    grads = {}
    d_rgb, d_sigma = d_out[:, :3], d_out[:, 3:4]
    g = acts['g']
    grads['linear_color.weight'] = d_rgb.T @ g
    grads['linear_color.bias'] = d_rgb.sum(0)
    dg = d_rgb @ params['linear_color.weight']
    dg = dg * (g > 0)
    grads['linear_d.weight'] = dg.T @ acts['view_in']
    grads['linear_d.bias'] = dg.sum(0)
    dhd = dg @ params['linear_d.weight']
    W = params['linear_feat.weight'].shape[0]
    dfeat = dhd[:, :W]
    h = acts['trunk']
    grads['linear_feat.weight'] = dfeat.T @ h
    grads['linear_feat.bias'] = dfeat.sum(0)
    grads['linear_density.weight'] = d_sigma.T @ h
    grads['linear_density.bias'] = d_sigma.sum(0)
    dh = dfeat @ params['linear_feat.weight'] + d_sigma @ params['linear_density.weight']
    for i in reversed(range(D)):
        if i in skips:
            dh = dh[:, input_ch:]
        # output of layer i (post relu); recompute mask from the stored next input
        nxt = acts[f'in{i + 1}'] if i + 1 < D else acts['trunk']
        post = nxt[:, input_ch:] if i in skips else nxt
        dh = dh * (post > 0)
        grads[f'linear_x.{i}.weight'] = dh.T @ acts[f'in{i}']
        grads[f'linear_x.{i}.bias'] = dh.sum(0)
        dh = dh @ params[f'linear_x.{i}.weight']
    return {k: v.astype(np.float32) for k, v in grads.items()}


# ----------------------------------------------------------------------------------------------
# K5: compositing
# ----------------------------------------------------------------------------------------------
def _sigmoid(x):
    return (F32(1.) / (F32(1.) + np.exp(-x))).astype(np.float32)


def post_process(outputs, z_vals, rays_d):
    """nerf_process.py:89-140.  raw[N,S,4], z[N,S], d[N,3] ->
    (rgb_map[N,3], disp_map[N], acc_map[N], weights[N,S], depth_map[N])."""
    raw = _f32(outputs)
    z = _f32(z_vals)
    d = _f32(rays_d)
    dists = z[..., 1:] - z[..., :-1]
    dists = np.concatenate([dists, np.full_like(dists[..., :1], 1e10)], -1)
    norm = np.sqrt((d * d).sum(-1, dtype=np.float32)).astype(np.float32)
    dists = dists * norm[:, None]
    rgb = _sigmoid(raw[..., :3])
    with np.errstate(over='ignore', invalid='ignore', divide='ignore'):
        alpha = F32(1.) - np.exp(-np.maximum(raw[..., 3], F32(0)) * dists)
        alpha = alpha.astype(np.float32)
        ones = np.ones((alpha.shape[0], 1), dtype=np.float32)
        trans = np.cumprod(np.concatenate([ones, F32(1.) - alpha + F32(1e-10)], -1), -1, dtype=np.float32)[:, :-1]
        weights = (alpha * trans).astype(np.float32)
        rgb_map = (weights[..., None] * rgb).sum(-2, dtype=np.float32)
        depth_map = (weights * z).sum(-1, dtype=np.float32)
        acc_map = weights.sum(-1, dtype=np.float32)
        q = depth_map / acc_map
        # torch.max(1e-10, nan) = nan -> where(isnan) -> 0
        disp = F32(1.) / np.where(np.isnan(q), q, np.maximum(F32(1e-10), q))
        disp = np.where(np.isnan(disp), F32(0), disp)
        disp = np.where(disp > F32(5.), F32(5.), disp)
        rgb_map = rgb_map + (F32(1.) - acc_map[..., None])
    return (rgb_map.astype(np.float32), disp.astype(np.float32), acc_map.astype(np.float32),
            weights, depth_map.astype(np.float32))


def post_process_backward(outputs, z_vals, rays_d, d_rgb_map):
    """Autograd of post_process wrt raw for an upstream gradient on rgb_map only
    (train.py:60-66: only rgb_c / rgb_f feed the loss).  float64 internally."""
    raw = np.asarray(outputs, dtype=np.float64)
    z = np.asarray(z_vals, dtype=np.float64)
    d = np.asarray(rays_d, dtype=np.float64)
    g = np.asarray(d_rgb_map, dtype=np.float64)
    N, S, _ = raw.shape
    dists = np.concatenate([z[..., 1:] - z[..., :-1], np.full((N, 1), 1e10)], -1)
    dists = dists * np.linalg.norm(d, axis=-1)[:, None]
    c = 1. / (1. + np.exp(-raw[..., :3]))
    sig = np.maximum(raw[..., 3], 0.)
    e = np.exp(-sig * dists)
    alpha = 1. - e
    f = 1. - alpha + 1e-10
    T = np.cumprod(np.concatenate([np.ones((N, 1)), f], -1), -1)[:, :-1]
    w = alpha * T
    dw = (c * g[:, None, :]).sum(-1) - g.sum(-1)[:, None]
    # R_s = sum_{k>s} dw_k alpha_k prod_{s<j<k} f_j  (reverse linear recurrence)
    R = np.zeros((N, S))
    for s in range(S - 2, -1, -1):
        R[:, s] = dw[:, s + 1] * alpha[:, s + 1] + f[:, s + 1] * R[:, s + 1]
    dalpha = dw * T - T * R
    dsigma = dalpha * dists * e * (raw[..., 3] > 0)
    draw = np.empty_like(raw)
    draw[..., :3] = (w[..., None] * g[:, None, :]) * c * (1. - c)
    draw[..., 3] = dsigma
    return draw.astype(np.float32)


# ----------------------------------------------------------------------------------------------
# drivers
# ----------------------------------------------------------------------------------------------
def render_rays(rays, params_coarse, params_fine, opts, t_rand, u, D=8, skips=(4,), L_x=10, L_d=4,
                return_all=False):
    """nerf_process.py:185-216 with RNG injected.  opts needs near, far, N_samples_c, N_samples_f."""
    rays = _f32(rays)
    z_c = stratified_z(opts.near, opts.far, opts.N_samples_c, t_rand)
    emb = embed_points(rays, z_c, L_x, L_d)
    raw_c, acts_c = mlp_forward(params_coarse, emb, D, skips, 3 + 6 * L_x, return_acts=True)
    raw_c = raw_c.reshape(z_c.shape[0], z_c.shape[1], 4)
    rgb_c, disp_c, acc_c, w_c, depth_c = post_process(raw_c, z_c, rays[:, 3:])
    ret = {'rgb_c': rgb_c, 'disp_c': disp_c}
    if return_all:
        ret.update(z_c=z_c, raw_c=raw_c, weights_c=w_c, emb_c=emb, acts_c=acts_c)
    if opts.N_samples_f > 0:
        z_f, _, _ = fine_z(z_c, w_c, u)
        emb_f = embed_points(rays, z_f, L_x, L_d)
        raw_f, acts_f = mlp_forward(params_fine, emb_f, D, skips, 3 + 6 * L_x, return_acts=True)
        raw_f = raw_f.reshape(z_f.shape[0], z_f.shape[1], 4)
        rgb_f, disp_f, acc_f, w_f, depth_f = post_process(raw_f, z_f, rays[:, 3:])
        ret.update(rgb_f=rgb_f, disp_f=disp_f)
        if return_all:
            ret.update(z_f=z_f, raw_f=raw_f, weights_f=w_f, emb_f=emb_f, acts_f=acts_f)
    return ret


def render(ray_o, ray_d, params_coarse, params_fine, opts, t_rand, u, H=None, W=None, focal=None, **kw):
    """nerf_process.py:220-252 (single chunk; chunking does not change values)."""
    o = _f32(ray_o).reshape(-1, 3)
    d = _f32(ray_d).reshape(-1, 3)
    if getattr(opts, 'data_type', 'blender') == 'llff':
        o, d = ndc_rays(H, W, focal, 1., o, d)
    rays = np.concatenate([o, d], -1)
    return render_rays(rays, params_coarse, params_fine, opts, t_rand, u, **kw)


def train_grads(rays, target, params_coarse, params_fine, opts, t_rand, u, D=8, skips=(4,), L_x=10, L_d=4):
    """train.py:57-69: loss = mean((rgb_c-t)^2) + mean((rgb_f-t)^2); returns
    (loss_c, loss_f, grads_coarse, grads_fine)."""
    r = render_rays(rays, params_coarse, params_fine, opts, t_rand, u, D, skips, L_x, L_d, return_all=True)
    rays = _f32(rays)
    target = _f32(target)
    n3 = F32(target.size)
    out = []
    for tag, params in (('c', params_coarse), ('f', params_fine)):
        diff = r[f'rgb_{tag}'] - target
        loss = float((diff.astype(np.float64) ** 2).mean())
        d_rgb = (F32(2.) * diff / n3).astype(np.float32)
        z = r[f'z_{tag}']
        draw = post_process_backward(r[f'raw_{tag}'], z, rays[:, 3:], d_rgb)
        grads = mlp_backward(params, r[f'emb_{tag}'], draw.reshape(-1, 4), D, skips, 3 + 6 * L_x, acts=r[f'acts_{tag}'])
        out.append((loss, grads))
    return out[0][0], out[1][0], out[0][1], out[1][1]


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (main.py:79-80) single-tensor update, fp32, no weight decay.
    step is the 1-based step count AFTER increment."""
    p, g, m, v = _f32(p), _f32(g), _f32(m), _f32(v)
    m = (m + (g - m) * F32(1 - beta1)).astype(np.float32)          # exp_avg.lerp_(grad, 1-beta1)
    v = (v * F32(beta2) + (g * g) * F32(1 - beta2)).astype(np.float32)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = (np.sqrt(v) / F32(math.sqrt(bc2)) + F32(eps)).astype(np.float32)
    p = (p - F32(step_size) * (m / denom)).astype(np.float32)
    return p, m, v


def lr_at(step, lr_max=5e-4, lr_min=5e-5, warmup=10000, iter_n=200000):
    """scheduler.py:54-64 as used from main.py:82-90 (first_cycle_steps = iter_N+1)."""
    if step < warmup:
        return (lr_max - lr_min) * step / warmup + lr_min
    return lr_min + (lr_max - lr_min) * (1 + math.cos(math.pi * (step - warmup) / (iter_n + 1 - warmup))) / 2


def make_opts(**kw):
    base = dict(near=2., far=6., N_samples_c=64, N_samples_f=128, perturb=1., data_type='blender',
                gpu_ids=[0], rank=0, chunk_rays=4096, chunk_pts=524288, N_rays=4096,
                precrop_iters=0, precrop_frac=.5)
    base.update(kw)
    return SimpleNamespace(**base)

"""Pins oracle/nerf_oracle.py against vectors produced by the unmodified reference
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, split_params
from oracle import nerf_oracle as orc


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def test_make_o_d_bit_exact():
    g = load_golden('raygen.npz')
    o, d = orc.make_o_d(int(g['W']), int(g['H']), g['K'], g['pose'])
    assert np.array_equal(o, g['rays_o'])
    assert np.array_equal(d, g['rays_d'])          # bit-exact: fma chain
    o8, d8 = orc.make_o_d(int(g['W8']), int(g['H8']), g['K8'], g['pose8'])
    sel = g['sel8']
    assert np.array_equal(o8.reshape(-1, 3)[sel], g['rays_o8'])
    assert np.array_equal(d8.reshape(-1, 3)[sel], g['rays_d8'])


def test_get_rays_np():
    g = load_golden('raygen.npz')
    o, d = orc.get_rays_np(int(g['H']), int(g['W']), g['K'], g['pose'])
    if str(g['np_dtype']) == str(d.dtype):
        assert np.array_equal(d, g['np_rays_d'])
    else:  # fixture made under another NumPy major version (SURVEY B-3)
        np.testing.assert_allclose(d, g['np_rays_d'], rtol=1e-6)
    assert np.array_equal(np.ascontiguousarray(o), g['np_rays_o'])


def test_ndc_rays_bit_exact():
    g = load_golden('ndc.npz')
    o, d = orc.ndc_rays(int(g['H']), int(g['W']), float(g['focal']), 1., g['rays_o'], g['rays_d'])
    assert np.array_equal(o, g['ndc_o'])
    assert np.array_equal(d, g['ndc_d'])


def test_positional_encoding():
    g = load_golden('posenc.npz')
    ex = orc.positional_encoding(g['x'], 10)
    ed = orc.positional_encoding(g['d'], 4)
    assert ex.shape[1] == int(g['out_dim_x']) == 63 and ed.shape[1] == int(g['out_dim_d']) == 27
    # identity block and argument products are exact; sin/cos differ by libm ulps only
    assert np.array_equal(ex[:, :3], g['enc_x'][:, :3])
    assert np.abs(ex - g['enc_x']).max() <= 5e-7
    assert np.abs(ed - g['enc_d']).max() <= 5e-7


def test_torch_linspace():
    g = load_golden('pre_process_coarse.npz')
    assert np.array_equal(orc.torch_linspace01(64), g['t_vals'])
    assert np.array_equal(orc.torch_linspace01(128), g['t_vals_128'])
    assert np.array_equal(orc.torch_linspace01(192), g['t_vals_192'])


def test_pre_process_coarse():
    g = load_golden('pre_process_coarse.npz')
    z = orc.stratified_z(float(g['near']), float(g['far']), 64, g['t_rand'])
    assert np.array_equal(z, g['z_vals'])          # bit-exact: individually rounded ops
    emb = orc.embed_points(g['rays'], z)
    assert emb.shape == g['embedded'].shape
    # PE arguments reach 2^9*|x| so an ulp of the point (norm/sum order) moves sin/cos by ~1e-4 at the top band
    assert np.abs(emb[:, :3] - g['embedded'][:, :3]).max() <= 1e-6
    assert np.abs(emb - g['embedded']).max() <= 2e-3
    assert np.abs(emb[:, :33] - g['embedded'][:, :33]).max() <= 5e-5


def test_sample_pdf_from_reference_cdf_bit_exact():
    """Given the reference's own cdf the inverse-CDF stage is bit-exact (indices AND samples)."""
    g = load_golden('sample_pdf.npz')
    for tag in ('det', 'rnd'):
        u = g[f'u_{tag}']
        u = np.broadcast_to(u, (g['cdf'].shape[0], u.shape[-1]))
        s, inds = orc.invert_cdf(g['bins'], g['cdf'], u)
        assert np.array_equal(inds, g[f'inds_{tag}'])
        assert np.array_equal(s, g[f'samples_{tag}'])


def test_sample_pdf_own_cdf():
    """With the oracle's own (fp64-accumulated) cdf, indices can differ from the reference's
    device-specific summation order only at knot ties (SURVEY B-5)."""
    g = load_golden('sample_pdf.npz')
    cdf = orc.pdf_to_cdf(g['weights'][..., 1:-1])
    assert ulp_diff(cdf, g['cdf']).max() <= 4
    for tag in ('det', 'rnd'):
        s, inds = orc.sample_pdf(g['bins'], g['weights'][..., 1:-1], g[f'u_{tag}'])
        mism = inds != g[f'inds_{tag}']
        assert mism.mean() < 2e-3, mism.mean()
        assert np.abs(inds - g[f'inds_{tag}']).max() <= 1
        # same bin -> same sample up to the ulp of the cdf; at a flipped tie the sample may jump by at
        # most one bin (flat-pdf bins use denom=1, nerf_process.py:179, so the inverse CDF is not continuous)
        # conditioning: d(sample) <= ulp(cdf)/denom * bin_width with denom >= 1e-5 (nerf_process.py:179)
        err = np.abs(s - g[f'samples_{tag}'])[~mism]
        assert err.max() <= 2e-3 and np.quantile(err, 0.999) <= 2e-5
        width = np.diff(g['bins'], axis=-1).max()
        assert np.abs(s - g[f'samples_{tag}']).max() <= width


def test_fine_z_and_embedding():
    g = load_golden('sample_pdf.npz')
    z_f, _, _ = orc.fine_z(g['fine_z_in'], g['fine_w_in'], g['fine_u'])
    assert z_f.shape == g['fine_z'].shape == (8, 192)
    assert np.abs(z_f - g['fine_z']).max() <= 2e-5
    assert np.all(np.diff(z_f, axis=-1) >= 0)


@pytest.mark.parametrize('S', [64, 192])
def test_post_process(S):
    g = load_golden(f'post_process_S{S}.npz')
    rgb, disp, acc, w, depth = orc.post_process(g['raw'], g['z_vals'], g['rays_d'])
    assert np.abs(w - g['weights']).max() <= 5e-6   # exp() libm ulps x weights<=1
    assert np.abs(rgb - g['rgb_map']).max() <= 1e-5
    assert np.abs(acc - g['acc_map']).max() <= 1e-5
    assert np.abs(depth - g['depth_map']).max() <= 1e-4
    assert np.abs(disp - g['disp_map']).max() <= 1e-4
    # edge cases (SURVEY B-8): empty rays -> disp 0, acc 0, white background
    assert np.all(disp[:8] == 0) and np.all(acc[:8] == 0) and np.all(rgb[:8] == 1)
    assert np.all(disp[32:40] <= 5.0)
    d_raw = orc.post_process_backward(g['raw'], g['z_vals'], g['rays_d'], g['d_rgb'])
    scale = np.abs(g['d_raw']).max()
    assert np.abs(d_raw - g['d_raw']).max() <= 1e-4 * max(scale, 1.)


def test_mlp_w64_forward_backward():
    g = load_golden('mlp_w64.npz')
    p = split_params(g, 'p')
    yc = orc.mlp_forward(p['coarse'], g['x'])
    yf = orc.mlp_forward(p['fine'], g['x'])
    assert np.abs(yc - g['y_coarse']).max() <= 1e-5
    assert np.abs(yf - g['y_fine']).max() <= 1e-5
    grads = orc.mlp_backward(p['coarse'], g['x'], g['d_y'])
    for k in grads:
        ref = g['g/' + k]
        assert grads[k].shape == ref.shape
        assert np.abs(grads[k] - ref).max() <= 1e-4 * max(1., np.abs(ref).max()), k


def test_render_and_train_w64():
    g = load_golden('render_train_w64.npz')
    p = split_params(g, 'p')
    opts = orc.make_opts(near=float(g['near']), far=float(g['far']))
    rays = np.concatenate([g['rays_o'], g['rays_d']], -1)
    r = orc.render_rays(rays, p['coarse'], p['fine'], opts, g['t_rand'], g['u'])
    for k in ('rgb_c', 'rgb_f'):
        assert np.abs(r[k] - g[k]).max() <= 1e-4, k
    for k in ('disp_c', 'disp_f'):
        assert np.abs(r[k] - g[k]).max() <= 1e-3, k
    lc, lf, gc, gf = orc.train_grads(rays, g['target'], p['coarse'], p['fine'], opts, g['t_rand'], g['u'])
    assert abs(lc - float(g['loss_c'])) <= 1e-5 and abs(lf - float(g['loss_f'])) <= 1e-5
    for tag, grads in (('coarse', gc), ('fine', gf)):
        num = den = 0.
        for k, v in grads.items():
            ref = g[f'g/model_{tag}.{k}']
            num += float(((v - ref).astype(np.float64) ** 2).sum())
            den += float((ref.astype(np.float64) ** 2).sum())
        assert np.sqrt(num / den) <= 1e-3, (tag, np.sqrt(num / den))
    # one Adam step (main.py:79-80)
    a = split_params(g, 'a')
    for tag, grads in (('coarse', gc), ('fine', gf)):
        for k, v in grads.items():
            ref_g = g[f'g/model_{tag}.{k}']
            newp, _, _ = orc.adam_step(p[tag][k], ref_g, np.zeros_like(ref_g), np.zeros_like(ref_g), 1, float(g['lr']))
            assert np.abs(newp - a[tag][k]).max() <= 1e-6, k


def test_render_llff_w64():
    g = load_golden('render_llff_w64.npz')
    a = split_params(g, 'a')
    opts = orc.make_opts(near=0., far=1., data_type='llff', perturb=0.)
    r = orc.render(g['rays_o'], g['rays_d'], a['coarse'], a['fine'], opts, g['t_rand'], g['u_det'],
                   H=int(g['H']), W=int(g['W']), focal=float(g['focal']))
    for k in ('rgb_c', 'rgb_f'):
        assert np.abs(r[k] - g[k]).max() <= 1e-4, k


def test_lr_schedule():
    assert abs(orc.lr_at(0) - 5e-5) < 1e-12
    assert abs(orc.lr_at(10000) - 5e-4) < 1e-12
    assert abs(orc.lr_at(5000) - (5e-5 + 4.5e-4 * 0.5)) < 1e-12
    assert orc.lr_at(200000) < 5.1e-5


def _seed0_params_w256(scale):
    """The seed-0 init of the full-width network, built with the product's host-side module (same RNG consumption as the
    reference's constructor; verified against the fixture's parameter checksums below)."""
    import torch
    from nerf_pytorch_paeng_b200.model import NeRF
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None))
    sd = {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}
    out = {'coarse': {}, 'fine': {}}
    for k, v in sd.items():
        net_name, rest = k.split('.', 1)
        out[net_name.replace('model_', '')][rest] = v * np.float32(scale) if rest == 'linear_density.weight' else v
    return sd, out


@pytest.mark.parametrize('tag', ['plain', 'dens30'])
def test_render_train_w256_random_init(tag):
    """The benchmarked configuration (full-width net at its PLAIN seed-0 random init, and the density x30 variant; 1024 rays,
    64+128) pinned on the unmodified reference: render, per-sample densities, losses and the whole gradient vectors."""
    g = load_golden(f'render_train_w256_{tag}.npz')
    sd, p = _seed0_params_w256(float(g['density_scale']))
    sums = np.array([float(v.astype(np.float64).sum()) for v in sd.values()])
    assert list(sd.keys()) == [str(s) for s in g['param_names']]
    assert np.abs(sums - g['param_sums_seed0']).max() <= 1e-9 * max(1., np.abs(sums).max())      # same seed-0 weights as the reference
    opts = orc.make_opts(near=float(g['near']), far=float(g['far']))
    rays = np.concatenate([g['rays_o'], g['rays_d']], -1)
    r = orc.render_rays(rays, p['coarse'], p['fine'], opts, g['t_rand'], g['u'])
    # two fp32 implementations (numpy/OpenBLAS vs torch/MKL): a ray whose LAST sample has |sigma| ~ 1e-8 can land on either side of
    # the 1e10-interval step of nerf_process.py:98; count them instead of hiding them
    for k in ('rgb_c', 'rgb_f'):
        e = np.abs(r[k] - g[k]).max(-1)
        assert (e > 1e-4).sum() <= 2, (k, (e > 1e-4).sum(), e.max())
    lc, lf, gc, gf = orc.train_grads(rays, g['target'], p['coarse'], p['fine'], opts, g['t_rand'], g['u'])
    assert abs(lc - float(g['loss_c'])) <= 1e-4 and abs(lf - float(g['loss_f'])) <= 1e-4
    for tag_n, grads, ref in (('coarse', gc, g['grad_coarse']), ('fine', gf, g['grad_fine'])):
        flat = np.concatenate([grads[k].reshape(-1) for k in orc.mlp_param_names()])
        rel = np.linalg.norm((flat - ref).astype(np.float64)) / np.linalg.norm(ref.astype(np.float64))
        assert rel <= 2e-3, (tag_n, rel)


@pytest.mark.parametrize('rows', [24, 8512, 40000])
def test_cdf_torch_cuda_order_vs_b200_fixture(rows):
    """The oracle's restatement of torch.sum / torch.cumsum's CUDA summation order (ATen Reduce.cuh / ScanUtils.cuh) against vectors
    the unmodified reference produced on a B200 (oracle/make_golden_cuda.py): row sums, cdf, bin indices and samples bit for bit."""
    g = load_golden('sample_pdf_cuda.npz')
    assert int(g[f'n{rows}_rows']) == rows
    z, w = g[f'n{rows}_z'], g[f'n{rows}_w']
    ww = (w[:, 1:-1] + np.float32(1e-5)).astype(np.float32)
    assert np.array_equal(orc.aten_cuda_sum_lastdim(ww), g[f'n{rows}_sum'])
    cdf = orc.pdf_to_cdf(w[:, 1:-1], order='cuda', rows=rows)
    assert np.array_equal(cdf, g[f'n{rows}_cdf'])
    u = orc.torch_linspace01(128)
    _, zs, inds = orc.fine_z(z, w, u, order='cuda', rows=rows)
    assert np.array_equal(inds, g[f'n{rows}_inds'].astype(np.int64))
    assert np.array_equal(zs, g[f'n{rows}_samples'])
    # the CPU order (fp64 accumulation) is a different function: it must NOT be silently identical
    if rows > 100:
        assert not np.array_equal(orc.pdf_to_cdf(w[:, 1:-1]), g[f'n{rows}_cdf'])

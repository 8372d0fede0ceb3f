"""Same-device parity: the UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.py; travels to the GPU box,
never enters git) executed on the B200 next to this repo's kernels, with mismatch COUNTS.  Skipped when
baseline/_ref is absent.  The bit-exact claims of north_star (ray origins/directions, bin indices at perturb=0) are checked here
against the reference's own CUDA execution -- the CPU-generated fixtures pin the same functions on the host."""
import os
import sys
from types import SimpleNamespace
from unittest import mock

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu

sys.path.insert(0, ROOT)
from baseline import ref_shim  # noqa: E402

if not ref_shim.available():
    pytest.skip('baseline/_ref not installed (python baseline/install_ref.py in the build container)', allow_module_level=True)


@pytest.fixture(scope='module')
def ctx():
    from nerf_pytorch_paeng_b200.engine import get_engine
    dev = torch.device('cuda', 0)
    return SimpleNamespace(ref=ref_shim.import_reference(), eng=get_engine(dev), dev=dev)


def bits_differ(a, b):
    return int((a.contiguous().view(torch.int32) != b.contiguous().view(torch.int32)).sum())


def test_make_o_d_and_ndc_bit_exact_on_device(ctx):
    ref, eng, dev = ctx.ref, ctx.eng, ctx.dev
    poses = ref.get_render_pose(n_angle=120, single_angle=-1, phi=-30.0, nf=4.0)
    focal = 0.5 * 800 / np.tan(0.5 * 0.6911112070083618)
    K = np.array([[focal, 0, 400.], [0, focal, 400.], [0, 0, 1]])
    for pi in (3, 64):
        pose = poses[pi, :3, :4].to(dev)
        o_r, d_r = ref.rays.make_o_d(800, 800, torch.from_numpy(K).to(dev), pose)
        o_m, d_m = eng.raygen(800, 800, K, pose)
        assert bits_differ(d_r.reshape(-1, 3), d_m) == 0 and bits_differ(o_r.reshape(-1, 3).contiguous(), o_m) == 0
    g = load_golden('ndc.npz')
    Hl, Wl, fl = int(g['H']), int(g['W']), float(g['focal'])
    Kl = np.array([[fl, 0, .5 * Wl], [0, fl, .5 * Hl], [0, 0, 1]])
    o_r, d_r = ref.rays.make_o_d(Wl, Hl, torch.from_numpy(Kl).to(dev), torch.from_numpy(g['pose']).to(dev))
    o_r, d_r = o_r.reshape(-1, 3).contiguous(), d_r.reshape(-1, 3).contiguous()
    on_r, dn_r = ref.proc.ndc_rays(Hl, Wl, Kl[0][0], 1., o_r, d_r)
    on_m, dn_m = eng.ndc_rays(Hl, Wl, fl, 1., o_r, d_r)
    assert bits_differ(on_r, on_m) == 0 and bits_differ(dn_r, dn_m) == 0


@pytest.mark.parametrize('n_rows', [24, 4096, 40000])      # ATen's cumsum kernel uses 32 / 16 / 512 threads per row of 62
def test_sample_pdf_det_bit_exact_on_device(ctx, n_rows):
    """perturb=0: cdf, bin indices and samples equal the reference's CUDA execution bit for bit (north_star)."""
    ref, eng, dev = ctx.ref, ctx.eng, ctx.dev
    opts = SimpleNamespace(near=2., far=6., gpu_ids=[0], rank=0, N_samples_c=64, N_samples_f=128, perturb=0., chunk_pts=524288,
                           chunk_rays=4096, data_type='blender')
    gen = torch.Generator(device='cpu').manual_seed(n_rows)
    z = torch.sort(torch.rand(n_rows, 64, generator=gen) * 4 + 2, -1)[0].to(dev)
    w = (torch.rand(n_rows, 64, generator=gen) ** 8)
    w[:n_rows // 50] = 0.                       # empty rays: flat pdf
    w = w.to(dev)
    mids = .5 * (z[..., 1:] + z[..., :-1])
    rec = []
    real = torch.searchsorted
    with mock.patch('torch.searchsorted', lambda *a, **k: rec.append(real(*a, **k)) or rec[-1]):
        s_ref = ref.proc.sample_pdf(mids, w[..., 1:-1], 128, det=True, opts=opts)
    ww = w[..., 1:-1] + 1e-5
    pdf = ww / torch.sum(ww, -1, keepdim=True)
    cdf_ref = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    z_f, zs, inds, cdf = eng.sample_pdf(z, w, 128, u=torch.linspace(0., 1., steps=128, device=dev), want_samples=True,
                                        want_inds=True, want_cdf=True)
    assert bits_differ(cdf, cdf_ref) == 0
    assert int((inds != rec[-1]).sum()) == 0
    assert bits_differ(zs, s_ref) == 0
    assert torch.equal(z_f, torch.sort(torch.cat([z, s_ref], -1), -1)[0])


def test_post_process_on_device(ctx):
    ref, eng, dev = ctx.ref, ctx.eng, ctx.dev
    g = load_golden('post_process_S192.npz')
    raw, zz, dd = (torch.from_numpy(g[k]).to(dev) for k in ('raw', 'z_vals', 'rays_d'))
    outs = ref.proc.post_process(raw, zz, dd)
    rgb, disp, acc, wts, depth = eng.composite_forward(raw, zz, dd)
    assert float((rgb - outs[0]).abs().max()) <= 1e-6 and float((wts - outs[3]).abs().max()) <= 1e-6
    assert float((depth - outs[4]).abs().max()) <= 1e-5 and float((disp - outs[1]).abs().max()) <= 1e-6


@pytest.mark.parametrize('n_rows,s_c,s_f', [(1000, 16, 32), (5000, 34, 64), (777, 100, 128), (4096, 128, 256), (40, 10, 16)])
def test_sample_pdf_det_other_sizes_on_device(ctx, n_rows, s_c, s_f):
    """The torch-CUDA summation order is restated for every row width below 128 (ATen picks block widths from the shape): other
    N_samples_c than the shipped 64, checked against the reference's CUDA execution bit for bit."""
    ref, eng, dev = ctx.ref, ctx.eng, ctx.dev
    opts = SimpleNamespace(near=2., far=6., gpu_ids=[0], rank=0, N_samples_c=s_c, N_samples_f=s_f, perturb=0., chunk_pts=524288,
                           chunk_rays=4096, data_type='blender')
    gen = torch.Generator(device='cpu').manual_seed(1000 * s_c + n_rows)
    z = torch.sort(torch.rand(n_rows, s_c, generator=gen) * 4 + 2, -1)[0].to(dev)
    w = (torch.rand(n_rows, s_c, generator=gen) ** 6).to(dev)
    mids = .5 * (z[..., 1:] + z[..., :-1])
    rec = []
    real = torch.searchsorted
    with mock.patch('torch.searchsorted', lambda *a, **k: rec.append(real(*a, **k)) or rec[-1]):
        s_ref = ref.proc.sample_pdf(mids, w[..., 1:-1], s_f, det=True, opts=opts)
    ww = w[..., 1:-1] + 1e-5
    pdf = ww / torch.sum(ww, -1, keepdim=True)
    cdf_ref = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    _, zs, inds, cdf = eng.sample_pdf(z, w, s_f, u=torch.linspace(0., 1., steps=s_f, device=dev), want_samples=True, want_inds=True,
                                      want_cdf=True)
    assert bits_differ(cdf, cdf_ref) == 0, (n_rows, s_c)
    assert int((inds != rec[-1]).sum()) == 0
    assert bits_differ(zs, s_ref) == 0

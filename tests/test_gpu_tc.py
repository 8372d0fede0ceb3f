"""GPU tests of the tcgen05 (NB_BF16) MLP path: step-by-step against a numpy emulation of the same
bf16 data flow, and end to end against the fp32 path / oracle (north_star: >= 50 dB PSNR, gradients
within 1e-2 relative)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import nerf_oracle as orc

pytestmark = pytest.mark.gpu


def bf16(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).bfloat16().float().numpy()


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def npy(t):
    return t.detach().cpu().numpy()


def net_params(net):
    sd = {k: npy(v) for k, v in net.state_dict().items()}
    pc = {k[len('model_coarse.'):]: v for k, v in sd.items() if k.startswith('model_coarse.')}
    pf = {k[len('model_fine.'):]: v for k, v in sd.items() if k.startswith('model_fine.')}
    return pc, pf


def emulate_chain(p, emb):
    """Same data flow as nb_mlp_tc.cu: bf16 operands, fp32 accumulate, fp32 bias/ReLU, bf16 re-quantisation of
    every layer input; sigma / rgb heads from the fp32 activations.  Returns (accs per step, raw)."""
    ex, ed = bf16(emb[:, :63]), bf16(emb[:, 63:])
    W = {k: bf16(v) for k, v in p.items() if k.endswith('weight')}
    accs = []
    a = ex
    h32 = None
    for i in range(8):
        if i == 5:
            a = np.concatenate([ex, a], -1)
        acc = a @ W[f'linear_x.{i}.weight'].T
        accs.append(acc)
        h32 = np.maximum(acc + p[f'linear_x.{i}.bias'], 0)
        a = bf16(h32)
    sigma = h32 @ p['linear_density.weight'].T + p['linear_density.bias']
    acc = a @ W['linear_feat.weight'].T
    accs.append(acc)
    feat = bf16(acc + p['linear_feat.bias'])
    acc = np.concatenate([feat, ed], -1) @ W['linear_d.weight'].T
    accs.append(acc)
    g32 = np.maximum(acc + p['linear_d.bias'], 0)
    rgb = g32 @ p['linear_color.weight'].T + p['linear_color.bias']
    return accs, np.concatenate([rgb, sigma], -1).astype(np.float32)


@pytest.fixture(scope='module')
def setup():
    from nerf_pytorch_paeng_b200.engine import get_engine
    from nerf_pytorch_paeng_b200.model import NeRF
    eng = get_engine(torch.device('cuda', 0))
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).cuda()
    with torch.no_grad():
        for m in (net.model_coarse, net.model_fine):
            m.linear_density.weight.mul_(30.)
            for lin in list(m.linear_x) + [m.linear_feat, m.linear_d, m.linear_density, m.linear_color]:
                lin.bias.uniform_(-0.1, 0.1)
    g = load_golden('raygen.npz')
    return eng, net, g


def make_rays(g, n, s, seed=0):
    rs = np.random.RandomState(seed)
    rays = np.concatenate([np.tile(g['rays_o8'][:1], (n, 1)), g['rays_d8'][:n]], -1).astype(np.float32)
    z = np.sort(rs.rand(n, s).astype(np.float32) * 4 + 2, -1)
    return rays, z


@pytest.mark.parametrize('step', [0, 1, 5, 8, 9])
def test_tc_chain_steps(setup, step):
    eng, net, g = setup
    net.set_precision('bf16')
    m = net.model_coarse
    flat = m.flat_params()
    rays, z = make_rays(g, 5, 60)         # 300 points: 3 tiles (one partial), both slots
    acc, raw = eng.mlp_tc_probe(m.desc, flat, m.packed_weights(), cu(rays), cu(z), step)
    torch.cuda.synchronize()
    pc, _ = net_params(net)
    emb = orc.embed_points(rays, z)
    accs, raw_e = emulate_chain(pc, emb)
    n = 128 if step == 9 else 256
    err = np.abs(npy(acc)[:, :n] - accs[step]).max()
    scale = np.abs(accs[step]).max()
    assert err <= 2e-2 * max(1., scale), (step, err, scale)
    assert np.abs(npy(raw) - raw_e).max() <= 3e-2 * max(1., np.abs(raw_e).max())


def test_tc_forward_vs_fp32(setup):
    """bf16 tensor-core path against the fp32 CUDA-core path on the same rays (many tiles: persistent
    loop with several iterations per CTA, ragged tail)."""
    eng, net, g = setup
    n, s = 4096, 73
    rays, z = make_rays(g, n, s, seed=1)
    outs = {}
    for prec in ('fp32', 'bf16'):
        net.set_precision(prec)
        with torch.no_grad():
            outs[prec] = npy(net.model_fine.forward_rays(cu(rays), cu(z)))
    ref, got = outs['fp32'], outs['bf16']
    assert np.isfinite(got).all()
    rel = np.abs(got - ref).max() / np.abs(ref).max()
    rms = np.sqrt(((got - ref) ** 2).mean()) / np.sqrt((ref ** 2).mean())
    assert rel <= 5e-2 and rms <= 1e-2, (rel, rms)
    # materialised-embedding entry (model(x)) takes the same kernel
    emb = orc.embed_points(rays[:3], z[:3])
    with torch.no_grad():
        y = npy(net(cu(emb), is_fine=True))
    assert np.abs(y - ref[:3 * s]).max() <= 5e-2 * np.abs(ref).max()


def test_tc_render_psnr(setup):
    """End-to-end coarse+fine render, bf16 MLP vs fp32 MLP with identical random draws: >= 50 dB."""
    from types import SimpleNamespace
    from nerf_pytorch_paeng_b200 import nerf_process
    eng, net, g = setup
    n = 2048
    rs = np.random.RandomState(3)
    rays = np.concatenate([np.tile(g['rays_o8'][:1], (n, 1)), g['rays_d8'][:n]], -1).astype(np.float32)
    rng = {'t_rand': cu(rs.rand(n, 64)), 'u': cu(rs.rand(n, 128))}
    opts = SimpleNamespace(near=2., far=6., N_samples_c=64, N_samples_f=128, perturb=1., data_type='blender', gpu_ids=[0],
                           rank=0, chunk_rays=4096, chunk_pts=524288, seed=0, rng=rng)
    out = {}
    for prec in ('fp32', 'bf16'):
        net.set_precision(prec)
        with torch.no_grad():
            out[prec] = nerf_process.render_rays(cu(rays), net, None, opts)
    for k in ('rgb_c', 'rgb_f'):
        mse = float(((out['bf16'][k] - out['fp32'][k]) ** 2).mean())
        psnr = -10 * np.log10(max(mse, 1e-20))
        assert psnr >= 50., (k, psnr)

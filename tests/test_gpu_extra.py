"""GPU tests of the rows either side of the hot path and of the larger configurations:
frame output, checkpoint compatibility, the reference-signature train()/test()/render() entries, the config-5
sample-count sweep, and (when >= 2 GPUs are visible) N-GPU == 1-GPU equivalence."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import nerf_oracle as orc

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def npy(t):
    return t.detach().cpu().numpy()


def make_opts(**kw):
    base = dict(near=2., far=6., N_samples_c=64, N_samples_f=128, perturb=1., data_type='blender', gpu_ids=[0], rank=0,
                chunk_rays=4096, chunk_pts=524288, N_rays=1024, precrop_iters=0, precrop_frac=.5, seed=0, global_batch=False,
                idx_print=10 ** 9, idx_save=None, exp_name='t', cdf_order='fp64')    # fixtures / oracle: CPU summation order
    base.update(kw)
    return SimpleNamespace(**base)


def test_frame_to8b():
    """nb_frame_to8b == to8b(rgb), to8b(disp/nanmax(disp)) of test.py:50-61 / utils.py:11."""
    from nerf_pytorch_paeng_b200.engine import get_engine
    from nerf_pytorch_paeng_b200.utils import to8b
    eng = get_engine(torch.device('cuda', 0))
    rs = np.random.RandomState(0)
    rgb = (rs.rand(1000, 3) * 1.4 - 0.2).astype(np.float32)
    disp = (rs.rand(1000) * 5).astype(np.float32)
    disp[7] = np.nan
    r8, d8 = eng.frame_to8b(cu(rgb), cu(disp))
    assert np.array_equal(npy(r8), to8b(rgb))
    exp = to8b(np.nan_to_num(disp / np.nanmax(disp), nan=0.0))
    assert np.abs(npy(d8).astype(int) - exp.astype(int)).max() <= 1      # x/max rounding at a bin edge


def test_checkpoint_roundtrip_reference_format(tmp_path):
    """train.py:105-114 checkpoint dict written by one model loads into another (same keys as the reference)
    and reproduces its outputs; FlatAdam state survives a save/load."""
    from nerf_pytorch_paeng_b200 import trainer
    from nerf_pytorch_paeng_b200.model import NeRF
    torch.manual_seed(1)
    a = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).cuda().set_precision('bf16')
    opt = trainer.FlatAdam(a, lr=5e-4)
    x = torch.rand(300, 90, device='cuda')
    ya = a(x)
    ck = {'idx': 7, 'model_state_dict': a.state_dict(), 'optimizer_state_dict': opt.state_dict()}
    path = os.path.join(tmp_path, 'blender_lego_7.pth.tar')
    torch.save(ck, path)
    torch.manual_seed(2)
    b = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).cuda().set_precision('bf16')
    assert not torch.equal(b(x), ya)
    b.load_state_dict(torch.load(path, map_location='cpu')['model_state_dict'])
    assert torch.equal(b(x), ya)          # packed bf16 weights are refreshed after load_state_dict
    # a plain torch Adam over the same parameters also refreshes them after its in-place step
    o2 = torch.optim.Adam(b.parameters(), lr=1e-2)
    (b(x).sum()).backward()
    o2.step()
    assert not torch.equal(b(x), ya)


@pytest.mark.parametrize('sc,sf', [(128, 256), (256, 512)])
def test_sample_count_sweep(sc, sf):
    """BASELINE config 5: larger sample counts run through every kernel (fused bf16 path vs fp32 path)."""
    from nerf_pytorch_paeng_b200 import nerf_process
    from nerf_pytorch_paeng_b200.model import NeRF
    g = load_golden('raygen.npz')
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).cuda()
    n = 96
    rs = np.random.RandomState(sc)
    rays = cu(np.concatenate([np.tile(g['rays_o8'][:1], (n, 1)), g['rays_d8'][:n]], -1))
    opts = make_opts(N_samples_c=sc, N_samples_f=sf, rng={'t_rand': cu(rs.rand(n, sc)), 'u': cu(rs.rand(n, sf))})
    out = {}
    for prec in ('fp32', 'bf16'):
        net.set_precision(prec)
        with torch.no_grad():
            out[prec] = nerf_process.render_rays(rays, net, None, opts)
    for k in ('rgb_c', 'rgb_f'):
        assert out['fp32'][k].shape == (n, 3)
        mse = float(((out['bf16'][k] - out['fp32'][k]) ** 2).mean())
        assert -10 * np.log10(max(mse, 1e-20)) >= 50.
    # fp32 coarse pass against the oracle at this sample count
    sd = {k: npy(v) for k, v in net.state_dict().items()}
    pc = {k[len('model_coarse.'):]: v for k, v in sd.items() if k.startswith('model_coarse.')}
    pf = {k[len('model_fine.'):]: v for k, v in sd.items() if k.startswith('model_fine.')}
    t_dev = npy(torch.linspace(0., 1., steps=sc, device='cuda'))
    z = orc.stratified_z(2., 6., sc, npy(opts.rng['t_rand']), t_vals=t_dev)
    raw = orc.mlp_forward(pc, orc.embed_points(npy(rays), z)).reshape(n, sc, 4)
    rgb, *_ = orc.post_process(raw, z, npy(rays)[:, 3:])
    assert np.abs(npy(out['fp32']['rgb_c']) - rgb).max() <= 1e-4


@pytest.mark.parametrize('data_type', ['blender', 'llff'])
def test_train_entry_reference_signature(data_type):
    """nerf_pytorch_paeng_b200.train.train(...) (train.py:12 signature): per-image path with host images, both pixel
    selection modes, torch.optim.Adam and FlatAdam; the loss goes down on a constant-colour target."""
    from nerf_pytorch_paeng_b200 import train as train_mod, trainer
    from nerf_pytorch_paeng_b200.model import NeRF, get_positional_encoder
    H, W = 60, 80
    K = np.array([[100., 0, W / 2], [0, 100., H / 2], [0, 0, 1.]])
    g = load_golden('raygen.npz' if data_type == 'blender' else 'ndc.npz')
    poses = (g['all_poses'][:3] if data_type == 'blender' else g['llff_poses'][:3]).astype(np.float32)
    images = [np.full((H, W, 3), 0.25, np.float32) for _ in range(3)]
    posenc = [get_positional_encoder(10)[0], get_positional_encoder(4)[0]]
    near, far = (2., 6.) if data_type == 'blender' else (0., 1.)
    for fast in (False, True):
        torch.manual_seed(0)
        np.random.seed(0)
        net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(K, poses)).cuda().set_precision('bf16')
        opts = make_opts(data_type=data_type, near=near, far=far, N_rays=512, device_select=fast, precrop_iters=2)
        optimizer = trainer.FlatAdam(net, lr=2e-3) if fast else torch.optim.Adam(net.parameters(), lr=2e-3)
        losses = [float(train_mod.train(i, [0, 1, 2], images, (K, poses), (H, W), net, torch.nn.MSELoss(), posenc, optimizer,
                                        None, None, opts)) for i in range(1, 13)]
        assert np.isfinite(losses).all()
        assert np.mean(losses[-3:]) < 0.6 * np.mean(losses[:3]), losses


def test_test_and_render_entries():
    """test.test / test.render (test.py:17,111 signatures): frames come back as uint8 [H,W,3], PSNR finite."""
    from nerf_pytorch_paeng_b200 import test as test_mod
    from nerf_pytorch_paeng_b200.model import NeRF
    g = load_golden('raygen.npz')
    H, W = 24, 32
    K = np.array([[40., 0, W / 2], [0, 40., H / 2], [0, 0, 1.]])
    poses = g['all_poses'][:2].astype(np.float32)
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(K, poses)).cuda().set_precision('bf16')
    opts = make_opts(exp_name='nonexistent_ckpt')
    imgs = [np.random.rand(H, W, 3).astype(np.float32) for _ in range(2)]
    with pytest.raises(FileNotFoundError):          # like the reference (test.py:20-21), a missing checkpoint is an error
        test_mod.test(0, [0, 1], None, net, imgs, K, poses, (H, W), opts, save=False)
    opts.allow_missing_checkpoint = True
    res = test_mod.test(0, [0, 1], None, net, imgs, K, poses, (H, W), opts, save=False)
    assert len(res['frames']) == 2 and res['frames'][0].shape == (H, W, 3) and res['frames'][0].dtype == np.uint8
    assert np.isfinite(res['psnr']).all()
    rgbs, disps = test_mod.render(0, None, net, K, poses, (H, W), opts, save=False)
    assert rgbs.shape == (2, H, W, 3) and disps.shape == (2, H, W) and rgbs.dtype == np.uint8


def _ddp_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    from nerf_pytorch_paeng_b200 import distributed, trainer
    from nerf_pytorch_paeng_b200.model import NeRF
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world)
    ctx = distributed.DistContext()
    g = load_golden('raygen.npz')
    n = 512
    rs = np.random.RandomState(0)
    rays = np.concatenate([np.tile(g['rays_o8'][:1], (n, 1)), g['rays_d8'][:n]], -1).astype(np.float32)
    target, t_rand, u = rs.rand(n, 3).astype(np.float32), rs.rand(n, 64).astype(np.float32), rs.rand(n, 128).astype(np.float32)
    lo, hi = ctx.shard_range(n)
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).cuda().set_precision('fp32')
    dev = lambda a: torch.from_numpy(a[lo:hi].copy()).cuda()
    opts = make_opts(gpu_ids=list(range(world)), rank=rank, rng={'t_rand': dev(t_rand), 'u': dev(u)})
    out = trainer.render_losses_and_grads(net, dev(rays), dev(target), opts, n_global=n)
    ctx.allreduce_grads(net)
    ctx.allreduce_(out['loss_buf'])
    # render: bands gathered to the full frame on every rank
    full = ctx.gather_rows(out['rgb_f'], n)
    torch.save({'gc': net.model_coarse.flat_grad.cpu(), 'gf': net.model_fine.flat_grad.cpu(), 'loss': out['loss_buf'].cpu(),
                'rgb_f': full.cpu()}, os.path.join(out_dir, f'r{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs (gpurun --gpus 2)')
def test_two_gpu_equals_one_gpu(tmp_path):
    """ray-sharded data parallel over NCCL: summed rank gradients (loss normalised by the GLOBAL ray
    count) == the single-GPU gradient of the whole batch; gathered render bands == the single-GPU render."""
    import torch.multiprocessing as mp
    from nerf_pytorch_paeng_b200 import trainer
    from nerf_pytorch_paeng_b200.model import NeRF
    world, port = 2, 29600 + os.getpid() % 300
    mp.spawn(_ddp_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = load_golden('raygen.npz')
    n = 512
    rs = np.random.RandomState(0)
    rays = np.concatenate([np.tile(g['rays_o8'][:1], (n, 1)), g['rays_d8'][:n]], -1).astype(np.float32)
    target, t_rand, u = rs.rand(n, 3).astype(np.float32), rs.rand(n, 64).astype(np.float32), rs.rand(n, 128).astype(np.float32)
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).cuda().set_precision('fp32')
    opts = make_opts(rng={'t_rand': cu(t_rand), 'u': cu(u)})
    out = trainer.render_losses_and_grads(net, cu(rays), cu(target), opts)
    for r in range(world):
        d = torch.load(os.path.join(tmp_path, f'r{r}.pt'))
        for key, ref in (('gc', net.model_coarse.flat_grad), ('gf', net.model_fine.flat_grad)):
            rel = float((d[key] - ref.cpu()).norm() / ref.cpu().norm())
            assert rel <= 1e-4, (key, rel)        # fp32 atomics order only
        assert float((d['loss'] - out['loss_buf'].cpu()).abs().max()) <= 1e-6
        assert float((d['rgb_f'] - out['rgb_f'].cpu()).abs().max()) <= 1e-5


def _ddp_worker_modes(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    from nerf_pytorch_paeng_b200 import distributed, trainer
    from nerf_pytorch_paeng_b200.model import NeRF
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world)
    ctx = distributed.DistContext()
    g = load_golden('raygen.npz')
    n = 1024
    rs = np.random.RandomState(1)
    rays = np.concatenate([np.tile(g['rays_o8'][:1], (n, 1)), g['rays_d8'][:n]], -1).astype(np.float32)
    lo, hi = ctx.shard_range(n)
    dev = lambda a: torch.from_numpy(a[lo:hi].copy()).cuda()
    steps = [(rs.rand(n, 3).astype(np.float32), rs.rand(n, 64).astype(np.float32), rs.rand(n, 128).astype(np.float32)) for _ in range(3)]
    res = {}
    for mode in ('joint', 'peer'):
        os.environ['NB_DP_MODE'] = mode
        torch.manual_seed(0)
        net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).cuda().set_precision('bf16')
        opt = trainer.FlatAdam(net, lr=5e-4)
        losses = []
        for target, t_rand, u in steps:
            opts = make_opts(gpu_ids=list(range(world)), rank=rank, rng={'t_rand': dev(t_rand), 'u': dev(u)})
            losses.append(trainer.train_step(net, opt, dev(rays), dev(target), opts, dist_ctx=ctx).cpu())
        torch.cuda.synchronize()
        res[mode] = {'pc': net.model_coarse.flat.cpu(), 'pf': net.model_fine.flat.cpu(), 'gc': net.model_coarse.flat_grad.cpu(),
                     'gf': net.model_fine.flat_grad.cpu(), 'loss': torch.stack(losses)}
    torch.save(res, os.path.join(out_dir, f'm{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs (gpurun --gpus 2)')
def test_two_gpu_bf16_joint_and_peer_exchange(tmp_path):
    """The benchmarked data-parallel routes on the bf16 path: NB_DP_MODE=joint (one NCCL all-reduce of the joint buffer) and
    NB_DP_MODE=peer (copy-engine pushes + sum folded into Adam) give every rank the same gradients / losses / weights as each
    other and as ONE GPU stepping on the concatenated batch."""
    import torch.multiprocessing as mp
    from nerf_pytorch_paeng_b200 import trainer
    from nerf_pytorch_paeng_b200.model import NeRF
    world, port = 2, 29900 + os.getpid() % 300
    mp.spawn(_ddp_worker_modes, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = load_golden('raygen.npz')
    n = 1024
    rs = np.random.RandomState(1)
    rays = np.concatenate([np.tile(g['rays_o8'][:1], (n, 1)), g['rays_d8'][:n]], -1).astype(np.float32)
    steps = [(rs.rand(n, 3).astype(np.float32), rs.rand(n, 64).astype(np.float32), rs.rand(n, 128).astype(np.float32)) for _ in range(3)]
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).cuda().set_precision('bf16')
    opt = trainer.FlatAdam(net, lr=5e-4)
    losses = []
    for i, (target, t_rand, u) in enumerate(steps):
        opts = make_opts(rng={'t_rand': cu(t_rand), 'u': cu(u)})
        losses.append(trainer.train_step(net, opt, cu(rays), cu(target), opts).cpu())
        if i == 0:
            g1 = (net.model_coarse.flat_grad.cpu().clone(), net.model_fine.flat_grad.cpu().clone())
    one = {'gc': net.model_coarse.flat_grad.cpu(), 'gf': net.model_fine.flat_grad.cpu(), 'loss': torch.stack(losses)}
    rel = lambda a, b: float((a - b).norm() / b.norm())
    d = [torch.load(os.path.join(tmp_path, f'm{r}.pt')) for r in range(world)]
    for mode in ('joint', 'peer'):
        # replicas stay bit-identical (same summed gradient, same Adam) ...
        assert torch.equal(d[0][mode]['pc'], d[1][mode]['pc']) and torch.equal(d[0][mode]['pf'], d[1][mode]['pf']), mode
        # ... and equal the single-GPU run on the whole batch: first-step losses exactly comparable, later steps drift with Adam's
        # sign-like early updates
        assert float((d[0][mode]['loss'][0] - one['loss'][0]).abs().max()) <= 1e-5, mode
        assert float((d[0][mode]['loss'] - one['loss']).abs().max()) <= 5e-3, mode
    assert rel(d[0]['peer']['gc'], d[0]['joint']['gc']) <= 1e-2 and rel(d[0]['peer']['gf'], d[0]['joint']['gf']) <= 1e-2
    assert float((d[0]['peer']['loss'] - d[0]['joint']['loss']).abs().max()) <= 1e-3


def test_train_entry_global_batch():
    """train.py:25-32: the global-batch path (device-resident shuffled [N,3,3] ray/rgb table + GetterRayBatchIdx cursor)."""
    from nerf_pytorch_paeng_b200 import rays as rays_mod, train as train_mod, trainer
    from nerf_pytorch_paeng_b200.model import NeRF, get_positional_encoder
    from nerf_pytorch_paeng_b200.utils import GetterRayBatchIdx
    g = load_golden('raygen.npz')
    H, W = 30, 40
    K = np.array([[50., 0, W / 2], [0, 50., H / 2], [0, 0, 1.]])
    poses = g['all_poses'][:2].astype(np.float32)
    table = []
    for pose in poses:                                   # main.py:95-101
        o, d = rays_mod.get_rays_np(H, W, K, pose)
        table.append(np.stack([o, d, np.full_like(d, 0.3)], 0))
    rays_rgb = np.transpose(np.stack(table, 0), [0, 2, 3, 1, 4]).reshape(-1, 3, 3).astype(np.float32)
    np.random.seed(0)
    np.random.shuffle(rays_rgb)
    cursor = GetterRayBatchIdx(torch.from_numpy(rays_rgb).cuda())
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(K, poses)).cuda().set_precision('bf16')
    opts = make_opts(N_rays=600, global_batch=True)
    opt = trainer.FlatAdam(net, lr=2e-3)
    posenc = [get_positional_encoder(10)[0], get_positional_encoder(4)[0]]
    losses = [float(train_mod.train(i, [0, 1], None, (K, poses), (H, W), net, torch.nn.MSELoss(), posenc, opt, cursor, None, opts))
              for i in range(1, 11)]
    assert cursor.epoch >= 1                             # 2400 rays / 600 per step: wrapped and reshuffled (utils.py:54-58)
    assert np.isfinite(losses).all() and np.mean(losses[-3:]) < 0.7 * np.mean(losses[:3]), losses


def test_full_size_properties():
    """BASELINE configs[1] at full size (4096 rays, 64+128) through size-independent properties of the domain."""
    from nerf_pytorch_paeng_b200 import nerf_process, trainer
    from nerf_pytorch_paeng_b200.engine import get_engine
    from nerf_pytorch_paeng_b200.model import NeRF
    eng = get_engine(torch.device('cuda', 0))
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).cuda().set_precision('bf16')
    with torch.no_grad():
        net.model_coarse.linear_density.weight.mul_(30.)
        net.model_fine.linear_density.weight.mul_(30.)
    n = 4096
    K = np.array([[1111.111, 0, 400.], [0, 1111.111, 400.], [0, 0, 1.]])
    g = load_golden('raygen.npz')
    pix = eng.select_pixels(n, 800, 800, seed=3)
    o, d = eng.raygen(800, 800, K, cu(g['pose8']), pix_idx=pix)
    rays = torch.cat((o, d), -1)
    opts = make_opts(N_rays=n, seed=11)
    z_c = nerf_process._coarse_z(rays, opts)
    m = net.model_coarse
    raw, _ = eng.mlp_forward(m.desc, m.flat_params(), m.packed_weights(), m.precision, rays=rays, z=z_c)
    rgb, disp, acc, w, depth = eng.composite_forward(raw.view(n, 64, 4), z_c, d)
    # compositing: weights are a sub-probability vector; rgb = sum w*sigmoid(c) + (1 - acc); depth inside [near, far]
    assert float(w.min()) >= 0. and float(acc.max()) <= 1. + 1e-5
    recon = (w[..., None] * torch.sigmoid(raw.view(n, 64, 4)[..., :3])).sum(1) + (1. - acc[:, None])
    assert float((recon - rgb).abs().max()) <= 1e-5
    assert float(rgb.min()) >= -1e-6 and float(rgb.max()) <= 1. + 1e-5
    assert float((depth - acc * 6.0).max()) <= 1e-3 and float((depth - acc * 2.0).min()) >= -1e-3
    assert float(disp.min()) >= 0. and float(disp.max()) <= 5.
    # hierarchical sampling: sorted, 192 per ray, contains every coarse depth, new samples inside the coarse span
    z_f = nerf_process._fine_z(rays, opts, z_c, w)
    assert z_f.shape == (n, 192) and bool((z_f[:, 1:] >= z_f[:, :-1]).all())
    both = torch.sort(torch.cat([z_f, z_c], -1), -1)[0]
    assert int((both[:, 1:] == both[:, :-1]).sum(-1).min()) >= 64           # every coarse value appears again
    assert float(z_f.min()) >= float(z_c.min()) - 1e-6 and float(z_f.max()) <= float(z_c.max()) + 1e-6
    # MLP rows are independent: permuting the rays permutes raw bit-exactly (tile / slot / cluster placement invariance)
    perm = torch.randperm(n, device='cuda')
    raw_p, _ = eng.mlp_forward(m.desc, m.flat_params(), m.packed_weights(), m.precision, rays=rays[perm].contiguous(), z=z_c[perm].contiguous())
    assert torch.equal(raw_p.view(n, 64, 4), raw.view(n, 64, 4)[perm])
    # a ragged prefix gives the same rows as the full batch
    raw_r, _ = eng.mlp_forward(m.desc, m.flat_params(), m.packed_weights(), m.precision, rays=rays[:1001].contiguous(), z=z_c[:1001].contiguous())
    assert torch.equal(raw_r, raw[:1001 * 64])
    # backward is linear in the upstream gradient, and the step is deterministic up to fp32 atomic order
    target = torch.rand(n, 3, device='cuda')
    nerf_process._counter[0] = 0
    trainer.render_losses_and_grads(net, rays, target, opts)
    g1 = net.model_fine.flat_grad.clone()
    nerf_process._counter[0] = 0
    trainer.render_losses_and_grads(net, rays, target, opts)
    g2 = net.model_fine.flat_grad.clone()
    assert float((g1 - g2).norm() / g1.norm()) <= 1e-5
    nerf_process._counter[0] = 0
    trainer.render_losses_and_grads(net, rays, target, opts, n_global=n // 2)    # loss scaled x2 -> gradient x2
    assert float((net.model_fine.flat_grad - 2 * g1).norm() / g1.norm()) <= 1e-3


def test_chunked_gradient_accumulation():
    """Batches too large for one activation stash are processed in ray chunks with gradient accumulation
    (BASELINE config 5 sizes); the result equals the single-pass step."""
    from nerf_pytorch_paeng_b200 import trainer
    from nerf_pytorch_paeng_b200.model import NeRF
    g = load_golden('raygen.npz')
    n = 900
    rs = np.random.RandomState(1)
    rays = cu(np.concatenate([np.tile(g['rays_o8'][:1], (n, 1)), g['rays_d8'][:n]], -1))
    target = cu(rs.rand(n, 3))
    rng = {'t_rand': cu(rs.rand(n, 64)), 'u': cu(rs.rand(n, 128))}
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).cuda().set_precision('fp32')
    res = []
    for max_pts in (10 ** 9, 256 * 192):          # single pass vs chunks of 256 rays (last chunk ragged: 132 rays)
        opts = make_opts(rng=dict(rng), max_points_per_pass=max_pts)
        out = trainer.render_losses_and_grads(net, rays, target, opts)
        torch.cuda.synchronize()
        res.append((out['loss_buf'].clone(), out['rgb_f'].clone(), net.model_coarse.flat_grad.clone(), net.model_fine.flat_grad.clone()))
    a, b = res
    assert float((a[0] - b[0]).abs().max()) <= 1e-6
    assert float((a[1] - b[1]).abs().max()) <= 1e-5 and b[1].shape == (n, 3)
    assert float((a[2] - b[2]).norm() / a[2].norm()) <= 1e-4
    assert float((a[3] - b[3]).norm() / a[3].norm()) <= 1e-4


@pytest.mark.parametrize('n_pts', [1, 129])
def test_tiny_batches_bf16(n_pts):
    """One point / one ragged extra tile through the cluster kernels (ghost tile in the CTA pair), forward and backward."""
    from nerf_pytorch_paeng_b200._lib import NB_BF16, NB_FP32
    from nerf_pytorch_paeng_b200.engine import get_engine
    from nerf_pytorch_paeng_b200.model import NeRF
    eng = get_engine(torch.device('cuda', 0))
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).cuda()
    m = net.model_fine
    rays = torch.tensor([[0., 0., 4., 0.1, -0.2, -1.0]], device='cuda')
    z = torch.linspace(2., 6., n_pts, device='cuda').reshape(1, n_pts)
    d_raw = torch.full((n_pts, 4), 1e-2, device='cuda')
    res = {}
    for prec in (NB_FP32, NB_BF16):
        m.precision = prec
        flat = m.flat_params()
        raw, act = eng.mlp_forward(m.desc, flat, m.packed_weights(), prec, rays=rays, z=z, save=True)
        grad = torch.full_like(flat, 3.0)
        eng.mlp_backward(m.desc, flat, m.packed_weights(), prec, n_pts, act, d_raw, grad)
        res[prec] = (raw.clone(), grad.clone())
    assert torch.isfinite(res[NB_BF16][0]).all() and torch.isfinite(res[NB_BF16][1]).all()
    assert float((res[NB_BF16][0] - res[NB_FP32][0]).abs().max()) <= 3e-2 * max(1., float(res[NB_FP32][0].abs().max()))
    assert float((res[NB_BF16][1] - res[NB_FP32][1]).norm() / res[NB_FP32][1].norm()) <= 0.2


def test_training_converges_bf16():
    """300 fused train steps (bf16 tcgen05 forward/backward + Adam, lr 5e-4) on a synthetic per-view colour target: the fine
    loss must fall at least 3x (measured: 0.103 -> 0.020) -- an end-to-end check that the gradients point the right way."""
    from nerf_pytorch_paeng_b200 import trainer
    from nerf_pytorch_paeng_b200.engine import get_engine
    from nerf_pytorch_paeng_b200.model import NeRF
    eng = get_engine(torch.device('cuda', 0))
    g = load_golden('raygen.npz')
    torch.manual_seed(0)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).cuda().set_precision('bf16')
    opt = trainer.FlatAdam(net, lr=5e-4)
    K = np.array([[1111.111, 0, 400.], [0, 1111.111, 400.], [0, 0, 1.]])
    poses = cu(g['all_poses'])
    opts = make_opts(N_rays=2048, seed=3)
    first = last = None
    for it in range(300):
        pose = poses[it % 32]
        pix = eng.select_pixels(2048, 800, 800, seed=100 + it % 32, offset=it * 2048)
        o, d = eng.raygen(800, 800, K, pose, pix_idx=pix)
        rays = torch.cat((o, d), -1)
        # synthetic "image": a smooth function of the pixel position and the camera
        r, c = (pix // 800).float() / 800., (pix % 800).float() / 800.
        tgt = torch.stack([0.5 + 0.4 * torch.sin(6.28 * r + pose[0, 3]), 0.5 + 0.4 * torch.cos(6.28 * c + pose[1, 3]), 0.3 + 0.4 * r * c], -1).contiguous()
        loss = trainer.train_step(net, opt, rays, tgt, opts)
        if it < 5:
            first = float(loss.sum()) if first is None else max(first, float(loss.sum()))
        if it >= 290:
            last = float(loss[1]) if last is None else min(last, float(loss[1]))
    assert np.isfinite(last) and last < (first / 2) / 3, (first, last)
    assert -10 * np.log10(last) > 15., last


@pytest.mark.parametrize('precision,perturb,s_f', [('bf16', 1., 128), ('fp32', 0., 128), ('bf16', 1., 0)])
def test_fused_drivers_equal_stepwise_calls(precision, perturb, s_f):
    """nb_train_rays / nb_render_rays (one C-ABI call) == the same stages enqueued one by one: identical renders (same kernels,
    same Philox counters), gradients equal up to the order of the fp32 atomics; split coarse/fine calls == one call."""
    from nerf_pytorch_paeng_b200 import nerf_process as NP, trainer
    from nerf_pytorch_paeng_b200.model import NeRF
    dev = torch.device('cuda', 0)
    torch.manual_seed(3)
    model = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev)
    model.set_precision(precision)
    n = 300
    rays = torch.cat([torch.randn(n, 3, device=dev) * .1 + torch.tensor([0., 0., 4.], device=dev),
                      torch.nn.functional.normalize(torch.randn(n, 3, device=dev) * .2 + torch.tensor([0., 0., -1.], device=dev), dim=-1)], -1)
    tgt = torch.rand(n, 3, device=dev)
    res = {}
    for mode in ('steps', 'fused', 'split'):
        opts = make_opts(perturb=perturb, N_samples_f=s_f, seed=11, fused_driver=mode != 'steps')
        NP._counter[0] = 0
        for net in (model.model_coarse, model.model_fine):
            net.bind_flat_grad().zero_()
        hooks = []
        out = trainer.render_losses_and_grads(model, rays, tgt, opts, n_global=2 * n,
                                              on_net_done=(lambda net: hooks.append(net)) if mode == 'split' else None)
        torch.cuda.synchronize()
        if mode == 'split':
            assert hooks == ([model.model_coarse, model.model_fine] if s_f else [model.model_coarse])
        res[mode] = ({k: v.clone() for k, v in out.items()}, model.model_coarse.flat_grad.clone(), model.model_fine.flat_grad.clone())
        NP._counter[0] = 0
        fr = trainer.render_rays_fused(model, rays, opts)
        NP._counter[0] = 0
        st = trainer.render_losses_and_grads_free(model, rays, opts)
        assert set(fr) == set(st)
        for k in fr:
            assert torch.equal(fr[k], st[k]), k
    ref_out, ref_gc, ref_gf = res['steps']
    for mode in ('fused', 'split'):
        o, gc, gf = res[mode]
        assert set(o) == set(ref_out)
        for k in ref_out:
            if k == 'loss_buf':
                assert torch.allclose(o[k], ref_out[k], rtol=1e-5, atol=1e-8)
            else:
                assert torch.equal(o[k], ref_out[k]), (mode, k)
        assert float((gc - ref_gc).norm() / ref_gc.norm()) < 1e-4
        if s_f:
            assert float((gf - ref_gf).norm() / ref_gf.norm()) < 1e-4
        else:
            assert float(gf.abs().max()) == 0.


def test_fused_driver_rejects_small_workspace():
    from nerf_pytorch_paeng_b200._lib import RenderCfg, NB_BF16
    from nerf_pytorch_paeng_b200.engine import NBError, get_engine, _ptr
    from nerf_pytorch_paeng_b200.model import NeRF
    import ctypes as C
    dev = torch.device('cuda', 0)
    eng = get_engine(dev)
    m = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev).model_coarse
    cfg = RenderCfg(64, 128, NB_BF16, 2, 0, 0, 0, 0, None, 0, 0)
    need = C.c_size_t()
    eng._call('nb_render_workspace_bytes', C.byref(m.desc), 256, C.byref(cfg), 0, C.byref(need))
    assert need.value > 256 * 192 * 16
    ws = torch.empty(1024, dtype=torch.uint8, device=dev)
    rays = torch.zeros(256, 6, device=dev)
    lo = torch.zeros(64, device=dev)
    with pytest.raises(NBError, match='workspace'):
        eng._call('nb_render_rays', C.byref(m.desc), C.byref(cfg), _ptr(m.flat_params()), None, _ptr(m.flat_params()), None, 256, _ptr(rays),
                  _ptr(lo), _ptr(lo), None, None, None, None, None, None, _ptr(ws), ws.numel(), eng.stream)


def test_data_parallel_path_world_size_1(tmp_path):
    """The data-parallel train step (joint gradient+loss buffer, one NCCL all-reduce) with a 1-rank process group equals
    the plain step: same loss, same parameters after Adam.  (2-rank equivalence: test_two_gpu_equals_one_gpu.)"""
    import copy
    import torch.distributed as dist
    from nerf_pytorch_paeng_b200 import nerf_process as NP, trainer
    from nerf_pytorch_paeng_b200.distributed import DistContext
    from nerf_pytorch_paeng_b200.model import NeRF
    dev = torch.device('cuda', 0)
    created = False
    if not dist.is_initialized():
        dist.init_process_group('nccl', init_method=f'file://{tmp_path}/pg', rank=0, world_size=1)
        created = True
    try:
        torch.manual_seed(5)
        m0 = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev)
        m0.set_precision('bf16')
        m1 = copy.deepcopy(m0)
        n = 256
        rays = torch.cat([torch.randn(n, 3, device=dev) * .1 + torch.tensor([0., 0., 4.], device=dev),
                          torch.nn.functional.normalize(torch.randn(n, 3, device=dev) * .2 + torch.tensor([0., 0., -1.], device=dev), dim=-1)], -1)
        tgt = torch.rand(n, 3, device=dev)
        res = []
        for model, ctx in ((m0, None), (m1, DistContext())):
            opts = make_opts(seed=3)
            NP._counter[0] = 0
            opt = trainer.FlatAdam(model, lr=5e-4)
            losses = [trainer.train_step(model, opt, rays, tgt, opts, dist_ctx=ctx).clone() for _ in range(1)]
            torch.cuda.synchronize()
            res.append((torch.stack(losses), model.model_coarse.flat_grad.clone(), model.model_fine.flat_grad.clone(),
                        model.model_coarse.flat_params().clone()))
        assert torch.allclose(res[0][0], res[1][0], rtol=1e-4, atol=1e-7)
        for k in (1, 2):        # gradients of the step (fp32 atomics reorder the sums)
            assert float((res[0][k] - res[1][k]).norm() / res[0][k].norm()) <= 1e-4
        # one Adam step moves every parameter by ~lr; identical up to the sign flips of ~zero gradients
        assert float((res[0][3] - res[1][3]).abs().median()) <= 1e-6
        whole, loss = DistContext().joint_grad_buffer(m1)
        assert whole.numel() == 2 * 595844 + 2 and m1.model_fine.flat_grad.data_ptr() == whole[595844:].data_ptr()
    finally:
        if created:
            dist.destroy_process_group()


def test_fused_driver_workspace_is_not_overrun():
    """nb_train_rays stays inside the nb_render_workspace_bytes it announces (guard bands), incl. the activation stash."""
    import ctypes as C
    from nerf_pytorch_paeng_b200 import nerf_process as NP
    from nerf_pytorch_paeng_b200._lib import RenderCfg, NB_BF16
    from nerf_pytorch_paeng_b200.engine import get_engine, _ptr
    from nerf_pytorch_paeng_b200.model import NeRF
    dev = torch.device('cuda', 0)
    eng = get_engine(dev)
    torch.manual_seed(1)
    model = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev)
    model.set_precision('bf16')
    nc, nf = model.model_coarse, model.model_fine
    n = 131                                                    # ragged: 131*64 and 131*192 points are not tile multiples
    rays = torch.cat([torch.zeros(n, 3, device=dev) + torch.tensor([0., 0., 4.], device=dev),
                      torch.nn.functional.normalize(torch.randn(n, 3, device=dev) * .2 + torch.tensor([0., 0., -1.], device=dev), dim=-1)], -1)
    tgt = torch.rand(n, 3, device=dev)
    opts = make_opts()
    lower, span = NP._coarse_bins(opts, dev)
    cfg = RenderCfg(64, 128, NB_BF16, 2, 9, 0, 100000, 0, None, 0, 0)
    need = C.c_size_t()
    eng._call('nb_render_workspace_bytes', C.byref(nc.desc), n, C.byref(cfg), 1, C.byref(need))
    G = 4096
    whole = torch.full((G + need.value + G,), 0x5A, dtype=torch.uint8, device=dev)
    ws = whole[G:G + need.value]
    gc, gf = torch.zeros_like(nc.flat_params()), torch.zeros_like(nf.flat_params())
    loss = torch.zeros(2, device=dev)
    rgb_f = torch.empty(n, 3, device=dev)
    eng._call('nb_train_rays', C.byref(nc.desc), C.byref(cfg), _ptr(nc.flat_params()), _ptr(nc.packed_weights()), _ptr(nf.flat_params()),
              _ptr(nf.packed_weights()), n, _ptr(rays), _ptr(tgt), None, n, _ptr(lower), _ptr(span), None, None, _ptr(gc), _ptr(gf), 0,
              _ptr(loss), None, None, _ptr(rgb_f), None, 3, _ptr(ws), ws.numel(), eng.stream)
    torch.cuda.synchronize()
    assert bool((whole[:G] == 0x5A).all()) and bool((whole[G + need.value:] == 0x5A).all()), 'workspace guard band overwritten'
    assert torch.isfinite(loss).all() and float(loss.min()) > 0 and torch.isfinite(gc).all() and torch.isfinite(gf).all()
    assert float(gc.abs().max()) > 0 and float(gf.abs().max()) > 0 and torch.isfinite(rgb_f).all()


def test_resume_from_reference_checkpoint():
    """f4: a checkpoint WRITTEN BY THE REFERENCE (train.py:105-114, torch.optim.Adam state, 3 real reference steps; fixture from
    oracle/make_golden.py::gen_checkpoint) is loaded the way main.py:111-117 resumes -- model.load_state_dict +
    optimizer.load_state_dict, here into trainer.FlatAdam -- and ONE resumed train step on the recorded inputs lands on the
    parameters the reference itself reached with its 4th step."""
    from nerf_pytorch_paeng_b200 import train as train_mod, trainer
    from nerf_pytorch_paeng_b200.model import NeRF
    from conftest import GOLDEN
    r = load_golden('ref_checkpoint_w64_resume.npz')
    net = NeRF(8, 64, 63, 27, [4], gt_camera_param=(None, None)).cuda().set_precision('fp32')
    opt = trainer.FlatAdam(net, lr=5e-4)
    idx = train_mod.load_checkpoint(os.path.join(GOLDEN, 'ref_checkpoint_w64_3.pth.tar'), net, opt)
    assert idx == 3 and opt.step_count == 3
    assert abs(opt.param_groups[0]['lr'] - float(r['lr_resume'])) < 1e-12        # the schedule position travels in param_groups
    before = {k: v.detach().clone() for k, v in net.state_dict().items()}
    opts = make_opts(rng={'t_rand': cu(r['in/t_rand']), 'u': cu(r['in/u'])})
    rays = cu(np.concatenate([r['in/rays_o'], r['in/rays_d']], -1))
    loss = trainer.train_step(net, opt, rays, cu(r['in/target']), opts)
    torch.cuda.synchronize()
    assert abs(float(loss.sum()) - float(r['in/loss'])) <= 1e-5
    assert opt.step_count == 4
    worst = moved = 0.
    for k, v in net.state_dict().items():
        ref = torch.from_numpy(r['a/' + k]).cuda()
        worst = max(worst, float((v - ref).abs().max()))
        moved = max(moved, float((ref - before[k]).abs().max()))
    assert moved > 1e-5                      # the step did something
    # ... and it is the reference's step: an Adam update is ~lr = 1.85e-4 per element; elements whose gradient is ~0 amplify the
    # 1e-4-relative difference between two fp32 implementations of the gradient (update = lr * m / (sqrt(v) + eps))
    assert worst <= 1e-5, worst
    num = sum(float(((v - torch.from_numpy(r['a/' + k]).cuda()) ** 2).sum()) for k, v in net.state_dict().items())
    den = sum(float(((torch.from_numpy(r['a/' + k]).cuda() - before[k]) ** 2).sum()) for k in before)
    assert (num / den) ** 0.5 <= 1e-2, (num / den) ** 0.5     # whole-vector error of the UPDATE itself
    # the optimizer state after the resumed step keeps torch.optim.Adam's layout
    sd = opt.state_dict()
    assert float(sd['state'][0]['step']) == 4.0 and sd['state'][47]['exp_avg'].shape == (3,)


def test_cuda_graph_train_step_equals_eager():
    """f2: the captured CUDA graph of the train step (device-resident Philox counter) computes the same losses and gradients
    as the eagerly enqueued step, replay after replay, and trains."""
    from nerf_pytorch_paeng_b200 import nerf_process as NP, trainer
    from nerf_pytorch_paeng_b200.model import NeRF
    g = load_golden('raygen.npz')
    n = 1024
    rs = np.random.RandomState(3)
    rays = cu(np.concatenate([np.tile(g['rays_o8'][:1], (n, 1)), g['rays_d8'][:n]], -1))
    targets = [cu(rs.rand(n, 3).astype(np.float32)) for _ in range(3)]
    opts = make_opts(seed=11)
    opts.cdf_order = 'cuda'

    def fresh():
        torch.manual_seed(0)
        net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).cuda().set_precision('bf16')
        return net, trainer.FlatAdam(net, lr=5e-4)
    # eager
    net_e, opt_e = fresh()
    NP.set_rng_state(0)
    eager = []
    for t in targets:
        loss = trainer.train_step(net_e, opt_e, rays, t, opts).clone()
        eager.append((loss, net_e.model_coarse.flat_grad.clone(), net_e.model_fine.flat_grad.clone()))
    # graphed
    net_g, opt_g = fresh()
    gs = trainer.GraphedTrainStep(net_g, opts, n, torch.device('cuda', 0)).capture()
    gs.ctr.zero_()                                    # the warm-up replays advanced the device counter
    eng = gs.eng
    for i, t in enumerate(targets):
        before = eng.launch_count()
        loss = gs(opt_g, rays, t).clone()
        torch.cuda.synchronize()
        l_e, gc_e, gf_e = eager[i]
        assert torch.allclose(loss, l_e, rtol=1e-5, atol=1e-7), (i, loss, l_e)
        for got, ref in ((net_g.model_coarse.flat_grad, gc_e), (net_g.model_fine.flat_grad, gf_e)):
            assert float((got - ref).norm() / ref.norm()) <= (1e-5 if i == 0 else 1e-2), i      # later steps: Adam's sign-like first updates amplify atomics-order noise
        assert eng.launch_count() - before <= 6       # host-side launches per step: 2 x (fold + re-pack) + 2 Adam (the rest is ONE graph launch)
    assert int(gs.ctr) == 3 * gs._per_step


def test_batch_cursor_fast_path_equals_reference_protocol():
    """utils.GetterRayBatchIdx: next_batch() (composed permutation + row gather, table never rewritten) walks the same batches as
    the reference's protocol (utils.py:41-58 + the slice of train.py:29) across several reshuffles."""
    from nerf_pytorch_paeng_b200.utils import GetterRayBatchIdx
    n, bs = 1000, 96
    table = torch.arange(n * 9, dtype=torch.float32, device='cuda').view(n, 3, 3)
    torch.manual_seed(5)
    a = GetterRayBatchIdx(table.clone())
    ref = []
    for _ in range(40):
        i, t, ep = a(bs)
        ref.append((t[i - bs:i].clone(), ep))
    torch.manual_seed(5)
    b = GetterRayBatchIdx(table.clone())
    for k in range(40):
        o, d, c = b.next_batch(bs)
        assert torch.equal(torch.stack([o, d, c], 1), ref[k][0]) and b.epoch == ref[k][1], k
    assert b.epoch == 3 and torch.equal(b.rays_rgb, table)          # three reshuffles, the table itself untouched


def test_adam_step_sum_equals_adam_on_the_summed_gradient():
    """nb_adam_step_sum (the data-parallel update with the all-reduce folded into its load) on ONE GPU: k gradient buffers summed
    in order == nb_adam_step on their sum, bit for bit; the sum is written back to g."""
    from nerf_pytorch_paeng_b200.engine import get_engine
    eng = get_engine(torch.device('cuda', 0))
    n = 595844
    gen = torch.Generator(device='cuda').manual_seed(1)
    p0 = torch.randn(n, device='cuda', generator=gen)
    gs = [torch.randn(n, device='cuda', generator=gen) * 1e-3 for _ in range(4)]
    m0, v0 = torch.rand(n, device='cuda', generator=gen) * 1e-3, torch.rand(n, device='cuda', generator=gen) * 1e-6
    total = ((gs[0] + gs[1]) + gs[2]) + gs[3]                    # rank order, like the kernel
    pa, ma, va = p0.clone(), m0.clone(), v0.clone()
    eng.adam_step(pa, total, ma, va, 3e-4, 7)
    pb, mb, vb = p0.clone(), m0.clone(), v0.clone()
    g_own = gs[1].clone()                                        # "this rank" is rank 1: its own buffer receives the sum
    eng.adam_step_sum(pb, g_own, [gs[0], g_own, gs[2], gs[3]], mb, vb, 3e-4, 7)
    torch.cuda.synchronize()
    assert torch.equal(g_own, total)
    assert torch.equal(pa, pb) and torch.equal(ma, mb) and torch.equal(va, vb)


def test_opts_precision_selects_the_mlp_path():
    """config.py --precision / opts.precision is honoured by the drop-in entry points (the flag was dead)."""
    from nerf_pytorch_paeng_b200 import config, nerf_process
    from nerf_pytorch_paeng_b200._lib import NB_BF16, NB_FP32
    from nerf_pytorch_paeng_b200.model import NeRF
    assert config.get_args_parser([]).precision == 'bf16'                        # the drop-in default is the tensor-core path
    g = load_golden('raygen.npz')
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).cuda()
    assert net.model_fine.precision == NB_FP32                                      # a bare module starts on the parity path
    rays = cu(np.concatenate([np.tile(g['rays_o8'][:1], (64, 1)), g['rays_d8'][:64]], -1))
    for prec, code in (('bf16', NB_BF16), ('fp32', NB_FP32)):
        opts = make_opts(precision=prec)
        with torch.no_grad():
            out = nerf_process.render_rays(rays, net, None, opts)
        assert net.model_coarse.precision == code and net.model_fine.precision == code
        assert bool(torch.isfinite(out['rgb_f']).all())
    opts = make_opts()                                                             # no attribute: the module's setting is left alone
    net.set_precision('bf16')
    with torch.no_grad():
        nerf_process.render_rays(rays, net, None, opts)
    assert net.model_fine.precision == NB_BF16


def test_bf16_mlp_calls_on_concurrent_streams():
    """The bf16 MLP entries stage per-network constants in __constant__ memory; a bank per stream in flight (nb_cbank.h) makes calls
    for DIFFERENT networks on different streams safe (one stream per device at a time).  Two, then six streams
    (more than there are banks: the least recently used one is handed over) run the coarse and the fine network interleaved; every
    result equals the single-stream result bit for bit."""
    from nerf_pytorch_paeng_b200.model import NeRF
    torch.manual_seed(3)
    net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).cuda()
    with torch.no_grad():
        for m in (net.model_coarse, net.model_fine):
            for lin in list(m.linear_x) + [m.linear_feat, m.linear_d, m.linear_density, m.linear_color]:
                lin.bias.uniform_(-0.5, 0.5)                               # the constants that would be mixed up
    net.set_precision('bf16')
    g = load_golden('raygen.npz')
    n, s = 1024, 192                                                        # 196,608 points: a few hundred microseconds per call
    rs = np.random.RandomState(0)
    rays = cu(np.concatenate([np.tile(g['rays_o8'][:1], (n, 1)), g['rays_d8'][:n]], -1))
    z = cu(np.sort(rs.rand(n, s).astype(np.float32) * 4 + 2, -1))
    mods = (net.model_coarse, net.model_fine)
    with torch.no_grad():
        want = [m.forward_rays(rays, z).clone() for m in mods]
        for m in mods:
            m.packed_weights()                                              # packed once, outside the streams
    torch.cuda.synchronize()
    assert not torch.equal(want[0], want[1])
    for n_streams in (2, 6):
        streams = [torch.cuda.Stream() for _ in range(n_streams)]
        got = []
        for it in range(4 * n_streams):
            k = it % n_streams
            with torch.cuda.stream(streams[k]), torch.no_grad():
                got.append(((it + it // n_streams) % 2, mods[(it + it // n_streams) % 2].forward_rays(rays, z)))
        torch.cuda.synchronize()
        for which, out in got:
            assert torch.equal(out, want[which]), (n_streams, which)

"""Generate tests/golden/*.npz by EXECUTING the unmodified reference on CPU.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The reference (a Python/PyTorch project) ships no tests or golden vectors, so
the known answers are manufactured here: each hot-path function of the
reference is imported from /root/reference (never copied), run on seeded
inputs with its RNG draws recorded, and inputs + outputs are written as small
fixtures.  Shims: a stub ``IQA_pytorch`` module (utils.py:3
imports it), ``torch.device -> cpu`` and ``Tensor.get_device -> 'cpu'`` so the
hard-coded ``cuda:N`` devices resolve on this GPU-less box.  torch.rand and
torch.searchsorted are wrapped (record only) so the random draws and the
bit-exact bin indices can be stored.

TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import types
from types import SimpleNamespace
from unittest import mock

import numpy as np
import torch

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def import_reference():
    stub = types.ModuleType('IQA_pytorch')
    stub.SSIM = object
    stub.LPIPSvgg = object
    sys.modules['IQA_pytorch'] = stub
    if 'cv2' not in sys.modules:
        try:
            import cv2  # noqa: F401
        except Exception:
            sys.modules['cv2'] = types.ModuleType('cv2')
    sys.path.insert(0, REF)
    import rays as ref_rays
    import nerf_process as ref_np
    from model import NeRF as RefNeRF, get_positional_encoder as ref_posenc
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_render_pose', os.path.join(REF, 'dataset', 'render_pose.py'))
    rp = importlib.util.module_from_spec(spec)   # dataset/__init__ pulls matplotlib/imageio (absent): load the file directly
    spec.loader.exec_module(rp)
    get_render_pose = rp.get_render_pose
    import scheduler as ref_sched
    return SimpleNamespace(rays=ref_rays, proc=ref_np, NeRF=RefNeRF, posenc=ref_posenc,
                           get_render_pose=get_render_pose, scheduler=ref_sched)


class Recorder:
    """Wraps torch.rand / torch.searchsorted: passes through, records results."""

    def __init__(self):
        self.rand = []
        self.inds = []
        self._rand = torch.rand
        self._ss = torch.searchsorted

    def rand_fn(self, *a, **k):
        k.pop('device', None)
        r = self._rand(*a, **k)
        self.rand.append(r.clone())
        return r

    def ss_fn(self, *a, **k):
        r = self._ss(*a, **k)
        self.inds.append(r.clone())
        return r


def cpu_patches(rec):
    real_device = torch.device
    return [
        mock.patch('torch.device', lambda *a, **k: real_device('cpu')),
        mock.patch('torch.Tensor.get_device', lambda self: 'cpu'),
        mock.patch('torch.rand', rec.rand_fn),
        mock.patch('torch.searchsorted', rec.ss_fn),
    ]


def blender_K(H, W, angle_x=0.6911112070083618):
    focal = .5 * W / np.tan(.5 * angle_x)
    return np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])  # float64 like load_blender.py:66-70


def llff_poses(n, seed=0):
    rng = np.random.RandomState(seed)
    poses = []
    for _ in range(n):
        ax, ay, az = rng.uniform(-0.05, 0.05, 3)
        Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
        Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
        Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
        t = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(-0.05, 0.05)])
        poses.append(np.concatenate([Rx @ Ry @ Rz, t[:, None]], 1))
    return np.stack(poses).astype(np.float32)


def state_to_np(module):
    return {k: v.detach().numpy().copy() for k, v in module.state_dict().items()}


def gen_render_train_w256(ref, rec):
    """The BENCHMARKED configuration pinned on the reference: full-width NeRF(8,256,63,27,[4]) at its seed-0
    PLAIN random init (and a variant with linear_density.weight x30 so that densities are not ~0), 1024 rays of an 800x800
    Blender-shaped view, 64+128 samples, perturb=1 with the draws recorded, MSE_c + MSE_f, backward.  Stores the render, the
    per-sample densities the two networks produced (for the last-sample sign-decision count, nerf_process.py:98), the losses and
    the WHOLE gradient vector of both networks.  Weights are not stored: they are the seed-0 init (param checksums are)."""
    H8 = W8 = 800
    K8 = blender_K(H8, W8)
    poses = ref.get_render_pose(n_angle=120, single_angle=-1, phi=-30.0, nf=4.0).numpy()
    pose8 = torch.from_numpy(poses[33][:3, :4].copy())
    o8, d8 = ref.rays.make_o_d(W8, H8, torch.from_numpy(K8), pose8)
    fx, _ = ref.posenc(10)
    fd, _ = ref.posenc(4)
    opts = SimpleNamespace(near=2., far=6., gpu_ids=[0], rank=0, N_samples_c=64, N_samples_f=128,
                           perturb=1., chunk_pts=524288, chunk_rays=4096, data_type='blender')
    N = 1024
    sel = np.random.RandomState(6).choice(H8 * W8, N, replace=False).astype(np.int64)
    ro, rd_ = o8.reshape(-1, 3)[sel].contiguous(), d8.reshape(-1, 3)[sel].contiguous()
    target = torch.rand(N, 3, generator=torch.Generator().manual_seed(7))
    for tag, scale in (('plain', 1.0), ('dens30', 30.0)):
        torch.manual_seed(0)
        net = ref.NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None))
        sd0 = net.state_dict()
        sums = np.array([float(v.double().sum()) for v in sd0.values()])
        if scale != 1.0:
            with torch.no_grad():
                for m in (net.model_coarse, net.model_fine):
                    m.linear_density.weight.mul_(scale)
        raws = []
        hooks = [m.register_forward_hook(lambda mod, inp, out: raws.append(out.detach().clone()))
                 for m in (net.model_coarse, net.model_fine)]
        torch.manual_seed(100)
        rec.rand.clear()
        net.zero_grad()
        rgb_c, disp_c, rgb_f, disp_f = ref.proc.batchify_rays_and_render_by_chunk(ro, rd_, net, [fx, fd], H8, W8, torch.from_numpy(K8), opts)
        for h in hooks:
            h.remove()
        t_rand, u = rec.rand[0], rec.rand[1]
        assert len(raws) == 2 and raws[0].shape == (N * 64, 4) and raws[1].shape == (N * 192, 4)
        crit = torch.nn.MSELoss()
        loss_c, loss_f = crit(rgb_c, target), crit(rgb_f, target)
        (loss_c + loss_f).backward()
        gc = torch.cat([p.grad.reshape(-1) for p in net.model_coarse.parameters()]).numpy()
        gf = torch.cat([p.grad.reshape(-1) for p in net.model_fine.parameters()]).numpy()
        np.savez_compressed(os.path.join(OUT, f'render_train_w256_{tag}.npz'), rays_o=ro.numpy(), rays_d=rd_.numpy(), target=target.numpy(),
                            t_rand=t_rand.numpy(), u=u.numpy(), near=2., far=6., density_scale=scale,
                            rgb_c=rgb_c.detach().numpy(), disp_c=disp_c.detach().numpy(), rgb_f=rgb_f.detach().numpy(),
                            disp_f=disp_f.detach().numpy(), loss_c=float(loss_c), loss_f=float(loss_f),
                            sigma_c=raws[0][:, 3].reshape(N, 64).numpy(), sigma_f=raws[1][:, 3].reshape(N, 192).numpy(),
                            grad_coarse=gc, grad_fine=gf, param_sums_seed0=sums, param_names=np.array(list(sd0.keys())))


def gen_checkpoint(ref, rec):
    """A checkpoint WRITTEN BY THE REFERENCE's own code path (train.py:105-114: {'idx', 'model_state_dict', 'optimizer_state_dict'}
    of the reference NeRF module and torch.optim.Adam after 3 real reference train steps; small W=64 net so the file stays small),
    plus what the reference computes when it RESUMES from it (main.py:111-117) and takes one more step on recorded inputs."""
    H8 = W8 = 800
    K8 = blender_K(H8, W8)
    poses = ref.get_render_pose(n_angle=120, single_angle=-1, phi=-30.0, nf=4.0).numpy()
    o8, d8 = ref.rays.make_o_d(W8, H8, torch.from_numpy(K8), torch.from_numpy(poses[5][:3, :4].copy()))
    fx, _ = ref.posenc(10)
    fd, _ = ref.posenc(4)
    opts = SimpleNamespace(near=2., far=6., gpu_ids=[0], rank=0, N_samples_c=64, N_samples_f=128,
                           perturb=1., chunk_pts=524288, chunk_rays=4096, data_type='blender')
    N = 64
    torch.manual_seed(3)
    net = ref.NeRF(8, 64, 63, 27, [4], gt_camera_param=(None, None))
    with torch.no_grad():
        for m in (net.model_coarse, net.model_fine):
            m.linear_density.weight.mul_(30.)
    opt = torch.optim.Adam(net.parameters(), lr=5e-4, betas=(0.9, 0.999))                    # main.py:79-80
    sched = ref.scheduler.CosineAnnealingWarmupRestarts(opt, first_cycle_steps=1000, cycle_mult=1., max_lr=5e-4, min_lr=5e-5,
                                                        warmup_steps=10) if hasattr(ref, 'scheduler') else None
    crit = torch.nn.MSELoss()

    def step(i):
        sel = np.random.RandomState(50 + i).choice(H8 * W8, N, replace=False)
        ro, rd_ = o8.reshape(-1, 3)[sel].contiguous(), d8.reshape(-1, 3)[sel].contiguous()
        target = torch.rand(N, 3, generator=torch.Generator().manual_seed(60 + i))
        rec.rand.clear()
        rgb_c, _, rgb_f, _ = ref.proc.batchify_rays_and_render_by_chunk(ro, rd_, net, [fx, fd], H8, W8, torch.from_numpy(K8), opts)
        opt.zero_grad()
        loss = crit(rgb_c, target) + crit(rgb_f, target)
        loss.backward()
        opt.step()
        if sched is not None:
            sched.step()
        return dict(rays_o=ro.numpy(), rays_d=rd_.numpy(), target=target.numpy(), t_rand=rec.rand[0].numpy(), u=rec.rand[1].numpy(),
                    loss=float(loss), lr=float(opt.param_groups[0]['lr']))
    for i in range(3):
        step(i)
    ckpt = {'idx': 3, 'model_state_dict': net.state_dict(), 'optimizer_state_dict': opt.state_dict()}      # train.py:107-109
    torch.save(ckpt, os.path.join(OUT, 'ref_checkpoint_w64_3.pth.tar'))
    lr_resume = float(opt.param_groups[0]['lr'])
    rec4 = step(3)                                                                                         # the resumed step
    np.savez_compressed(os.path.join(OUT, 'ref_checkpoint_w64_resume.npz'), lr_resume=lr_resume,
                        **{'in/' + k: v for k, v in rec4.items()}, **{'a/' + k: v for k, v in state_to_np(net).items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = import_reference()
    import torch._dynamo  # noqa: F401  (optimizer construction imports it lazily; must happen before torch.device is patched)
    rec = Recorder()
    patches = cpu_patches(rec)
    for p in patches:
        p.start()
    only = [a for a in sys.argv[1:] if not a.startswith('-')]
    try:
        torch.set_num_threads(8)
        if only:                       # e.g. `python oracle/make_golden.py w256 ckpt`: regenerate just these fixtures
            if 'w256' in only:
                gen_render_train_w256(ref, rec)
            if 'ckpt' in only:
                gen_checkpoint(ref, rec)
            return
        gen_render_train_w256(ref, rec)
        gen_checkpoint(ref, rec)
        torch.manual_seed(0)
        np.random.seed(0)
        torch.set_num_threads(8)

        # ---------------- K1: make_o_d / get_rays_np (blender-shaped) ----------------
        poses = ref.get_render_pose(n_angle=120, single_angle=-1, phi=-30.0, nf=4.0).numpy()
        H, W = 60, 80
        K = blender_K(H, W)
        pose = torch.from_numpy(poses[7][:3, :4].copy())
        o, d = ref.rays.make_o_d(W, H, torch.from_numpy(K), pose)          # tensor-K (train.py:43)
        o2, d2 = ref.rays.make_o_d(W, H, K, pose)                          # numpy-K (test.py:38)
        assert torch.equal(d, d2)
        on, dn = ref.rays.get_rays_np(H, W, K, poses[7][:3, :4])
        # full-size 800x800, store a strided subset of pixels
        H8 = W8 = 800
        K8 = blender_K(H8, W8)
        pose8 = torch.from_numpy(poses[33][:3, :4].copy())
        o8, d8 = ref.rays.make_o_d(W8, H8, torch.from_numpy(K8), pose8)
        sel = np.random.RandomState(1).choice(H8 * W8, 4096, replace=False).astype(np.int64)
        d8s = d8.reshape(-1, 3)[sel].numpy()
        o8s = o8.reshape(-1, 3)[sel].numpy()
        np.savez_compressed(
            os.path.join(OUT, 'raygen.npz'),
            H=H, W=W, K=K, pose=pose.numpy(), rays_o=o.numpy().copy(), rays_d=d.numpy(),
            np_rays_o=np.ascontiguousarray(on), np_rays_d=dn, np_dtype=str(dn.dtype), numpy_version=np.__version__,
            H8=H8, W8=W8, K8=K8, pose8=pose8.numpy(), sel8=sel, rays_o8=o8s, rays_d8=d8s,
            all_poses=poses[:, :3, :4].copy())

        # ---------------- K1: ndc_rays (llff-shaped) ----------------
        Hl, Wl, focal = 756, 1008, 815.13158
        Kl = np.array([[focal, 0, .5 * Wl], [0, focal, .5 * Hl], [0, 0, 1]])
        lp = llff_poses(4)
        ol, dl = ref.rays.make_o_d(Wl, Hl, torch.from_numpy(Kl), torch.from_numpy(lp[1]))
        sel_l = np.random.RandomState(2).choice(Hl * Wl, 4096, replace=False).astype(np.int64)
        ol_s = ol.reshape(-1, 3)[sel_l].contiguous()
        dl_s = dl.reshape(-1, 3)[sel_l].contiguous()
        on_t, dn_t = ref.proc.ndc_rays(Hl, Wl, torch.from_numpy(Kl)[0][0], 1., ol_s, dl_s)   # tensor focal
        on_n, dn_n = ref.proc.ndc_rays(Hl, Wl, Kl[0][0], 1., ol_s, dl_s)                      # numpy focal
        assert torch.equal(on_t, on_n) and torch.equal(dn_t, dn_n)
        np.savez_compressed(os.path.join(OUT, 'ndc.npz'), H=Hl, W=Wl, focal=focal, K=Kl, pose=lp[1], sel=sel_l,
                            rays_o=ol_s.numpy(), rays_d=dl_s.numpy(), ndc_o=on_t.numpy(), ndc_d=dn_t.numpy(),
                            llff_poses=lp)

        # ---------------- K3: positional encoding ----------------
        x = (torch.rand(512, 3) * 8 - 4)
        fx, dx = ref.posenc(10)
        fd, dd = ref.posenc(4)
        vd = torch.nn.functional.normalize(torch.randn(512, 3), dim=-1)
        np.savez_compressed(os.path.join(OUT, 'posenc.npz'), x=x.numpy(), enc_x=fx(x).numpy(), out_dim_x=dx,
                            d=vd.numpy(), enc_d=fd(vd).numpy(), out_dim_d=dd)

        # ---------------- K2 coarse + embedded input: pre_process(isFine=False) ----------------
        opts = SimpleNamespace(near=2., far=6., gpu_ids=[0], rank=0, N_samples_c=64, N_samples_f=128,
                               perturb=1., chunk_pts=524288, chunk_rays=4096, data_type='blender')
        N = 24
        sel_r = np.random.RandomState(3).choice(H8 * W8, N, replace=False)
        rays = torch.cat([o8.reshape(-1, 3)[sel_r], d8.reshape(-1, 3)[sel_r]], -1).contiguous()
        rec.rand.clear()
        emb, z_c, rd = ref.proc.pre_process(rays, [fx, fd], opts, isFine=False)
        t_rand = rec.rand[-1]
        t_vals = torch.linspace(0., 1., steps=64)
        np.savez_compressed(os.path.join(OUT, 'pre_process_coarse.npz'), rays=rays.numpy(), t_rand=t_rand.numpy(),
                            near=opts.near, far=opts.far, embedded=emb.numpy(), z_vals=z_c.numpy(),
                            t_vals=t_vals.numpy(), t_vals_128=torch.linspace(0., 1., steps=128).numpy(),
                            t_vals_192=torch.linspace(0., 1., steps=192).numpy())

        # ---------------- K2 fine: sample_pdf det / random, + full pre_process(isFine=True) ----------------
        Ns = 320
        g = torch.Generator().manual_seed(5)
        z = torch.sort(torch.rand(Ns, 64, generator=g) * 4 + 2, -1)[0]
        # peaky weights like a real transmittance profile, plus rows of exact zeros (empty rays)
        w = torch.rand(Ns, 64, generator=g) ** 8
        w[:16] = 0.
        w[16:32, 10:50] = 0.
        mids = .5 * (z[..., 1:] + z[..., :-1])
        opts_det = SimpleNamespace(**{**vars(opts), 'perturb': 0.})
        rec.inds.clear()
        s_det = ref.proc.sample_pdf(mids, w[..., 1:-1], 128, det=True, opts=opts_det)
        inds_det = rec.inds[-1]
        rec.rand.clear()
        rec.inds.clear()
        s_rnd = ref.proc.sample_pdf(mids, w[..., 1:-1], 128, det=False, opts=opts)
        u_rnd = rec.rand[-1]
        inds_rnd = rec.inds[-1]
        # the reference's cdf for the same weights (for the from-cdf bit-exact test)
        ww = w[..., 1:-1] + 1e-5
        pdf = ww / torch.sum(ww, -1, keepdim=True)
        cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
        # fine pre_process end to end (sort + embed) on a few rays
        rays_f = rays[:8].contiguous()
        rec.rand.clear()
        emb_f, z_f, _ = ref.proc.pre_process(rays_f, [fx, fd], opts, z_vals=z[:8].contiguous(), weights=w[40:48].contiguous(), isFine=True)
        u_f = rec.rand[-1]
        np.savez_compressed(os.path.join(OUT, 'sample_pdf.npz'), z_vals=z.numpy(), weights=w.numpy(), bins=mids.numpy(),
                            u_det=torch.linspace(0., 1., steps=128).numpy(), samples_det=s_det.numpy(),
                            inds_det=inds_det.numpy(), u_rnd=u_rnd.numpy(), samples_rnd=s_rnd.numpy(),
                            inds_rnd=inds_rnd.numpy(), cdf=cdf.numpy(),
                            fine_rays=rays_f.numpy(), fine_z_in=z[:8].numpy(), fine_w_in=w[40:48].numpy(), fine_u=u_f.numpy(),
                            fine_z=z_f.numpy(), fine_embedded=emb_f.numpy())

        # ---------------- K5: post_process incl. edge cases ----------------
        for S in (64, 192):
            Np = 160
            g = torch.Generator().manual_seed(10 + S)
            raw = torch.randn(Np, S, 4, generator=g)
            raw[..., 3] = raw[..., 3] * 3.
            zz = torch.sort(torch.rand(Np, S, generator=g) * 4 + 2, -1)[0]
            dd_ = torch.randn(Np, 3, generator=g)
            raw[0:8, :, 3] = -1.             # empty rays: acc=0 -> disp NaN->0 (last alpha = 0)
            raw[8:16, :, 3] = 50.            # opaque at first sample
            raw[16:24, :, 3] = 1e-3          # thin: disp clamp / last-sample alpha=1
            raw[24:32, :-1, 3] = -1.         # everything on the last (1e10) sample
            raw[24:32, -1, 3] = 1.
            zz[32:40] = zz[32:40] * 0.01     # tiny depths -> disp > 5 clamp
            zz[40:48, 10:20] = zz[40:48, 10:11]   # repeated depths: zero dists
            raw = raw.requires_grad_(True)
            outs = ref.proc.post_process(raw, zz, dd_)
            gup = torch.randn(Np, 3, generator=g)
            (outs[0] * gup).sum().backward()
            np.savez_compressed(os.path.join(OUT, f'post_process_S{S}.npz'), raw=raw.detach().numpy(), z_vals=zz.numpy(),
                                rays_d=dd_.numpy(), rgb_map=outs[0].detach().numpy(), disp_map=outs[1].detach().numpy(),
                                acc_map=outs[2].detach().numpy(), weights=outs[3].detach().numpy(),
                                depth_map=outs[4].detach().numpy(), d_rgb=gup.numpy(), d_raw=raw.grad.numpy())

        # ---------------- K4: MLP (small width stored whole; full width via seeded init) ----------------
        torch.manual_seed(0)
        net_s = ref.NeRF(8, 64, 63, 27, [4], gt_camera_param=(None, None))
        xs = torch.cat([fx(x[:256]), fd(vd[:256])], -1)
        ys_c = net_s(xs)
        ys_f = net_s(xs, is_fine=True)
        gy = torch.randn(256, 4)
        net_s.zero_grad()
        (ys_c * gy).sum().backward()
        grads_c = {k: v.grad.numpy().copy() for k, v in net_s.model_coarse.named_parameters()}
        sd = state_to_np(net_s)
        np.savez_compressed(os.path.join(OUT, 'mlp_w64.npz'), x=xs.numpy(), y_coarse=ys_c.detach().numpy(),
                            y_fine=ys_f.detach().numpy(), d_y=gy.numpy(),
                            **{'p/' + k: v for k, v in sd.items()}, **{'g/' + k: v for k, v in grads_c.items()})

        torch.manual_seed(0)
        net = ref.NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None))
        with torch.no_grad():
            y_c = net(xs)
            y_f = net(xs, is_fine=True)
        sd = net.state_dict()
        np.savez_compressed(os.path.join(OUT, 'mlp_w256_seed0.npz'), x=xs.numpy(), y_coarse=y_c.numpy(), y_fine=y_f.numpy(),
                            param_names=np.array(list(sd.keys())),
                            param_sums=np.array([float(v.double().sum()) for v in sd.values()]),
                            param_abs_sums=np.array([float(v.double().abs().sum()) for v in sd.values()]),
                            param_first=np.array([float(v.flatten()[0]) for v in sd.values()]),
                            n_params=sum(v.numel() for v in sd.values()))

        # ---------------- render_rays end to end + train-step grads (W=64 net, stored whole) ----------------
        Nr = 48
        sel_r = np.random.RandomState(4).choice(H8 * W8, Nr, replace=False)
        ro, rd_ = o8.reshape(-1, 3)[sel_r].contiguous(), d8.reshape(-1, 3)[sel_r].contiguous()
        target = torch.rand(Nr, 3)
        # scale the last layers so densities are not ~0 (otherwise the pdf is flat and the test is weak)
        with torch.no_grad():
            for m in (net_s.model_coarse, net_s.model_fine):
                m.linear_density.weight.mul_(30.)
                m.linear_color.weight.mul_(4.)
        sd = state_to_np(net_s)
        rec.rand.clear()
        net_s.zero_grad()
        rgb_c, disp_c, rgb_f, disp_f = ref.proc.batchify_rays_and_render_by_chunk(ro, rd_, net_s, [fx, fd], H8, W8, torch.from_numpy(K8), opts)
        t_rand_r, u_r = rec.rand[0], rec.rand[1]
        crit = torch.nn.MSELoss()
        loss_c, loss_f = crit(rgb_c, target), crit(rgb_f, target)
        (loss_c + loss_f).backward()
        grads = {k: v.grad.numpy().copy() for k, v in net_s.named_parameters()}
        # one Adam step exactly as main.py:79-80 / train.py:70
        opt = torch.optim.Adam(net_s.parameters(), lr=5e-4, betas=(0.9, 0.999))
        opt.step()
        sd_after = state_to_np(net_s)
        np.savez_compressed(os.path.join(OUT, 'render_train_w64.npz'), rays_o=ro.numpy(), rays_d=rd_.numpy(), target=target.numpy(),
                            t_rand=t_rand_r.numpy(), u=u_r.numpy(), near=2., far=6.,
                            rgb_c=rgb_c.detach().numpy(), disp_c=disp_c.detach().numpy(), rgb_f=rgb_f.detach().numpy(),
                            disp_f=disp_f.detach().numpy(), loss_c=float(loss_c), loss_f=float(loss_f), lr=5e-4,
                            **{'p/' + k: v for k, v in sd.items()}, **{'g/' + k: v for k, v in grads.items()},
                            **{'a/' + k: v for k, v in sd_after.items()})

        # ---------------- llff end to end (NDC inside batchify), deterministic u (perturb=0.) ----------------
        opts_l = SimpleNamespace(near=0., far=1., gpu_ids=[0], rank=0, N_samples_c=64, N_samples_f=128,
                                 perturb=0., chunk_pts=524288, chunk_rays=4096, data_type='llff')
        Nl = 32
        rec.rand.clear()
        with torch.no_grad():
            lc, ldc, lf, ldf = ref.proc.batchify_rays_and_render_by_chunk(ol_s[:Nl].contiguous(), dl_s[:Nl].contiguous(), net_s, [fx, fd],
                                                                          Hl, Wl, torch.from_numpy(Kl), opts_l)
        assert len(rec.rand) == 1
        np.savez_compressed(os.path.join(OUT, 'render_llff_w64.npz'), rays_o=ol_s[:Nl].numpy(), rays_d=dl_s[:Nl].numpy(),
                            H=Hl, W=Wl, focal=focal, t_rand=rec.rand[0].numpy(), u_det=torch.linspace(0., 1., steps=128).numpy(),
                            rgb_c=lc.numpy(), disp_c=ldc.numpy(), rgb_f=lf.numpy(), disp_f=ldf.numpy(),
                            **{'a/' + k: v for k, v in sd_after.items()})
    finally:
        for p in patches:
            p.stop()
    tot = 0
    for f in sorted(os.listdir(OUT)):
        sz = os.path.getsize(os.path.join(OUT, f))
        tot += sz
        print(f'{f:32s} {sz / 1024:9.1f} KiB')
    print(f'total {tot / 1024:.1f} KiB')


if __name__ == '__main__':
    main()

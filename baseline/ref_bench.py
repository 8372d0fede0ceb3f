"""Timing of the UNMODIFIED reference (baseline/_ref, imported through ref_shim) on this box: its CPU path on the host cores
(BASELINE configs[0] / SURVEY 8(d): 1024 rays, 64+128, fp32, random-init) and its own eager-PyTorch CUDA path on the GPU
(configs[1] 4096-ray train step, configs[2] 800x800 render).  BASELINE INFRASTRUCTURE ONLY: nothing under
nerf_pytorch_paeng_b200/ imports this; bench.py uses it for `--impl reference`, `cpu_baseline` and `gpu_baseline`.

The reference's train.py cannot be imported (visdom / matplotlib / configargparse are absent), so one train step is its
body restated over the reference's OWN functions: rays.make_o_d -> rays.sample_rays_and_pixel ->
nerf_process.batchify_rays_and_render_by_chunk -> nn.MSELoss x2 -> backward -> torch.optim.Adam.step (train.py:35-70,
main.py:79-80)."""
import os
import time
from types import SimpleNamespace

import numpy as np
import torch

from . import ref_shim

H = W = 800
FOCAL = 0.5 * 800 / np.tan(0.5 * 0.6911112070083618)            # load_blender.py:51-52


def make_opts(n_rays, device_index=0, **kw):
    base = dict(near=2., far=6., gpu_ids=[device_index], rank=0, N_samples_c=64, N_samples_f=128, perturb=1., chunk_pts=524288,
                chunk_rays=4096, data_type='blender', N_rays=n_rays, precrop_iters=0, precrop_frac=.5)
    base.update(kw)
    return SimpleNamespace(**base)


def setup(device, n_rays, seed=0):
    """Reference model (random init under torch.manual_seed(seed)), encoders, Adam, synthetic Blender-shaped inputs."""
    ref = ref_shim.import_reference()
    torch.manual_seed(seed)
    np.random.seed(seed)
    model = ref.NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(device)
    fx, _ = ref.posenc(10)
    fd, _ = ref.posenc(4)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, betas=(0.9, 0.999))
    K = np.array([[FOCAL, 0, 0.5 * W], [0, FOCAL, 0.5 * H], [0, 0, 1]])
    poses = ref.get_render_pose(n_angle=120, single_angle=-1, phi=-30.0, nf=4.0)
    images = torch.rand(4, H, W, 3)
    return SimpleNamespace(ref=ref, model=model, posenc=[fx, fd], opt=opt, K=K, poses=poses, images=images,
                           opts=make_opts(n_rays, device.index or 0), device=device, crit=torch.nn.MSELoss())


def train_step(s, i):
    """train.py:35-70 (per-image path) on the reference's functions."""
    dev = s.device
    i_img = i % s.images.shape[0]
    target_img = s.images[i_img].to(dev)                                                   # train.py:37-38
    pose = s.poses[i % s.poses.shape[0], :3, :4].to(dev)
    rays_o, rays_d = s.ref.rays.make_o_d(W, H, torch.from_numpy(s.K).to(dev), pose)        # train.py:43
    rays_o, rays_d, target = s.ref.rays.sample_rays_and_pixel(i, rays_o, rays_d, target_img, s.opts)
    rgb_c, disp_c, rgb_f, disp_f = s.ref.proc.batchify_rays_and_render_by_chunk(rays_o, rays_d, s.model, s.posenc, H, W, s.K, s.opts)
    s.opt.zero_grad()
    loss = s.crit(rgb_c, target) + s.crit(rgb_f, target)
    loss.backward()
    s.opt.step()
    return loss


def render_rays_only(s, n_rays, pose_idx=0):
    """configs[0]: no_grad render of the first n_rays pixel rays of a seeded permutation of one 800x800 view."""
    dev = s.device
    pose = s.poses[pose_idx, :3, :4].to(dev)
    rays_o, rays_d = s.ref.rays.make_o_d(W, H, torch.from_numpy(s.K).to(dev), pose)
    sel = torch.from_numpy(np.random.RandomState(0).permutation(H * W)[:n_rays]).to(dev)
    o, d = rays_o.reshape(-1, 3)[sel], rays_d.reshape(-1, 3)[sel]
    with torch.no_grad():
        return s.ref.proc.batchify_rays_and_render_by_chunk(o, d, s.model, s.posenc, H, W, s.K, s.opts)


def cpu_info():
    model = ''
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                model = line.split(':', 1)[1].strip()
                break
    except OSError:
        pass
    return model


def time_cpu(n_rays=1024, steps=3, warmup=1, threads=None):
    """The reference's CPU path on this host, all cores.  torchrun sets OMP_NUM_THREADS=1 for nproc>1: the thread count is
    set explicitly here and reported.  Returns dict(train_rays_per_s, render_rays_per_s, threads, ...)."""
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    dev = torch.device('cpu')
    with ref_shim.on_cpu():
        s = setup(dev, n_rays)
        for i in range(warmup):
            train_step(s, i)
        ts = []
        for i in range(steps):
            t0 = time.perf_counter()
            train_step(s, warmup + i)
            ts.append(time.perf_counter() - t0)
        render_rays_only(s, n_rays)
        tr = []
        for _ in range(max(1, min(steps, 3))):
            t0 = time.perf_counter()
            render_rays_only(s, n_rays)
            tr.append(time.perf_counter() - t0)
    return {'train_rays_per_s': n_rays / float(np.mean(ts)), 'train_ms_per_step': 1e3 * float(np.mean(ts)), 'train_best_ms': 1e3 * min(ts),
            'render_rays_per_s': n_rays / min(tr), 'render_ms': 1e3 * min(tr), 'n_rays': n_rays, 'steps': steps, 'warmup': warmup,
            'threads': torch.get_num_threads(), 'cpu_count': os.cpu_count(), 'cpu_model': cpu_info()}


def time_gpu(device, n_rays=4096, steps=10, warmup=3, render_frames=1):
    """The reference's own eager CUDA path on this GPU (like-for-like baseline): 4096-ray train steps and 800x800 renders."""
    s = setup(device, n_rays)
    for i in range(warmup):
        train_step(s, i)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        loss = train_step(s, warmup + i)
    e1.record()
    torch.cuda.synchronize(device)
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / steps
    out = {'train_rays_per_s': n_rays / (ms / 1e3), 'train_ms_per_step': ms, 'train_wall_ms_per_step': 1e3 * wall / steps, 'n_rays': n_rays,
           'steps': steps, 'warmup': warmup, 'loss': float(loss), 'peak_mem_GB': torch.cuda.max_memory_allocated(device) / 1e9,
           'what': 'unmodified reference, PyTorch eager fp32 (TF32 off), per-image path incl. its host pixel selection'}
    if render_frames > 0:
        pose = s.poses[0, :3, :4].to(device)
        with torch.no_grad():
            def frame():
                rays_o, rays_d = s.ref.rays.make_o_d(W, H, s.K, pose)                          # test.py:38
                return s.ref.proc.batchify_rays_and_render_by_chunk(rays_o, rays_d, s.model, s.posenc, H, W, s.K, s.opts)
            frame()
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            for _ in range(render_frames):
                frame()
            torch.cuda.synchronize(device)
            fr = (time.perf_counter() - t0) / render_frames
        out.update({'render_frames_per_s': 1.0 / fr, 'render_ms_per_frame': 1e3 * fr, 'render_frames': render_frames})
    return out

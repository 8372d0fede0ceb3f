#!/usr/bin/env python
"""Benchmark of the ray-batch hot path (BASELINE.json: train rays/s at 4096 rays, 64+128 samples).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]

One "step" = one full train step (ray batch -> coarse+fine render -> MSE_c+MSE_f -> backward ->
Adam) on 4096 rays PER GPU (weak scaling) of the Blender-lego-shaped synthetic workload
(BASELINE.json configs[1]).  Prints ONE JSON line (rank 0).  `value` times the step with the ray
batch already resident in HBM; `e2e` times the reference-facing call train.train(...) with the
target image and pixel indices coming from pinned HOST memory every step and the loss read back.
`--impl reference` times the CPU restatement of the reference (oracle/, numpy+BLAS on all host
cores) on a bounded sample of the same workload; under torchrun only rank 0 runs it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS, S_C, S_F = 4096, 64, 128
H = W = 800
FOCAL = 0.5 * 800 / np.tan(0.5 * 0.6911112070083618)        # load_blender.py:51-52 -> 1111.111
FLOP_PER_POINT_TRAIN = 3489024                              # fwd 1,186,816 + bwd 2,302,208
FLOP_PER_POINT_FWD = 1186816
POINTS_PER_RAY = S_C + (S_C + S_F)                          # coarse net sees 64, fine net all 192
WORKLOAD = 'Blender lego-shaped train step: 4096 rays/batch per GPU, 64+128 samples, PE L=10/4, 8x256 skip MLP x2, Adam'
WORKLOAD_LLFF = ('LLFF fern-shaped train step (BASELINE configs[3]): NDC rays at 1008x756, 4096 rays/batch per GPU, 64+128 samples, '
                 'near 0 / far 1, Adam')


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', 1373.4), d.get('hbm_gbs', 6549.8), 'measured (MEASURED_PEAKS.json, sustained bf16)'
    return 1400.0, 6650.0, 'fallback (B200_PROFILING.md)'


def synthetic_poses(n, seed=0):
    """Blender-shaped c2w poses on a radius-4 sphere (the generator of dataset/render_pose.py:28-34,
    theta~U(-180,180), phi~U(-90,0)), restated with numpy."""
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        th, phi = np.deg2rad(rs.uniform(-180, 180)), np.deg2rad(rs.uniform(-90, 0))
        t = np.eye(4); t[2, 3] = 4.0
        rp = np.array([[1, 0, 0, 0], [0, np.cos(phi), -np.sin(phi), 0], [0, np.sin(phi), np.cos(phi), 0], [0, 0, 0, 1]])
        rt = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]])
        c2w = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]) @ rt @ rp @ t
        out.append(c2w)
    return np.stack(out).astype(np.float32)


def llff_poses(n, seed=0):
    """Forward-facing LLFF-shaped c2w poses: identity +- U(-0.3,0.3) translation in x,y and +-0.05 rad rotations."""
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        ax, ay, az = rs.uniform(-0.05, 0.05, 3)
        rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
        ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
        rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
        t = np.array([rs.uniform(-0.3, 0.3), rs.uniform(-0.3, 0.3), rs.uniform(-0.05, 0.05)])
        out.append(np.concatenate([np.concatenate([rx @ ry @ rz, t[:, None]], 1), [[0, 0, 0, 1]]], 0))
    return np.stack(out).astype(np.float32)


def make_opts(rank_dev=0, **kw):
    base = dict(near=2., far=6., N_samples_c=S_C, N_samples_f=S_F, perturb=1., data_type='blender', gpu_ids=[rank_dev], rank=0,
                chunk_rays=N_RAYS, chunk_pts=524288, N_rays=N_RAYS, precrop_iters=0, precrop_frac=.5, seed=0,
                global_batch=False, idx_print=10 ** 9, idx_save=None, exp_name='bench')
    base.update(kw)
    return SimpleNamespace(**base)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20',
                                          '-i', str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the numpy oracle on all host cores, bounded sample
# ----------------------------------------------------------------------------------------------
def cpu_port_step(n_rays, seed=0):
    """One train step (render + grads + Adam) of the oracle port on n_rays rays; returns seconds."""
    from oracle import nerf_oracle as orc
    rs = np.random.RandomState(seed)
    if not hasattr(cpu_port_step, 'state'):
        import torch
        torch.manual_seed(0)
        from nerf_pytorch_paeng_b200.model import NeRF           # host-side module: only for the reference's seeded init
        net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None))
        sd = {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}
        pc = {k[len('model_coarse.'):]: v for k, v in sd.items() if k.startswith('model_coarse.')}
        pf = {k[len('model_fine.'):]: v for k, v in sd.items() if k.startswith('model_fine.')}
        pose = synthetic_poses(1)[0]
        K = np.array([[FOCAL, 0, 400.], [0, FOCAL, 400.], [0, 0, 1.]])
        o, d = orc.make_o_d(W, H, K, pose[:3, :4])
        cpu_port_step.state = (pc, pf, o.reshape(-1, 3), d.reshape(-1, 3))
    pc, pf, o, d = cpu_port_step.state
    sel = rs.choice(H * W, n_rays, replace=False)
    rays = np.concatenate([o[sel], d[sel]], -1)
    target = rs.rand(n_rays, 3).astype(np.float32)
    t_rand, u = rs.rand(n_rays, S_C).astype(np.float32), rs.rand(n_rays, S_F).astype(np.float32)
    t0 = time.perf_counter()
    lc, lf, gc, gf = orc.train_grads(rays, target, pc, pf, orc.make_opts(), t_rand, u)
    for p, g in ((pc, gc), (pf, gf)):
        for k in p:
            p[k], _, _ = orc.adam_step(p[k], g[k], np.zeros_like(g[k]), np.zeros_like(g[k]), 1, 5e-4)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = args.cpu_rays
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_port_step(n)
    times = [cpu_port_step(n, seed=i + 1) for i in range(max(1, min(args.steps, 5)))]
    ms = 1e3 * float(np.mean(times))
    val = n / (ms / 1e3)
    cores = os.cpu_count()
    line = {'impl': 'reference', 'metric': 'train_rays_per_s', 'value': val, 'unit': 'rays/s', 'n_gpus': args.gpus,
            'steps': len(times), 'warmup': max(1, min(args.warmup, 2)), 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'sample': f'{n} rays per step (bounded sample of the 4096-ray batch)'},
            'cpu_baseline': {'value': val, 'unit': 'rays/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{n}-ray train step (render+grads+Adam) of oracle/nerf_oracle.py, numpy/BLAS threads={cores}'},
            'e2e': {'value': val, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    emit(line)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from nerf_pytorch_paeng_b200 import distributed, train as train_mod, trainer
    from nerf_pytorch_paeng_b200.engine import get_engine
    from nerf_pytorch_paeng_b200.model import NeRF, get_positional_encoder

    dctx = distributed.init_from_env('nccl')
    rank = dctx.rank if dctx else 0
    world = dctx.world_size if dctx else 1
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    eng = get_engine(dev)
    torch.manual_seed(0)
    model = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev)      # random init, identical on every rank
    model.set_precision(args.precision)
    global H, W
    llff = args.workload == 'llff'
    if llff:
        H, W = 756, 1008
        focal = 815.13158
        opts = make_opts(rank_dev=local, seed=1000 + rank, device_select=True, data_type='llff', near=0., far=1.)
        K = np.array([[focal, 0, .5 * W], [0, focal, .5 * H], [0, 0, 1.]])
        poses = llff_poses(16, seed=0)
    else:
        opts = make_opts(rank_dev=local, seed=1000 + rank, device_select=True)
        K = np.array([[FOCAL, 0, 400.], [0, FOCAL, 400.], [0, 0, 1.]])
        poses = synthetic_poses(16, seed=0)
    optimizer = trainer.FlatAdam(model, lr=5e-4)
    posenc = [get_positional_encoder(10)[0], get_positional_encoder(4)[0]]
    poses_dev = torch.from_numpy(poses).to(dev)

    # ---- resident inputs: a ring of pre-generated ray batches (different pose/pixels per step, per rank)
    n_ring = 8
    gen = torch.Generator(device='cpu').manual_seed(1234 + rank)
    ring = []
    for i in range(n_ring):
        pix = torch.randperm(H * W, generator=gen)[:N_RAYS].to(dev)
        o, d = eng.raygen(H, W, K, poses_dev[i % len(poses), :3, :4], pix_idx=pix, ndc=llff, ndc_focal=float(K[0][0]), ndc_near=1.)
        ring.append((torch.cat((o, d), -1), torch.rand(N_RAYS, 3, device=dev)))

    def barrier():
        if dctx:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step(i):
        rays, tgt = ring[i % n_ring]
        return trainer.train_step(model, optimizer, rays, tgt, opts, dist_ctx=dctx)

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        loss = step(i)
    ev1.record()
    barrier()
    launches = eng.launch_count() - l0
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device=dev)
    if dctx:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_step = float(t) / args.steps
    value = N_RAYS * world / (ms_step / 1e3)

    # ---- roofline: CUDA events (on the launching stream) around each of the three MLP kernels during a few extra steps.
    # engine.mlp_forward == one launch of mlp_fwd_chain_kernel<train>; the backward is issued as its two ABI stages so that
    # mlp_dgrad_chain_kernel and mlp_wgrad_kernel are timed separately.  The DOMINANT kernel by launch-list share
    # (profiles/*launches*.csv) is mlp_wgrad_kernel, which is HBM-bound: it is the headline `roofline`.
    of, ob = eng.mlp_forward, eng.mlp_backward
    pend = {'fwd': [], 'dgrad': [], 'wgrad': []}

    def ev_pair(tag, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        pend[tag].append((e0, e1))
        return r

    def fwd_timed(*a, **k):
        return ev_pair('fwd', lambda: of(*a, **k))

    def bwd_timed(*a, **k):
        if args.precision != 'bf16':
            return ev_pair('wgrad', lambda: ob(*a, **k))
        ev_pair('dgrad', lambda: ob(*a, stage=1, **k))
        return ev_pair('wgrad', lambda: ob(*a, stage=2, **k))
    n_prof = min(args.steps, 5)
    opts.fused_driver = False         # same kernels, enqueued stage by stage so that events can be placed between them
    step(0)                           # untimed: lets the caching allocator create the stage-by-stage buffers
    torch.cuda.synchronize()
    eng.mlp_forward, eng.mlp_backward = fwd_timed, bwd_timed
    for i in range(n_prof):
        step(i)
    torch.cuda.synchronize()
    opts.fused_driver = True
    eng.mlp_forward, eng.mlp_backward = of, ob
    ms = {k: [a.elapsed_time(b) for a, b in v] for k, v in pend.items()}
    avg = {k: (sum(v) / len(v) if v else 0.0) for k, v in ms.items()}
    mlp_ms = sum(sum(v) for v in ms.values()) / n_prof
    peak_tf, peak_hbm, peak_src = peaks()
    pts_per_launch = N_RAYS * POINTS_PER_RAY / 2.0                       # two launches per step: 64- and 192-sample nets
    traffic = {}
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    flop_step = FLOP_PER_POINT_TRAIN * N_RAYS * POINTS_PER_RAY
    fwd_tf = FLOP_PER_POINT_FWD * pts_per_launch / (avg['fwd'] / 1e3) / 1e12
    # wgrad: algorithmic HBM bytes = every 16 KB operand blob its 12 jobs read: 85 blobs per 128-point tile
    WGRAD_BYTES_PER_POINT = 85 * 16384 / 128.0
    other = {'mlp_fwd_chain_kernel<train>': {'bound': 'tensor', 'achieved': fwd_tf, 'unit': 'TFLOP/s', 'frac': fwd_tf / peak_tf,
                                             'avg_launch_ms': avg['fwd'], 'traffic': traffic.get('mlp_fwd_chain_kernel_bytes_per_launch')}}
    if args.precision == 'bf16':
        wg_gbs = WGRAD_BYTES_PER_POINT * pts_per_launch / (avg['wgrad'] / 1e3) / 1e9
        dg_tf = 1115392 * pts_per_launch / (avg['dgrad'] / 1e3) / 1e12       # dgrad: 557,696 MAC per point
        other['mlp_dgrad_chain_kernel'] = {'bound': 'tensor', 'achieved': dg_tf, 'unit': 'TFLOP/s', 'frac': dg_tf / peak_tf,
                                           'avg_launch_ms': avg['dgrad'], 'traffic': traffic.get('mlp_dgrad_chain_kernel_bytes_per_launch')}
        roofline = {'bound': 'hbm', 'achieved': wg_gbs, 'peak': peak_hbm, 'unit': 'GB/s', 'frac': wg_gbs / peak_hbm,
                    'traffic': traffic.get('mlp_wgrad_kernel_bytes_per_launch'),
                    'kernel': 'mlp_wgrad_kernel (tcgen05 weight-gradient GEMMs, MN-major operands streamed from the activation stash)',
                    'algorithmic_bytes_per_launch': WGRAD_BYTES_PER_POINT * pts_per_launch, 'avg_launch_ms': avg['wgrad'],
                    'launches_per_step': 2, 'peak_source': peak_src.replace('sustained bf16', 'hbm_gbs')}
    else:
        roofline = dict(other.pop('mlp_fwd_chain_kernel<train>'), peak=peak_tf, kernel='sgemm_kernel chain (fp32 CUDA-core parity path)',
                        peak_source=peak_src)
    roofline['other_kernels'] = other
    roofline['whole_step'] = {'algorithmic_flop_per_step': flop_step, 'mlp_ms_per_step': mlp_ms, 'mlp_share_of_step': mlp_ms / ms_step,
                              'achieved_tflops': flop_step / (mlp_ms / 1e3) / 1e12, 'frac_of_bf16_peak': flop_step / (mlp_ms / 1e3) / 1e12 / peak_tf}
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: the reference-facing call train.train(...) with HOST inputs every step
    n_img = 4
    images = [torch.rand(H, W, 3).pin_memory() for _ in range(n_img)]
    gt_cam = (K, poses[:n_img])
    crit = torch.nn.MSELoss()
    np.random.seed(rank)
    host_loss = torch.zeros(1).pin_memory()

    def e2e_step(i):
        loss = train_mod.train(i + 1, list(range(n_img)), images, gt_cam, (H, W), model, crit, posenc, optimizer, None, None, opts,
                               dist_ctx=dctx)
        host_loss.copy_(loss.reshape(1), non_blocking=False)      # D2H read of the step's result
        return float(host_loss)
    for i in range(max(3, args.warmup // 2)):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], device=dev)
    if dctx:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_ms_step = float(t) / args.steps
    e2e = {'value': N_RAYS * world / (e2e_ms_step / 1e3), 'unit': 'rays/s', 'ms_per_step': e2e_ms_step,
           'h2d_bytes_per_step': H * W * 3 * 4, 'd2h_bytes_per_step': 4,
           'api': 'nerf_pytorch_paeng_b200.train.train (per-image path, train.py:35-45: pinned host image -> H2D every step, '
                  'pixel selection + ray-gen + target gather + fused step + Adam on device, loss D2H)'}

    # ---- second half of BASELINE.json's metric: full 800x800 coarse+fine frames/s, pixel bands sharded over ranks
    render = None
    if not args.no_render:
        pose = poses_dev[0, :3, :4]
        for _ in range(2):
            trainer.render_frame(model, H, W, K, pose, opts, dist_ctx=dctx)
        barrier()
        n_fr = 3
        t0 = time.perf_counter()
        for i in range(n_fr):
            rgb, disp = trainer.render_frame(model, H, W, K, poses_dev[i % len(poses), :3, :4], opts, dist_ctx=dctx)
            frame8 = eng.frame_to8b(rgb, disp)[0].cpu()                      # to8b on device + D2H of the finished frame
        barrier()
        fr_ms = (time.perf_counter() - t0) * 1e3 / n_fr
        t = torch.tensor([fr_ms], device=dev)
        if dctx:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        fr_ms = float(t)
        fl = FLOP_PER_POINT_FWD * H * W * POINTS_PER_RAY
        render = {'frames_per_s': 1e3 / fr_ms, 'ms_per_frame': fr_ms, 'rays_per_s': H * W / (fr_ms / 1e3), 'frame': f'{W}x{H} coarse+fine, 64+128',
                  'achieved_tflops': fl / (fr_ms / 1e3) / 1e12 / 1.0, 'frac_of_peak_all_gpus': fl / (fr_ms / 1e3) / 1e12 / (peak_tf * world),
                  'includes': 'ray-gen, sampling, MLP x2, compositing, band all-gather, uint8 frame D2H'}
    if rank != 0:
        return
    # ---- cpu baseline on this box's host cores (bounded sample)
    cpu = None
    if not args.no_cpu:
        cpu_port_step(args.cpu_rays)
        ts = [cpu_port_step(args.cpu_rays, seed=i + 1) for i in range(2)]
        cpu = {'value': args.cpu_rays / float(np.mean(ts)), 'unit': 'rays/s', 'cores': os.cpu_count(), 'kind': 'port',
               'sample': f'{args.cpu_rays}-ray train step (render+grads+Adam) of oracle/nerf_oracle.py, numpy/BLAS on all host cores'}
    line = {'metric': 'train_rays_per_s', 'value': value, 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD_LLFF if llff else WORKLOAD, 'rays_per_gpu': N_RAYS, 'global_rays': N_RAYS * world, 'samples': [S_C, S_F],
                       'precision': args.precision, 'parallelism': f'ray-sharded data parallel x{world}, one NCCL all-reduce of 1,191,690 fp32 (gradients of both nets + the two losses)',
                       'l2': 'per-step working set (activation stash >= 5 GB) exceeds the 126 MB L2; ring of 8 distinct ray batches',
                       'loss': float(loss.sum())},
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'render': render, 'gpu_launches': int(launches), 'clocks': clocks,
            'gpu_launches_per_step': launches / args.steps}
    emit(line)


def _claim_stdout():
    """Keep stdout for the single JSON line: everything else this process (or NCCL, which prints its version banner
    on stdout) writes to fd 1 is sent to stderr; returns a file object on the original stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    return real


_REAL_STDOUT = None


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + '\n')
    _REAL_STDOUT.flush()


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', type=str, default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', type=str, default=os.environ.get('NB_PRECISION', 'bf16'), choices=['fp32', 'bf16'])
    ap.add_argument('--cpu-rays', dest='cpu_rays', type=int, default=256)
    ap.add_argument('--no-cpu', dest='no_cpu', action='store_true')
    ap.add_argument('--no-render', dest='no_render', action='store_true')
    ap.add_argument('--workload', type=str, default='blender', choices=['blender', 'llff'],
                    help='blender = BASELINE configs[1] (the headline); llff = configs[3] (NDC rays at 1008x756)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

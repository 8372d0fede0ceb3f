#!/usr/bin/env python
"""Benchmark of the ray-batch hot path (BASELINE.json: train rays/s at 4096 rays, 64+128 samples).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]

One "step" = one full train step (ray batch -> coarse+fine render -> MSE_c+MSE_f -> backward ->
Adam) on 4096 rays PER GPU (weak scaling) of the Blender-lego-shaped synthetic workload
(BASELINE.json configs[1]).  Prints ONE JSON line (rank 0).  `value` times the step with the ray
batch already resident in HBM; `e2e` times the reference-facing call train.train(...) driven from
the host every step (camera pose from pinned host memory, on-device pixel selection + ray
generation + target gather from the device-resident image stack, loss copied back to the host).
The line also carries `roofline` (tensor pipe of the MLP kernels, HBM figures beside it),
`cpu_baseline` (the UNMODIFIED reference, baseline/_ref, on this box's host cores: BASELINE
configs[0], 1024 rays) and `gpu_baseline` (the unmodified reference's own eager-PyTorch CUDA path on
this GPU: the like-for-like figure); for N > 1 `dp_check` (sum of the rank gradients against the
single-GPU gradient of the concatenated batch) and `strong_scaling` (4096/N rays per GPU).
`--impl reference` times the reference's CPU path (all host threads, thread count set explicitly
because torchrun exports OMP_NUM_THREADS=1) on a bounded sample of the workload; under torchrun
only rank 0 runs it.  oracle/ is used only as the fallback CPU port when baseline/_ref is absent.
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


def cpu_reference(n_rays, steps, warmup):
    """CPU baseline on this box: the unmodified reference (kind 'reference') when baseline/_ref travelled here, else the numpy
    port in oracle/ (kind 'port').  Returns the cpu_baseline object (value = train rays/s)."""
    from baseline import ref_shim
    cores = os.cpu_count()
    if ref_shim.available():
        from baseline import ref_bench
        r = ref_bench.time_cpu(n_rays=n_rays, steps=steps, warmup=warmup, threads=cores)
        return {'value': r['train_rays_per_s'], 'unit': 'rays/s', 'cores': r['threads'], 'kind': 'reference',
                'sample': f'{n_rays}-ray train step (make_o_d + sample_rays_and_pixel + render 64+128 + MSE_c+MSE_f + backward + Adam) of the '
                          f'unmodified reference on CPU, mean of {steps} after {warmup} warm-up (BASELINE configs[0] shape)',
                'ms_per_step': r['train_ms_per_step'], 'render_rays_per_s': r['render_rays_per_s'], 'cpu_count': r['cpu_count'],
                'cpu_model': r['cpu_model'], 'torch_threads': r['threads']}
    for _ in range(max(1, warmup)):
        cpu_port_step(min(n_rays, 256))
    ts = [cpu_port_step(min(n_rays, 256), seed=i + 1) for i in range(max(1, steps))]
    return {'value': min(n_rays, 256) / float(np.mean(ts)), 'unit': 'rays/s', 'cores': cores, 'kind': 'port',
            'sample': f'{min(n_rays, 256)}-ray train step of oracle/nerf_oracle.py (numpy/BLAS); baseline/_ref not installed',
            'ms_per_step': 1e3 * float(np.mean(ts))}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 20)), max(1, min(args.warmup, 2))
    cpu = cpu_reference(args.cpu_rays, steps, warmup)
    line = {'impl': 'reference', 'metric': 'train_rays_per_s', 'value': cpu['value'], 'unit': 'rays/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': warmup, 'ms_per_step': cpu['ms_per_step'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'sample': f'{args.cpu_rays} rays per step (bounded sample of the 4096-ray batch; the reference\'s CPU path, '
                                                       'a reported baseline: the like-for-like GPU figure is gpu_baseline of the other arm)'},
            'cpu_baseline': cpu,
            'e2e': {'value': cpu['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
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

    # the step as a captured CUDA graph (render + loss + backward [+ all-reduce]; Adam and the weight re-pack follow
    # as ordinary launches).  Measured equal to the eagerly enqueued step within noise (the step is GPU-bound): `value` times the eager step, `--graph`
    # times the replay; at N = 1 the replay time is reported beside it as `cuda_graph`.
    graphed = None
    if args.graph and args.precision == 'bf16':
        graphed = trainer.GraphedTrainStep(model, opts, N_RAYS, dev, dist_ctx=dctx).capture()

    def step_eager(i):
        rays, tgt = ring[i % n_ring]
        return trainer.train_step(model, optimizer, rays, tgt, opts, dist_ctx=dctx)

    if graphed is not None:
        def step(i):                                                   # noqa: F811
            rays, tgt = ring[i % n_ring]
            return graphed(optimizer, rays, tgt)
    for i in range(args.warmup):
        step(i)
    barrier()
    # the eagerly enqueued step, for comparison (same kernels, 18 launches + memsets/copies from the host per step)
    eager_ms = None
    if graphed is not None:
        for i in range(3):
            step_eager(i)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(args.steps):
            step_eager(i)
        g1.record()
        barrier()
        tt = torch.tensor([g0.elapsed_time(g1)], device=dev)
        if dctx:
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        eager_ms = float(tt) / args.steps
        for i in range(3):
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
    launches = eng.launch_count() - l0                      # launches issued through the C ABI from the host in the timed region
    if graphed is not None:                                 # + the library's kernels inside each graph replay (counted at capture)
        launches += graphed.kernels_per_replay * args.steps
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device=dev)
    if dctx:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_step = float(t) / args.steps
    value = N_RAYS * world / (ms_step / 1e3)

    # ---- f2 evidence at N = 1: the same step replayed as a captured CUDA graph
    graph_info = None
    if graphed is None and dctx is None and args.precision == 'bf16':
        gs = trainer.GraphedTrainStep(model, opts, N_RAYS, dev).capture()
        for i in range(3):
            gs(optimizer, *ring[i % n_ring])
        torch.cuda.synchronize()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for i in range(args.steps):
            gs(optimizer, *ring[i % n_ring])
        q1.record()
        torch.cuda.synchronize()
        graph_info = {'ms_per_step': q0.elapsed_time(q1) / args.steps, 'kernels_per_replay': int(gs.kernels_per_replay), 'host_launches_per_step': 6,
                      'what': 'trainer.GraphedTrainStep: render + loss + backward captured once, Philox counters on the device; Adam x2 and (weight fold + re-pack) x2 from the host'}
        del gs

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
    step_eager(0)                     # untimed: lets the caching allocator create the stage-by-stage buffers
    torch.cuda.synchronize()
    eng.mlp_forward, eng.mlp_backward = fwd_timed, bwd_timed
    for i in range(n_prof):
        step_eager(i)
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
    tsrc = traffic.get('source', 'profiles/traffic.json (ncu --set full dram__bytes of an earlier run of this command; not measured in this run)')
    flop_step = FLOP_PER_POINT_TRAIN * N_RAYS * POINTS_PER_RAY
    fwd_tf = FLOP_PER_POINT_FWD * pts_per_launch / (avg['fwd'] / 1e3) / 1e12
    # wgrad's algorithmic HBM bytes: every 16 KB operand blob its 11 jobs read: 78 blobs per 128-point tile
    WGRAD_BYTES_PER_POINT = 78 * 16384 / 128.0
    other = {'mlp_fwd_chain_kernel<train>': {'bound': 'tensor', 'achieved': fwd_tf, 'unit': 'TFLOP/s', 'frac': fwd_tf / peak_tf,
                                             'avg_launch_ms': avg['fwd'], 'traffic': traffic.get('mlp_fwd_chain_kernel_bytes_per_launch')}}
    if args.precision == 'bf16':
        wg_gbs = WGRAD_BYTES_PER_POINT * pts_per_launch / (avg['wgrad'] / 1e3) / 1e9
        wg_tf = FLOP_PER_POINT_FWD * pts_per_launch / (avg['wgrad'] / 1e3) / 1e12    # dW = dY^T X: the forward's 593,408 MAC per point
        dg_tf = 1115392 * pts_per_launch / (avg['dgrad'] / 1e3) / 1e12               # dgrad: 557,696 MAC per point
        other['mlp_dgrad_chain_kernel'] = {'bound': 'tensor', 'achieved': dg_tf, 'unit': 'TFLOP/s', 'frac': dg_tf / peak_tf,
                                           'avg_launch_ms': avg['dgrad'], 'traffic': traffic.get('mlp_dgrad_chain_kernel_bytes_per_launch')}
        other['mlp_wgrad_kernel as an HBM stream'] = {
            'bound': 'hbm', 'achieved': wg_gbs, 'peak': peak_hbm, 'unit': 'GB/s', 'frac': wg_gbs / peak_hbm,
            'algorithmic_bytes_per_launch': WGRAD_BYTES_PER_POINT * pts_per_launch,
            'note': 'the kernel streams 9.98 KB/point of stashed operands; this is the roofline that binds it in practice (DESIGN.md section 4)'}
        # headline: the dominant kernel by time share (profiles/*launches*.csv) against SURVEY 8(d)'s bound for K4, the tensor pipe
        roofline = {'bound': 'tensor', 'achieved': wg_tf, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': wg_tf / peak_tf,
                    'traffic': traffic.get('mlp_wgrad_kernel_bytes_per_launch'), 'traffic_source': tsrc,
                    'kernel': 'mlp_wgrad_kernel (tcgen05 weight-gradient GEMMs, MN-major operands streamed from the activation stash)',
                    'algorithmic_flop_per_launch': FLOP_PER_POINT_FWD * pts_per_launch, 'avg_launch_ms': avg['wgrad'],
                    'launches_per_step': 2, 'peak_source': peak_src}
    else:
        roofline = dict(other.pop('mlp_fwd_chain_kernel<train>'), peak=peak_tf, kernel='sgemm_kernel chain (fp32 CUDA-core parity path)',
                        peak_source=peak_src)
    roofline['executed_vs_algorithmic'] = {
        'note': 'achieved = SURVEY 8(d) algorithmic FLOP / time.  The bf16 path folds the activation-free feature layer into the view layer '
                '(W\' = Wd[:, :256] . Wf, DESIGN.md section 4), so the tensor cores execute fewer MACs per point than the reference\'s op sequence',
        'algorithmic_mac_per_point': {'fwd': 593408, 'dgrad': 557696, 'wgrad': 593408},
        'executed_mac_per_point': {'fwd': 527872, 'dgrad': 492160, 'wgrad': 527872}}
    roofline['other_kernels'] = other
    roofline['whole_step'] = {'algorithmic_flop_per_step': flop_step, 'mlp_ms_per_step': mlp_ms, 'mlp_share_of_step': mlp_ms / ms_step,
                              'achieved_tflops': flop_step / (mlp_ms / 1e3) / 1e12, 'frac_of_bf16_peak': flop_step / (mlp_ms / 1e3) / 1e12 / peak_tf,
                              'whole_step_incl_everything_tflops': flop_step / (ms_step / 1e3) / 1e12,
                              'whole_step_frac_of_bf16_peak': flop_step / (ms_step / 1e3) / 1e12 / peak_tf}
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: the reference-facing call train.train(...) driven from the host every step.  The training images are a
    # device-resident stack (uploaded once); what crosses PCIe per step is the selected camera pose (pinned host
    # memory -> device) and the step's loss (device -> pinned host memory, consumed two steps later so that the host never
    # blocks the queue).  `per_step_image_upload` repeats the measurement with the reference's own behaviour (train.py:37-38:
    # the whole 7.68 MB image uploaded every step) for comparison.
    n_img = 4
    images = [torch.rand(H, W, 3).pin_memory() for _ in range(n_img)]
    poses_pin = torch.from_numpy(poses[:n_img]).pin_memory()
    gt_cam = (K, poses_pin)
    crit = torch.nn.MSELoss()
    np.random.seed(rank)
    host_loss = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_ev = [None, None]

    def run_e2e(cache_images):
        opts.cache_images, opts.cache_poses = cache_images, False
        seen = []

        def e2e_step(i):
            slot = i % 2
            if loss_ev[slot] is not None:
                loss_ev[slot].synchronize()
                seen.append(float(host_loss[slot]))                    # the loss of step i-2, read on the host
            loss = train_mod.train(i + 1, list(range(n_img)), images, gt_cam, (H, W), model, crit, posenc, optimizer, None, None, opts,
                                   dist_ctx=dctx)
            host_loss[slot].copy_(loss.detach().reshape(1), non_blocking=True)   # D2H read of the step's result
            loss_ev[slot] = torch.cuda.Event()
            loss_ev[slot].record()
        for i in range(max(4, args.warmup)):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(i)
        barrier()
        ms = (time.perf_counter() - t0) * 1e3
        tt = torch.tensor([ms], device=dev)
        if dctx:
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        assert len(seen) >= args.steps and np.isfinite(seen).all()
        return float(tt) / args.steps
    e2e_ms_step = run_e2e(True)
    e2e_upload_ms = run_e2e(False)
    opts.cache_images = opts.cache_poses = True
    e2e = {'value': N_RAYS * world / (e2e_ms_step / 1e3), 'unit': 'rays/s', 'ms_per_step': e2e_ms_step,
           'h2d_bytes_per_step': 48, 'd2h_bytes_per_step': 4,
           'api': 'nerf_pytorch_paeng_b200.train.train (per-image path, train.py:35-45): pose from pinned host memory every step, on-device '
                  'pixel selection + ray-gen + target gather from the device-resident image stack + fused step + Adam, loss D2H every step',
           'per_step_image_upload': {'value': N_RAYS * world / (e2e_upload_ms / 1e3), 'ms_per_step': e2e_upload_ms,
                                     'h2d_bytes_per_step': H * W * 3 * 4 + 48, 'note': 'opts.cache_images=False: the reference\'s per-step upload of the whole image'}}

    # ---- N > 1: (a) the all-reduced gradient of a sharded batch against the single-GPU gradient of the concatenated batch
    # (bf16 joint-buffer path, injected draws), (b) strong scaling: the same 4096 rays split over the ranks
    dp_check = strong = None
    if dctx is not None:
        gg = torch.Generator(device='cpu').manual_seed(4242)
        n_all = 1024 * world
        pix_all = torch.randperm(H * W, generator=gg)[:n_all].to(dev)
        o, d = eng.raygen(H, W, K, poses_dev[3, :3, :4], pix_idx=pix_all, ndc=llff, ndc_focal=float(K[0][0]), ndc_near=1.)
        rays_all = torch.cat((o, d), -1)
        tgt_all = torch.rand(n_all, 3, generator=gg).to(dev)
        tr_all, u_all = torch.rand(n_all, S_C, generator=gg).to(dev), torch.rand(n_all, S_F, generator=gg).to(dev)
        lo, hi = rank * 1024, (rank + 1) * 1024
        from types import SimpleNamespace as NS
        o_sh = NS(**{**vars(opts), 'rng': {'t_rand': tr_all[lo:hi].contiguous(), 'u': u_all[lo:hi].contiguous()}})
        whole, lossv = dctx.joint_grad_buffer(model)
        lossv.zero_()
        trainer.render_losses_and_grads(model, rays_all[lo:hi].contiguous(), tgt_all[lo:hi].contiguous(), o_sh, n_global=n_all, loss_buf=lossv)
        dctx.allreduce_(whole)
        summed = whole.clone()
        o_one = NS(**{**vars(opts), 'rng': {'t_rand': tr_all, 'u': u_all}})
        lossv.zero_()
        trainer.render_losses_and_grads(model, rays_all, tgt_all, o_one, n_global=n_all, loss_buf=lossv)     # every rank: the whole batch alone
        torch.cuda.synchronize()
        ng = whole.numel() - 2
        nc = model.model_coarse.flat_params().numel()
        rel = lambda a, b: float((a - b).norm() / b.norm())
        dp_check = {'what': f'sum over {world} ranks of the gradients of 1024-ray shards (bf16, joint buffer, one all-reduce) vs the single-GPU '
                            f'gradient of the concatenated {n_all}-ray batch, identical injected draws',
                    'grad_rel_err_coarse': rel(summed[:nc], whole[:nc]), 'grad_rel_err_fine': rel(summed[nc:ng], whole[nc:ng]),
                    'loss_abs_err': float((summed[ng:] - whole[ng:]).abs().max())}
        # strong scaling: global batch stays 4096 rays
        n_loc = N_RAYS // world
        sring = [(r[rank * n_loc:(rank + 1) * n_loc].contiguous(), t[rank * n_loc:(rank + 1) * n_loc].contiguous()) for r, t in ring]
        for i in range(3):
            trainer.train_step(model, optimizer, *sring[i % n_ring], opts, dist_ctx=dctx)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(args.steps):
            trainer.train_step(model, optimizer, *sring[i % n_ring], opts, dist_ctx=dctx)
        s1.record()
        barrier()
        tt = torch.tensor([s0.elapsed_time(s1)], device=dev)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        sms = float(tt) / args.steps
        strong = {'scaling': 'strong', 'global_rays': N_RAYS, 'rays_per_gpu': n_loc, 'ms_per_step': sms, 'value': N_RAYS / (sms / 1e3), 'unit': 'rays/s'}
        # at ~1 ms per step the host's 22 launches matter: the same step as ONE captured CUDA graph (render + loss + backward +
        # the NCCL all-reduce of the joint buffer), Adam and the re-pack following from the host
        if args.precision == 'bf16':
            try:
                gs = trainer.GraphedTrainStep(model, opts, n_loc, dev, dist_ctx=dctx).capture()
                for i in range(3):
                    gs(optimizer, *sring[i % n_ring])
                barrier()
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record()
                for i in range(args.steps):
                    gs(optimizer, *sring[i % n_ring])
                s1.record()
                barrier()
                tt = torch.tensor([s0.elapsed_time(s1)], device=dev)
                torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
                gms = float(tt) / args.steps
                strong['cuda_graph'] = {'ms_per_step': gms, 'value': N_RAYS / (gms / 1e3)}
                del gs
            except Exception as e:          # a capture problem must not cost the bench line
                strong['cuda_graph'] = {'error': f'{type(e).__name__}: {e}'[:200]}

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
    # ---- baselines measured on this box in the same run (rank 0, N = 1 only): the unmodified reference on the host cores
    # (BASELINE configs[0]: 1024 rays) and its own eager-PyTorch CUDA path on this GPU (configs[1] train step, configs[2] render)
    cpu = gpu_base = None
    if not args.no_cpu and world == 1:
        cpu = cpu_reference(args.cpu_rays, 3, 1)
        from baseline import ref_shim
        if ref_shim.available() and not llff:
            from baseline import ref_bench
            del ring
            torch.cuda.empty_cache()
            gb = ref_bench.time_gpu(dev, n_rays=N_RAYS, steps=8, warmup=2, render_frames=0 if args.no_render else 1)
            gpu_base = {'value': gb['train_rays_per_s'], 'unit': 'rays/s', 'ms_per_step': gb['train_ms_per_step'], 'kind': 'reference',
                        'what': 'the unmodified reference (baseline/_ref) on this GPU: PyTorch eager fp32, its own train step at 4096 rays '
                                '(make_o_d of the full image + host pixel selection + render + backward + Adam), CUDA events, 8 steps after 2',
                        'peak_mem_GB': gb['peak_mem_GB'], 'render_frames_per_s': gb.get('render_frames_per_s'),
                        'speedup_value': value / gb['train_rays_per_s'], 'speedup_e2e': e2e['value'] / gb['train_rays_per_s'],
                        'speedup_render': (render['frames_per_s'] / gb['render_frames_per_s']) if (render and gb.get('render_frames_per_s')) else None}
    line = {'metric': 'train_rays_per_s', 'value': value, 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD_LLFF if llff else WORKLOAD, 'rays_per_gpu': N_RAYS, 'global_rays': N_RAYS * world, 'samples': [S_C, S_F],
                       'precision': args.precision, 'parallelism': f'ray-sharded data parallel x{world}, ' + (
                           'gradients exchanged by copy-engine pushes into peer symmetric memory, sum folded into Adam (PeerGradExchange)'
                           if (dctx is not None and getattr(dctx, '_peer_exchange', None) is not None and graphed is None)
                           else 'one NCCL all-reduce of 1,191,690 fp32 (gradients of both nets + the two losses)') if world > 1 else 'single GPU',
                       'l2': 'per-step working set (activation stash >= 5 GB) exceeds the 126 MB L2; ring of 8 distinct ray batches',
                       'timed_region': f'{args.steps} steps = {ms_step * args.steps:.0f} ms; a sustained figure over 12,000 steps of the same step is in profiles/ (train_demo)',
                       'loss': float(loss.sum())},
            'roofline': roofline, 'cpu_baseline': cpu, 'gpu_baseline': gpu_base, 'e2e': e2e, 'render': render, 'gpu_launches': int(launches),
            'clocks': clocks, 'gpu_launches_per_step': launches / args.steps, 'dp_check': dp_check, 'strong_scaling': strong,
            'step_mode': ('cuda_graph: render+loss+backward' + ('+all-reduce' if dctx else '') + ' captured once (the 18 kernels + 2 memsets + 3 copies of a step replay as one '
                          'graph launch); 2 x (weight fold + re-pack) + 2 Adam launches follow from the host') if graphed is not None else 'eager launches',
            'eager_ms_per_step': eager_ms, 'cuda_graph': graph_info, 'host_launches_per_step': (6 if graphed is not None else launches / args.steps)}
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
    ap.add_argument('--cpu-rays', dest='cpu_rays', type=int, default=1024, help='rays per CPU-baseline step (BASELINE configs[0]: 1024)')
    ap.add_argument('--no-cpu', dest='no_cpu', action='store_true')
    ap.add_argument('--no-render', dest='no_render', action='store_true')
    ap.add_argument('--graph', dest='graph', action='store_true', help='time the captured CUDA graph of the step (trainer.GraphedTrainStep) as `value`; '
                    'by default the eagerly enqueued step is timed and the graph replay is reported beside it at N=1')
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

"""Soak test of the data-parallel step (torchrun, >= 2 GPUs): N steps of trainer.train_step in the current NB_DP_MODE; at the end
every rank's parameters must be bit-identical (replicas never diverge) and finite.  usage:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_soak.py [steps]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerf_pytorch_paeng_b200 import distributed, trainer  # noqa: E402
from nerf_pytorch_paeng_b200.engine import get_engine  # noqa: E402
from nerf_pytorch_paeng_b200.model import NeRF  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ctx = distributed.init_from_env('nccl')
local = int(os.environ.get('LOCAL_RANK', '0'))
dev = torch.device('cuda', local)
eng = get_engine(dev)
torch.manual_seed(0)
model = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev).set_precision('bf16')
opt = trainer.FlatAdam(model, lr=5e-4)
opts = bench.make_opts(rank_dev=local, seed=1000 + ctx.rank)
K = np.array([[bench.FOCAL, 0, 400.], [0, bench.FOCAL, 400.], [0, 0, 1.]])
poses = torch.from_numpy(bench.synthetic_poses(8)).to(dev)
gen = torch.Generator(device='cpu').manual_seed(7 + ctx.rank)
ring = []
for i in range(8):
    pix = torch.randperm(640000, generator=gen)[:4096].to(dev)
    o, d = eng.raygen(800, 800, K, poses[i, :3, :4], pix_idx=pix)
    ring.append((torch.cat((o, d), -1), torch.rand(4096, 3, generator=gen).to(dev)))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(20):
    trainer.train_step(model, opt, *ring[i % 8], opts, dist_ctx=ctx)
torch.cuda.synchronize()
ev0.record()
for i in range(steps):
    loss = trainer.train_step(model, opt, *ring[i % 8], opts, dist_ctx=ctx)
ev1.record()
torch.cuda.synchronize()
flat = torch.cat([model.model_coarse.flat, model.model_fine.flat])
digest = torch.stack([flat.double().sum(), flat.double().abs().sum(), flat.view(torch.int32).sum().double()])
all_d = [torch.empty_like(digest) for _ in range(ctx.world_size)]
dist.all_gather(all_d, digest)
same = all(torch.equal(all_d[0], x) for x in all_d)
if ctx.rank == 0:
    print({'mode': os.environ.get('NB_DP_MODE', 'auto'), 'peer': getattr(ctx, '_peer_exchange', None) is not None, 'world': ctx.world_size, 'steps': steps,
           'ms_per_step': ev0.elapsed_time(ev1) / steps, 'loss': [float(x) for x in loss], 'finite': bool(torch.isfinite(flat).all()),
           'replicas_bit_identical': same}, flush=True)
assert same and bool(torch.isfinite(flat).all())
dist.destroy_process_group()

"""Kernel timeline of one train step (torch.profiler / CUPTI): per-kernel start, duration and the idle gap before it.
Usage: python scripts/step_timeline.py  -> gpurun_out/step_timeline.json + a summary on stdout."""
import json
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nerf_pytorch_paeng_b200 import trainer  # noqa: E402
from nerf_pytorch_paeng_b200.engine import get_engine  # noqa: E402
from nerf_pytorch_paeng_b200.model import NeRF  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    eng = get_engine(dev)
    torch.manual_seed(0)
    model = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev).set_precision('bf16')
    opts = bench.make_opts(seed=1, device_select=True)
    opt = trainer.FlatAdam(model, lr=5e-4)
    K = np.array([[bench.FOCAL, 0, 400.], [0, bench.FOCAL, 400.], [0, 0, 1.]])
    poses = torch.from_numpy(bench.synthetic_poses(4)).to(dev)
    ring = []
    for i in range(4):
        pix = torch.randperm(640000)[:4096].to(dev)
        o, d = eng.raygen(800, 800, K, poses[i, :3, :4], pix_idx=pix)
        ring.append((torch.cat((o, d), -1), torch.rand(4096, 3, device=dev)))
    for i in range(10):
        trainer.train_step(model, opt, *ring[i % 4], opts)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(6):
            trainer.train_step(model, opt, *ring[i % 4], opts)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    rows = [{'name': e.name[:70], 'start_us': e.time_range.start, 'dur_us': e.time_range.end - e.time_range.start} for e in evs]
    # one step = from one stratified kernel to the next
    idx = [i for i, r in enumerate(rows) if 'stratified' in r['name']]
    a, b = idx[2], idx[3]
    step = rows[a:b]
    t0 = step[0]['start_us']
    prev_end = t0
    busy = 0.
    out = []
    for r in step:
        gap = r['start_us'] - prev_end
        out.append({'name': r['name'], 't_us': round(r['start_us'] - t0, 1), 'dur_us': round(r['dur_us'], 1), 'gap_before_us': round(gap, 1)})
        busy += r['dur_us']
        prev_end = max(prev_end, r['start_us'] + r['dur_us'])
    total = rows[b]['start_us'] - t0
    summary = {'step_us': total, 'busy_us': busy, 'idle_us': total - busy, 'n_events': len(step)}
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump({'summary': summary, 'events': out}, open(os.path.join(ROOT, 'gpurun_out', 'step_timeline.json'), 'w'), indent=1)
    print(json.dumps(summary))
    for r in out:
        print(f"{r['t_us']:9.1f} {r['dur_us']:9.1f} gap {r['gap_before_us']:7.1f}  {r['name']}")


if __name__ == '__main__':
    main()

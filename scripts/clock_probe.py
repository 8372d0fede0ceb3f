"""Sustained bf16 forward (inference) loop with nvidia-smi clock / power sampling: is the chain kernel power-capped?"""
import os, subprocess, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_pytorch_paeng_b200.engine import get_engine
from nerf_pytorch_paeng_b200.model import NeRF
dev = torch.device('cuda', 0)
eng = get_engine(dev)
torch.manual_seed(0)
net = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev).set_precision('bf16')
m = net.model_fine
flat = m.flat_params(); pk = m.packed_weights()
n, S = 16384, 192
rays = torch.cat([torch.zeros(n, 3, device=dev), torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)], -1)
z = torch.sort(torch.rand(n, S, device=dev) * 4 + 2, -1)[0]
rows = []
proc = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,temperature.gpu',
                         '--format=csv,noheader,nounits', '-lms', '100', '-i', '0'], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l.strip()) for l in proc.stdout], daemon=True).start()
for _ in range(5):
    eng.mlp_forward(m.desc, flat, pk, m.precision, rays=rays, z=z)
torch.cuda.synchronize()
for dur in (0.2, 2.0):
    t0 = time.time(); it = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < dur:
        for _ in range(20):
            eng.mlp_forward(m.desc, flat, pk, m.precision, rays=rays, z=z)
        it += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    print(f'dur={dur}s ms/launch={ms:.3f} TFLOP/s={1186816 * n * S / ms / 1e9:.1f}')
time.sleep(0.2)
proc.terminate()
print('samples (sm_mhz, W, sw_power_cap, hw_slowdown, temp):')
for r in rows[::3][:40]:
    print('  ', r)

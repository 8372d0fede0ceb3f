import sys, torch
sys.path.insert(0, '/root/repo')
from nerf_pytorch_paeng_b200.engine import get_engine
dev = torch.device('cuda', 0); eng = get_engine(dev)
for (n, sc, sf) in ((4096, 64, 128), (65536, 64, 128), (4096, 128, 256), (4096, 256, 512)):
    z = torch.sort(torch.rand(n, sc, device=dev) * 4 + 2, -1)[0]; w = torch.rand(n, sc, device=dev)
    for _ in range(3): eng.sample_pdf(z, w, sf, seed=1, offset=0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): out = eng.sample_pdf(z, w, sf, seed=1, offset=0)
    e1.record(); torch.cuda.synchronize()
    zf = out[0]
    ok = bool((zf[:, 1:] >= zf[:, :-1]).all())
    print(n, sc, sf, f'{e0.elapsed_time(e1)/20*1e3:.1f} us', 'sorted' if ok else 'NOT SORTED')

"""Run the UNMODIFIED reference (baseline/_ref) on the B200 itself and compare the CUDA kernels of this repo with it ON THE
SAME DEVICE (VERDICT r1 item 2 / SURVEY section 7 step 0):

  (i)  CUDA-side golden comparisons with mismatch COUNTS: make_o_d (800x800 and 1008x756 + ndc_rays), sample_pdf with
       perturb=0 (cdf, bin indices, samples), post_process, and the whole fp32 render_rays with injected draws;
  (ii) dumps the reference's CUDA intermediates (row sums / cumsum of the pdf) so that the summation order of torch.sum /
       torch.cumsum on this device can be restated offline (gpurun_out/ref_cuda_probe.npz);
  (iii) times the eager reference (train step at 4096 rays, 800x800 render) -> gpurun_out/ref_probe.json.

    python scripts/ref_probe.py [--no-time]
"""
import argparse
import json
import os
import sys
from types import SimpleNamespace
from unittest import mock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_bench, ref_shim  # noqa: E402
from nerf_pytorch_paeng_b200 import nerf_process as NP  # noqa: E402
from nerf_pytorch_paeng_b200.engine import get_engine  # noqa: E402
from nerf_pytorch_paeng_b200.model import NeRF, get_positional_encoder  # noqa: E402

OUT = os.path.join(ROOT, 'gpurun_out')


def bits_differ(a, b):
    a = a.detach().contiguous().view(torch.int32) if a.dtype == torch.float32 else a
    b = b.detach().contiguous().view(torch.int32) if b.dtype == torch.float32 else b
    return int((a != b).sum())


class Recorder:
    def __init__(self):
        self.rand, self.inds = [], []
        self._rand, self._ss = torch.rand, torch.searchsorted

    def rand_fn(self, *a, **k):
        r = self._rand(*a, **k)
        self.rand.append(r.clone())
        return r

    def ss_fn(self, *a, **k):
        r = self._ss(*a, **k)
        self.inds.append(r.clone())
        return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--no-time', action='store_true')
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    dev = torch.device('cuda', 0)
    eng = get_engine(dev)
    ref = ref_shim.import_reference()
    rec = Recorder()
    res = {'device': torch.cuda.get_device_name(0), 'torch': torch.__version__}
    dump = {}

    # ---------------------------------------------------------------- K1 make_o_d / ndc on CUDA
    poses = ref.get_render_pose(n_angle=120, single_angle=-1, phi=-30.0, nf=4.0)
    K = np.array([[ref_bench.FOCAL, 0, 400.], [0, ref_bench.FOCAL, 400.], [0, 0, 1]])
    tot = mism = 0
    for pi in (0, 33, 77):
        pose = poses[pi, :3, :4].to(dev)
        o_r, d_r = ref.rays.make_o_d(800, 800, torch.from_numpy(K).to(dev), pose)
        o_m, d_m = eng.raygen(800, 800, K, pose)
        mism += bits_differ(d_r.reshape(-1, 3), d_m) + bits_differ(o_r.reshape(-1, 3).contiguous(), o_m)
        tot += 2 * d_m.numel()
    res['make_o_d_800'] = {'values': tot, 'bit_mismatches': mism}
    Hl, Wl, fl = 756, 1008, 815.13158
    Kl = np.array([[fl, 0, .5 * Wl], [0, fl, .5 * Hl], [0, 0, 1]])
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ndc.npz'))
    pose_l = torch.from_numpy(g['pose']).to(dev)
    o_r, d_r = ref.rays.make_o_d(Wl, Hl, torch.from_numpy(Kl).to(dev), pose_l)
    o_m, d_m = eng.raygen(Hl, Wl, Kl, pose_l)
    res['make_o_d_llff'] = {'values': 2 * d_m.numel(), 'bit_mismatches': bits_differ(d_r.reshape(-1, 3), d_m) + bits_differ(o_r.reshape(-1, 3).contiguous(), o_m)}
    on_r, dn_r = ref.proc.ndc_rays(Hl, Wl, Kl[0][0], 1., o_r.reshape(-1, 3).contiguous(), d_r.reshape(-1, 3).contiguous())
    on_m, dn_m = eng.ndc_rays(Hl, Wl, fl, 1., o_r.reshape(-1, 3).contiguous(), d_r.reshape(-1, 3).contiguous())
    res['ndc_rays_llff'] = {'values': 2 * dn_m.numel(), 'bit_mismatches': bits_differ(on_r, on_m) + bits_differ(dn_r, dn_m),
                            'max_abs': float(max((on_r - on_m).abs().max(), (dn_r - dn_m).abs().max()))}

    # ---------------------------------------------------------------- K2 sample_pdf with perturb=0 on CUDA
    # three row-count regimes of ATen's cumsum kernel (ScanUtils.cuh: 32 / 16 / 512 threads per row for rows of 62)
    opts_det = SimpleNamespace(near=2., far=6., gpu_ids=[0], rank=0, N_samples_c=64, N_samples_f=128, perturb=0., chunk_pts=524288,
                               chunk_rays=4096, data_type='blender')
    sp = np.load(os.path.join(ROOT, 'tests', 'golden', 'sample_pdf.npz'))
    u_det = torch.linspace(0., 1., steps=128, device=dev)
    for n_rows in (24, 8512, 40000):
        gen = torch.Generator(device='cpu').manual_seed(11 + n_rows)
        z_all = torch.sort(torch.rand(n_rows, 64, generator=gen) * 4 + 2, -1)[0]
        w_all = torch.rand(n_rows, 64, generator=gen) ** 8
        w_all[:n_rows // 64] = 0.
        if n_rows == 8512:              # the CPU fixture's rays (peaky weights, empty rays) ride along
            z_all[:320], w_all[:320] = torch.from_numpy(sp['z_vals']), torch.from_numpy(sp['weights'])
        z_all, w_all = z_all.to(dev), w_all.to(dev)
        mids = .5 * (z_all[..., 1:] + z_all[..., :-1])
        with mock.patch('torch.searchsorted', rec.ss_fn):
            s_ref = ref.proc.sample_pdf(mids, w_all[..., 1:-1], 128, det=True, opts=opts_det)
        inds_ref = rec.inds[-1]
        # the reference's own intermediates, op for op (nerf_process.py:150-154), on this device
        ww = w_all[..., 1:-1] + 1e-5
        wsum = torch.sum(ww, -1, keepdim=True)
        pdf = ww / wsum
        csum = torch.cumsum(pdf, -1)
        cdf_ref = torch.cat([torch.zeros_like(csum[..., :1]), csum], -1)
        keep = min(n_rows, 768)
        dump.update({f'n{n_rows}_z': z_all[:keep].cpu().numpy(), f'n{n_rows}_w': w_all[:keep].cpu().numpy(),
                     f'n{n_rows}_sum': wsum[:keep].cpu().numpy(), f'n{n_rows}_cdf': cdf_ref[:keep].cpu().numpy(),
                     f'n{n_rows}_inds': inds_ref[:keep].cpu().numpy().astype(np.int16), f'n{n_rows}_samples': s_ref[:keep].cpu().numpy()})
        for mode_name, rows in (('cuda_order', 0), ('fp64_order', -1)):
            _, zs, inds, cdf = eng.sample_pdf(z_all, w_all, 128, u=u_det, want_samples=True, want_inds=True, want_cdf=True, cdf_rows=rows)
            res[f'sample_pdf_det_n{n_rows}_{mode_name}'] = {
                'rays': int(n_rows), 'inds': int(inds.numel()), 'inds_mismatches': int((inds != inds_ref).sum()),
                'cdf_bit_mismatches': bits_differ(cdf, cdf_ref), 'cdf_max_abs': float((cdf - cdf_ref).abs().max()),
                'samples_bit_mismatches': bits_differ(zs, s_ref), 'samples_max_abs': float((zs - s_ref).abs().max())}
        # the sub-batch call the fixture test makes: first `keep` rows, regime pinned by cdf_rows
        _, zs, inds, cdf = eng.sample_pdf(z_all[:keep].contiguous(), w_all[:keep].contiguous(), 128, u=u_det, want_samples=True,
                                          want_inds=True, want_cdf=True, cdf_rows=n_rows)
        res[f'sample_pdf_det_n{n_rows}_subbatch'] = {'inds_mismatches': int((inds != inds_ref[:keep]).sum()),
                                                     'cdf_bit_mismatches': bits_differ(cdf, cdf_ref[:keep].contiguous())}
    # with the reference's cdf injected: search + interpolation alone
    _, zs, inds, _ = eng.sample_pdf(z_all, w_all, 128, u=u_det, cdf_in=cdf_ref.contiguous(), want_samples=True, want_inds=True)
    res['sample_pdf_det_ref_cdf'] = {'inds_mismatches': int((inds != inds_ref).sum()), 'samples_bit_mismatches': bits_differ(zs, s_ref)}

    # ---------------------------------------------------------------- K5 post_process on CUDA
    for S in (64, 192):
        gp = np.load(os.path.join(ROOT, 'tests', 'golden', f'post_process_S{S}.npz'))
        raw, zz, dd = (torch.from_numpy(gp[k]).to(dev) for k in ('raw', 'z_vals', 'rays_d'))
        raw_r = raw.clone().requires_grad_(True)
        outs = ref.proc.post_process(raw_r, zz, dd)
        gup = torch.from_numpy(gp['d_rgb']).to(dev)
        (outs[0] * gup).sum().backward()
        rgb, disp, acc, wts, depth = eng.composite_forward(raw, zz, dd)
        d_raw = eng.composite_backward(raw, zz, dd, gup)
        res[f'post_process_S{S}'] = {
            'rgb_max_abs': float((rgb - outs[0]).abs().max()), 'disp_max_abs': float((disp - outs[1]).abs().max()),
            'weights_max_abs': float((wts - outs[3]).abs().max()), 'depth_max_abs': float((depth - outs[4]).abs().max()),
            'd_raw_max_abs': float((d_raw - raw_r.grad).abs().max()),
            'vs_cpu_golden_rgb_max_abs': float((outs[0].detach().cpu() - torch.from_numpy(gp['rgb_map'])).abs().max())}

    # ---------------------------------------------------------------- whole render (W=256, random init) fp32 + bf16 vs reference on CUDA
    torch.manual_seed(0)
    net_r = ref.NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev)
    torch.manual_seed(0)
    net_m = NeRF(8, 256, 63, 27, [4], gt_camera_param=(None, None)).to(dev)
    net_m.load_state_dict(net_r.state_dict())
    fx, fd = ref.posenc(10)[0], ref.posenc(4)[0]
    opts = SimpleNamespace(near=2., far=6., gpu_ids=[0], rank=0, N_samples_c=64, N_samples_f=128, perturb=1., chunk_pts=524288,
                           chunk_rays=4096, data_type='blender', seed=0)
    pose = poses[33, :3, :4].to(dev)
    o_r, d_r = ref.rays.make_o_d(800, 800, torch.from_numpy(K).to(dev), pose)
    sel = torch.from_numpy(np.random.RandomState(5).choice(640000, 2048, replace=False)).to(dev)
    ro, rd = o_r.reshape(-1, 3)[sel].contiguous(), d_r.reshape(-1, 3)[sel].contiguous()
    target = torch.rand(2048, 3, device=dev)
    for scale in (1.0, 30.0):
        if scale != 1.0:
            with torch.no_grad():
                for m in (net_r.model_coarse, net_r.model_fine):
                    m.linear_density.weight.mul_(scale)
            net_m.load_state_dict(net_r.state_dict())
        rec.rand.clear()
        net_r.zero_grad()
        with mock.patch('torch.rand', rec.rand_fn):
            rgb_c, disp_c, rgb_f, disp_f = ref.proc.batchify_rays_and_render_by_chunk(ro, rd, net_r, [fx, fd], 800, 800, K, opts)
        t_rand, u = rec.rand[0], rec.rand[1]
        crit = torch.nn.MSELoss()
        (crit(rgb_c, target) + crit(rgb_f, target)).backward()
        g_ref = {n: p.grad.detach().clone() for n, p in net_r.named_parameters()}
        for prec in ('fp32', 'bf16'):
            net_m.set_precision(prec)
            o2 = SimpleNamespace(**vars(opts), rng={'t_rand': t_rand, 'u': u})
            with torch.no_grad():
                mc, mdc, mf, mdf = NP.batchify_rays_and_render_by_chunk(ro, rd, net_m, None, 800, 800, K, o2)

            def psnr(a, b):
                return float(-10 * torch.log10(((a - b) ** 2).mean()))
            ec, ef = (mc - rgb_c).abs().max(dim=-1)[0], (mf - rgb_f).abs().max(dim=-1)[0]
            from nerf_pytorch_paeng_b200 import trainer
            for net in (net_m.model_coarse, net_m.model_fine):
                net.bind_flat_grad().zero_()
            trainer.render_losses_and_grads(net_m, torch.cat((ro, rd), -1), target, o2)
            gerr = {}
            for tag, mod in (('coarse', net_m.model_coarse), ('fine', net_m.model_fine)):
                gm = mod.flat_grad
                gr = torch.cat([g_ref[f'model_{tag}.' + n].reshape(-1) for n, _ in mod.named_parameters()])
                gerr[tag] = float((gm - gr).norm() / gr.norm())
            res[f'render_w256_scale{int(scale)}_{prec}'] = {
                'rays': 2048, 'psnr_c_dB': psnr(mc, rgb_c.detach()), 'psnr_f_dB': psnr(mf, rgb_f.detach()),
                'rgb_c_max_abs': float(ec.max()), 'rgb_f_max_abs': float(ef.max()),
                'rays_over_1e-4_c': int((ec > 1e-4).sum()), 'rays_over_1e-4_f': int((ef > 1e-4).sum()),
                'rays_over_1e-2_c': int((ec > 1e-2).sum()), 'rays_over_1e-2_f': int((ef > 1e-2).sum()),
                'grad_rel_err': gerr}
    np.savez_compressed(os.path.join(OUT, 'ref_cuda_probe.npz'), **dump)

    # ---------------------------------------------------------------- timings
    if not args.no_time:
        res['gpu_baseline'] = ref_bench.time_gpu(dev, n_rays=4096, steps=10, warmup=3, render_frames=1)
        res['cpu_baseline'] = ref_bench.time_cpu(n_rays=1024, steps=3, warmup=1)
    print(json.dumps(res, indent=1))
    json.dump(res, open(os.path.join(OUT, 'ref_probe.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()

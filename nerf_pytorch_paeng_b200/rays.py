"""Drop-in for the reference's rays.py (file:line refs are into the reference)."""
import numpy as np
import torch

from .engine import get_engine


def _as_cuda_pose(pose, device=None):
    if isinstance(pose, np.ndarray):
        pose = torch.from_numpy(np.ascontiguousarray(pose, dtype=np.float32))
    if not pose.is_cuda:
        pose = pose.to(device if device is not None else torch.device('cuda', torch.cuda.current_device()))
    return pose.float()


def get_rays_np(H, W, K, c2w):
    """rays.py:7-17.  Host NumPy arrays for the global-batch precompute (main.py:95-101), bit-exact with the reference under
    NumPy >= 2: K's float64 entries promote the computation to float64, so rays_d is float64 [H,W,3] (computed by
    the fp64 ray-gen kernel with the same individually rounded operations) and rays_o is the float32 c2w[:3,-1] broadcast to
    that shape (a read-only view, like np.broadcast_to in the reference).  main.py:101 casts the stack to float32."""
    c2w_np = np.asarray(c2w.detach().cpu() if isinstance(c2w, torch.Tensor) else c2w)
    pose = _as_cuda_pose(c2w)
    d = get_engine(pose.device).raygen_f64(H, W, _host_K(K), pose)
    rays_d = d.reshape(H, W, 3).cpu().numpy()
    return np.broadcast_to(c2w_np[:3, -1], rays_d.shape), rays_d


def make_o_d(img_w, img_h, img_k, pose):
    """rays.py:20-34.  pose: CUDA tensor [3|4, 4]; img_k: 3x3 tensor or ndarray (float64 in the reference).
    Returns rays_o (a stride-0 expand of pose[:3,-1], exactly like rays.py:33) and rays_d [H,W,3]."""
    pose = _as_cuda_pose(pose)
    _, d = get_engine(pose.device).raygen(img_h, img_w, _host_K(img_k), pose)
    rays_d = d.view(img_h, img_w, 3)
    rays_o = pose[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def make_o_d_selected(img_w, img_h, img_k, pose, pix_idx, ndc=False, near=1.):
    """Fast path: generate only the selected pixels' rays (flat indices r*W+c),
    optionally already NDC-warped.  Equivalent to make_o_d(...)[...][selected] (+ ndc_rays)."""
    pose = _as_cuda_pose(pose)
    K = _host_K(img_k)
    return get_engine(pose.device).raygen(img_h, img_w, K, pose, pix_idx=pix_idx, ndc=ndc, ndc_focal=K[0][0], ndc_near=near)


_K_cache = {}


def _host_K(img_k):
    """K as a host 3x3 of Python floats.  A CUDA tensor K (train.py:18-19) costs one D2H copy the
    first time it is seen; results are cached by (data_ptr, version)."""
    if isinstance(img_k, torch.Tensor):
        if img_k.is_cuda:
            key = (img_k.data_ptr(), img_k._version)
            if key not in _K_cache:
                if len(_K_cache) > 64:
                    _K_cache.clear()
                _K_cache[key] = img_k.detach().double().cpu().numpy()
            return _K_cache[key]
        return img_k.detach().double().numpy()
    return np.asarray(img_k, dtype=np.float64)


def sample_rays_and_pixel(i, rays_o, rays_d, target_img, opts):
    """rays.py:37-64.  Same host RNG call (np.random.choice, replace=False) so a seeded run picks the
    same pixels as the reference; the three gathers run on the GPU.  precrop as rays.py:40-45."""
    img_h, img_w = target_img.shape[:2]
    if i < opts.precrop_iters:
        dH = int(img_h // 2 * opts.precrop_frac)
        dW = int(img_w // 2 * opts.precrop_frac)
        rows = np.arange(img_h // 2 - dH, img_h // 2 + dH)
        cols = np.arange(img_w // 2 - dW, img_w // 2 + dW)
    else:
        rows = np.arange(img_h)
        cols = np.arange(img_w)
    n_coords = rows.size * cols.size
    selected_idx = np.random.choice(a=n_coords, size=opts.N_rays, replace=False)
    r = rows[selected_idx // cols.size]
    c = cols[selected_idx % cols.size]
    flat = torch.from_numpy((r * img_w + c).astype(np.int64)).to(rays_d.device)
    eng = get_engine(rays_d.device)
    d_sel = eng.gather_rows(rays_d.reshape(-1, 3), flat)
    if rays_o.stride(0) == 0 and rays_o.stride(1) == 0:          # the expand view make_o_d returns
        o_sel = rays_o[0, 0].expand(opts.N_rays, 3)
    else:
        o_sel = eng.gather_rows(rays_o.reshape(-1, 3).contiguous(), flat)
    t_sel = eng.gather_rows(target_img.reshape(-1, 3), flat)
    return o_sel, d_sel, t_sel

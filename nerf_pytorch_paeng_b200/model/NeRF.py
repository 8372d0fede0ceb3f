"""Drop-in for the reference's model/NeRF.py (file:line refs are into the reference).

Same constructor, sub-module names and state_dict keys as NeRF.py:10-30,55-65
(``model_{coarse,fine}.linear_x.{0..D-1}``, ``linear_d``, ``linear_feat``, ``linear_density``,
``linear_color``), same Xavier-uniform init consumed in the same RNG order, so checkpoints are
interchangeable (optimizer state: trainer.FlatAdam speaks torch.optim.Adam's state_dict format).  ``forward`` runs the CUDA MLP (K4) instead of nn.Linear/cuBLAS:

* every parameter of one NeRFModule is a view into ONE flat fp32 buffer laid out in
  ``parameters()`` order (the layout of nb_mlp_desc in include/nerf_b200.h); gradients likewise.
  The flat buffers are what the kernels, the fused Adam and the NCCL all-reduce operate on.
* ``precision``: NB_FP32 (CUDA-core parity path) or NB_BF16 (tcgen05 path, weights re-packed to
  the tile layout whenever the flat buffer's version changes).
"""
import torch
import torch.nn as nn

from ..engine import NB_BF16, NB_FP32, MlpDesc, NBError, get_engine


class _MlpFunction(torch.autograd.Function):
    """raw = MLP(x) with gradients for the parameters only (sample positions are data,
    nerf_process.py:66).  ``src`` is ('emb', x) or ('rays', rays, z)."""

    @staticmethod
    def forward(ctx, module, src, *params):
        eng = get_engine(module.flat.device)
        need_grad = any(ctx.needs_input_grad)   # (grad mode is off inside Function.forward)
        kw = dict(x=src[1]) if src[0] == 'emb' else dict(rays=src[1], z=src[2])
        raw, act = eng.mlp_forward(module.desc, module.flat, module.packed_weights(), module.precision, save=need_grad, **kw)
        ctx.module = module
        ctx.act = act
        ctx.n_pts = raw.shape[0]
        return raw

    @staticmethod
    def backward(ctx, d_raw):
        m = ctx.module
        eng = get_engine(m.flat.device)
        grad = torch.empty_like(m.flat)
        eng.mlp_backward(m.desc, m.flat, m.packed_weights(), m.precision, ctx.n_pts, ctx.act, d_raw.contiguous(), grad)
        ctx.act = None
        outs = [grad[o:o + n].view(s) for (o, n, s) in m.slices]
        return (None, None, *outs)


class NeRFModule(nn.Module):
    def __init__(self, D: int, W: int, input_ch: int, input_ch_d: int, skips=[4]):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch_x, self.input_ch_d = input_ch, input_ch_d
        self.skips = skips
        if len(skips) > 1:
            raise NBError('the CUDA MLP supports at most one skip layer (reference configs use skips=[4])')
        # identical construction order to NeRF.py:24-30 (=> identical default bias init draws)
        self.linear_x = nn.ModuleList(
            [nn.Linear(input_ch, W)] + [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W) for i in range(D - 1)])
        self.linear_d = nn.Linear(input_ch_d + W, W // 2)
        self.linear_feat = nn.Linear(W, W)
        self.linear_density = nn.Linear(W, 1)
        self.linear_color = nn.Linear(W // 2, 3)
        self.precision = NB_FP32
        self.desc = MlpDesc(D, W, input_ch, input_ch_d, skips[0] if skips else -1,
                            (input_ch - 3) // 6, (input_ch_d - 3) // 6)
        self.flat = None
        self.flat_grad = None
        self.slices = None
        self._packed = None
        self._packed_version = None
        self._plist = None          # cached tuple(self.parameters()); nn.Module traversal is too slow for the per-step path
        self._flat_dirty = True

    def _apply(self, fn, *a, **k):                       # .to() / .cuda() / .float(): storages may move
        self._flat_dirty = True
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):            # load_state_dict(assign=True) may replace the Parameters
        self._flat_dirty = True
        return super()._load_from_state_dict(*a, **k)

    # ---- flat parameter storage -------------------------------------------------------------
    def _flatten(self):
        """(Re)build the flat buffer and make every Parameter a view of it.  Called lazily: .to(),
        .cuda() or load_state_dict(assign=True) may have replaced the storages."""
        if not self._flat_dirty and self.flat is not None:
            # fast path: spot-check the two ends of the buffer; the full check below runs after any _apply / load
            p0, p1 = self._plist[0], self._plist[-1]
            if p0.data_ptr() == self.flat.data_ptr() and p1.data_ptr() + 4 * p1.numel() == self.flat.data_ptr() + 4 * self.flat.numel():
                return
        params = self._plist = tuple(self.parameters())
        self._flat_dirty = False
        # host storage is allowed (checkpoint / optimizer-state handling); every COMPUTE entry point goes through get_engine(),
        # which raises on a non-CUDA device: there is no CPU fallback
        dev = params[0].device
        total = sum(p.numel() for p in params)
        ok = self.flat is not None and self.flat.device == dev and self.flat.numel() == total
        if ok:
            off = 0
            base = self.flat.data_ptr()
            for p in params:
                if p.data_ptr() != base + 4 * off or not p.is_contiguous():
                    ok = False
                    break
                off += p.numel()
        if ok:
            return
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        slices, off = [], 0
        for p in params:
            n = p.numel()
            flat[off:off + n].copy_(p.data.reshape(-1).float())
            p.data = flat[off:off + n].view(p.shape)
            slices.append((off, n, tuple(p.shape)))
            off += n
        self.flat, self.slices = flat, slices
        self.flat_grad = None
        self._packed_version = None

    def flat_params(self):
        self._flatten()
        return self.flat

    def bind_flat_grad(self):
        """Point every p.grad at a slice of one flat gradient buffer (for fused backward / all-reduce)."""
        self._flatten()
        if self.flat_grad is None:
            self.flat_grad = torch.zeros_like(self.flat)
        g0, g1 = self._plist[0].grad, self._plist[-1].grad
        if (g0 is not None and g1 is not None and g0.data_ptr() == self.flat_grad.data_ptr()
                and g1.data_ptr() + 4 * g1.numel() == self.flat_grad.data_ptr() + 4 * self.flat_grad.numel()):
            return self.flat_grad                        # still bound (zero_grad(set_to_none=True) would have cleared both ends)
        for p, (o, n, s) in zip(self._plist, self.slices):
            g = self.flat_grad[o:o + n].view(s)
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g
        return self.flat_grad

    def packed_weights(self):
        if self.precision != NB_BF16:
            return None
        self._flatten()
        eng = get_engine(self.flat.device)
        # Parameters keep their own version counters (set_data), so track those: any in-place update
        # (optimizer.step, load_state_dict) bumps them
        ver = (self.flat.data_ptr(), self.flat._version) + tuple(p._version for p in self._plist)
        if self._packed is None or self._packed.device != self.flat.device:
            self._packed = torch.empty(eng.mlp_packed_bytes(self.desc), dtype=torch.uint8, device=self.flat.device)
            self._packed_version = None
        if self._packed_version != ver:
            eng.mlp_pack(self.desc, self.flat, self._packed)
            self._packed_version = ver
        return self._packed

    def mark_weights_changed(self):
        self._packed_version = None

    # ---- forward ----------------------------------------------------------------------------
    def forward(self, x):
        """NeRF.py:33-52 on a materialised embedding x[n, input_ch+input_ch_d] -> [n, 4] = [rgb, sigma]."""
        self._flatten()
        return _MlpFunction.apply(self, ('emb', x), *self._plist)

    def forward_rays(self, rays, z_vals):
        """Fused form used by render_rays: points + both encodings are generated in-kernel."""
        self._flatten()
        return _MlpFunction.apply(self, ('rays', rays, z_vals), *self._plist)


class NeRF(nn.Module):
    def __init__(self, D: int, W: int, input_ch: int, input_ch_d: int, skips=[4], gt_camera_param=None, device=None,
                 precision=NB_FP32):
        super().__init__()
        self.model_coarse = NeRFModule(D, W, input_ch, input_ch_d, skips)
        self.model_fine = NeRFModule(D, W, input_ch, input_ch_d, skips)
        self.apply(self._init_weights)                               # NeRF.py:60,63-65
        if gt_camera_param is None:
            gt_camera_param = (None, None)
        self.gt_intrinsic, self.gt_extrinsic = gt_camera_param
        self.set_precision(precision)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)

    def set_precision(self, precision):
        if precision in ('fp32', 'bf16'):
            precision = {'fp32': NB_FP32, 'bf16': NB_BF16}[precision]
        if precision not in (NB_FP32, NB_BF16):
            raise NBError(f'unknown precision {precision!r} (use "bf16" or "fp32")')
        self.model_coarse.precision = precision
        self.model_fine.precision = precision
        return self

    def get_camera_gt(self):
        return self.gt_intrinsic, self.gt_extrinsic

    def forward(self, x, is_fine: bool = False):
        """NeRF.py:70-78."""
        if is_fine:
            return self.model_fine(x)
        return self.model_coarse(x)

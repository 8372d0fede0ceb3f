"""Import the unmodified reference from baseline/_ref (see install_ref.py).  BASELINE / TEST INFRASTRUCTURE ONLY: the
product package never imports this.

Shims: utils.py:3 imports the absent third-party `IQA_pytorch` -> stub module with dummy SSIM / LPIPSvgg.
On a GPU nothing else is needed (the reference hard-codes cuda:N devices).  For CPU execution (`cpu=True`) two patches make
the unmodified source run on the host: torch.device -> cpu and Tensor.get_device -> 'cpu' (nerf_process.py:45-59,94,158-163,
rays.py:23-24).
"""
import contextlib
import importlib.util
import os
import sys
import types
from types import SimpleNamespace
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


def available():
    return os.path.exists(os.path.join(REF_DIR, 'nerf_process.py'))


def import_reference():
    """-> namespace(rays, proc, NeRF, posenc, scheduler, get_render_pose) of the reference's own modules."""
    if not available():
        raise RuntimeError(f'{REF_DIR} is missing: run `python baseline/install_ref.py` in the build container')
    if 'IQA_pytorch' not in sys.modules:
        stub = types.ModuleType('IQA_pytorch')
        stub.SSIM = object
        stub.LPIPSvgg = object
        sys.modules['IQA_pytorch'] = stub
    # the reference's top-level module names (rays, model, utils, ...) must resolve to ITS files: import them with REF_DIR first
    # on sys.path, then hide them from sys.modules again so that nothing else picks them up by accident
    names = ['rays', 'nerf_process', 'utils', 'scheduler', 'model', 'model.NeRF', 'model.NeRFHelper', 'model.PositionalEncoding']
    saved = {n: sys.modules.pop(n) for n in list(sys.modules) if n in names}
    sys.path.insert(0, REF_DIR)
    try:
        import rays as ref_rays
        import nerf_process as ref_np
        import scheduler as ref_sched
        from model import NeRF as RefNeRF, get_positional_encoder as ref_posenc
        spec = importlib.util.spec_from_file_location('ref_render_pose', os.path.join(REF_DIR, 'dataset', 'render_pose.py'))
        rp = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(rp)
    finally:
        sys.path.remove(REF_DIR)
        for n in names:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
    for m in (ref_rays, ref_np, ref_sched):
        assert os.path.dirname(os.path.abspath(m.__file__)) == REF_DIR, m.__file__
    return SimpleNamespace(rays=ref_rays, proc=ref_np, NeRF=RefNeRF, posenc=ref_posenc, scheduler=ref_sched,
                           get_render_pose=rp.get_render_pose)


@contextlib.contextmanager
def on_cpu():
    """Run the unmodified reference on the host: its hard-coded cuda devices resolve to cpu inside this context."""
    import torch
    import torch._dynamo  # noqa: F401  (optimizer construction imports it lazily; must happen before torch.device is patched)
    real_device = torch.device
    patches = [mock.patch('torch.device', lambda *a, **k: real_device('cpu')),
               mock.patch('torch.Tensor.get_device', lambda self: 'cpu')]
    for p in patches:
        p.start()
    try:
        yield
    finally:
        for p in patches:
            p.stop()

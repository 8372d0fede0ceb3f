"""CPU: the C-ABI library loads and exports every symbol include/nerf_b200.h declares; the ctypes
signature table matches the header's parameter counts; the product path refuses to run without a GPU."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, 'include', 'nerf_b200.h')


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    src = re.sub(r'^\s*#.*$', '', src, flags=re.M)
    out = {}
    for m in re.finditer(r'\b(?:int|int64_t|const char\*)\s+(nb_\w+)\s*\(([^;]*?)\)\s*;', src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ('', 'void') else len([a for a in args.split(',') if a.strip()])
        out[m.group(1)] = n
    return out


def test_library_exports_every_declared_symbol():
    from nerf_pytorch_paeng_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), 'run __graft_entry__.build() first'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    fns = header_functions()
    assert len(fns) >= 25
    for name in fns:
        assert hasattr(lib, name), f'{name} declared in nerf_b200.h but not exported'
    assert lib.nb_abi_version() == 2


def test_ctypes_table_matches_header():
    from nerf_pytorch_paeng_b200 import _lib
    fns = header_functions()
    assert set(fns) == set(_lib.SIGNATURES), set(fns) ^ set(_lib.SIGNATURES)
    for name, n in fns.items():
        assert len(_lib.SIGNATURES[name][1]) == n, name


def test_mlp_desc_struct_layout():
    from nerf_pytorch_paeng_b200._lib import MlpDesc
    assert ctypes.sizeof(MlpDesc) == 7 * 4
    assert [f[0] for f in MlpDesc._fields_] == ['D', 'W', 'in_x', 'in_d', 'skip', 'L_x', 'L_d']


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-only check')
def test_no_cpu_fallback():
    """Without a CUDA device the product path raises instead of silently computing elsewhere."""
    from nerf_pytorch_paeng_b200 import _lib
    from nerf_pytorch_paeng_b200.engine import NBError, get_engine
    from nerf_pytorch_paeng_b200.model import NeRF, get_positional_encoder
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.nb_create(ctypes.byref(h), 0, 0) != 0
    with pytest.raises(NBError):
        get_engine()
    net = NeRF(8, 64, 63, 27, [4], gt_camera_param=(None, None))
    with pytest.raises(NBError):
        net(torch.zeros(4, 90))
    fn, _ = get_positional_encoder(10)
    with pytest.raises(NBError):
        fn(torch.zeros(4, 3))


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under nerf_pytorch_paeng_b200/ may reference it."""
    pkg = os.path.join(ROOT, 'nerf_pytorch_paeng_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f)).read()
                assert 'nerf_oracle' not in txt and 'from oracle' not in txt and 'import oracle' not in txt, f

"""ctypes binding of libnerf_b200.so (include/nerf_b200.h).

The shared library is built in-tree by ``make -C nerf_pytorch_paeng_b200/csrc`` (or
``__graft_entry__.build()``); it is the only compute backend: if it is missing, or no sm_100 GPU is
present, the product path raises -- there is no CPU or eager-PyTorch fallback.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libnerf_b200.so')
CSRC = os.path.join(_HERE, 'csrc')

NB_FP32, NB_BF16 = 0, 1
NB_RAYGEN_NDC = 1


class NBError(RuntimeError):
    pass


class MlpDesc(C.Structure):
    _fields_ = [('D', C.c_int32), ('W', C.c_int32), ('in_x', C.c_int32), ('in_d', C.c_int32),
                ('skip', C.c_int32), ('L_x', C.c_int32), ('L_d', C.c_int32)]


class RenderCfg(C.Structure):
    _fields_ = [('S_c', C.c_int32), ('S_f', C.c_int32), ('precision', C.c_int32), ('u_mode', C.c_int32),
                ('seed', C.c_uint64), ('offset_c', C.c_uint64), ('offset_f', C.c_uint64), ('cdf_rows', C.c_int64), ('ctr', C.c_void_p), ('exact_last', C.c_int32), ('reserved', C.c_int32)]


_p = C.c_void_p
_i32, _i64, _u64, _f32, _f64, _u32, _sz = C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double, C.c_uint, C.c_size_t
_desc = C.POINTER(MlpDesc)
_cfg = C.POINTER(RenderCfg)

# name -> (restype, argtypes); mirrors include/nerf_b200.h one to one (tests/test_abi.py checks it)
SIGNATURES = {
    'nb_abi_version': (C.c_int, []),
    'nb_create': (C.c_int, [C.POINTER(_p), C.c_int, _u32]),
    'nb_destroy': (C.c_int, [_p]),
    'nb_last_error': (C.c_char_p, [_p]),
    'nb_device_info': (C.c_int, [_p, C.POINTER(_i32 * 4)]),
    'nb_launch_count': (_i64, [_p]),
    'nb_raygen_pinhole': (C.c_int, [_p, _i32, _i32, _f64, _f64, _f64, _f64, _p, _i64, _p, _i64, _p, _p, _u32, _f64, _f64, _p]),
    'nb_raygen_pinhole_f64': (C.c_int, [_p, _i32, _i32, _f64, _f64, _f64, _f64, _p, _i64, _p, _p]),
    'nb_ndc_rays': (C.c_int, [_p, _i64, _i32, _i32, _f64, _f64, _p, _p, _p, _p, _p]),
    'nb_gather_rows': (C.c_int, [_p, _i64, _i32, _p, _p, _p, _p]),
    'nb_select_pixels': (C.c_int, [_p, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _u64, _u64, _p, _p]),
    'nb_stratified': (C.c_int, [_p, _i64, _i32, _p, _p, _p, _u64, _u64, _p, _p, _p]),
    'nb_counter_add': (C.c_int, [_p, _p, _u64, _p]),
    'nb_sample_pdf': (C.c_int, [_p, _i64, _i32, _i32, _p, _p, _p, _i32, _u64, _u64, _p, _p, _p, _p, _p, _p, _i64, _p, _p]),
    'nb_posenc': (C.c_int, [_p, _i64, _i32, _p, _p, _p]),
    'nb_embed_points': (C.c_int, [_p, _i64, _i32, _i32, _i32, _p, _p, _p, _i64, _p]),
    'nb_mlp_act_bytes': (C.c_int, [_p, _desc, _i64, _i32, C.POINTER(_sz)]),
    'nb_mlp_workspace_bytes': (C.c_int, [_p, _desc, _i64, _i32, _i32, C.POINTER(_sz)]),
    'nb_mlp_packed_bytes': (C.c_int, [_p, _desc, C.POINTER(_sz)]),
    'nb_mlp_pack': (C.c_int, [_p, _desc, _p, _p, _p]),
    'nb_mlp_forward_emb': (C.c_int, [_p, _desc, _p, _p, _i64, _p, _i64, _p, _p, _i32, _p, _sz, _p]),
    'nb_mlp_forward_rays': (C.c_int, [_p, _desc, _p, _p, _i64, _i32, _p, _p, _p, _p, _i32, _p, _sz, _p]),
    'nb_mlp_backward': (C.c_int, [_p, _desc, _p, _p, _i64, _p, _p, _p, _i32, _i32, _p, _sz, _p]),
    'nb_mlp_backward_stage': (C.c_int, [_p, _desc, _p, _p, _i64, _p, _p, _p, _i32, _i32, _p, _sz, _i32, _p]),
    'nb_mlp_tc_probe': (C.c_int, [_p, _desc, _p, _p, _i64, _i32, _p, _p, _i32, _p, _p, _p]),
    'nb_composite_forward': (C.c_int, [_p, _i64, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    'nb_composite_backward': (C.c_int, [_p, _i64, _i32, _p, _p, _p, _p, _p, _p]),
    'nb_render_workspace_bytes': (C.c_int, [_p, _desc, _i64, _cfg, _i32, C.POINTER(_sz)]),
    'nb_render_rays': (C.c_int, [_p, _desc, _cfg, _p, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    'nb_train_rays': (C.c_int, [_p, _desc, _cfg, _p, _p, _p, _p, _i64, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _i32, _p,
                                _p, _p, _p, _p, _i32, _p, _sz, _p]),
    'nb_frame_to8b': (C.c_int, [_p, _i64, _p, _p, _p, _p, _p, _p]),
    'nb_mse_grad': (C.c_int, [_p, _i64, _p, _p, _f32, _f32, _p, _p, _p]),
    'nb_adam_step': (C.c_int, [_p, _i64, _p, _p, _p, _p, _f32, _f32, _f32, _f32, _i32, _p]),
    'nb_adam_step_sum': (C.c_int, [_p, _i64, _p, _p, _p, _i32, _p, _p, _f32, _f32, _f32, _f32, _i32, _p]),
}

_lib = None


def build(verbose=False):
    """Compile libnerf_b200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(['make', '-C', CSRC, '-j8'], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise NBError('building libnerf_b200.so failed')
    return LIB_PATH


def load():
    """dlopen the library and attach signatures.  Raises NBError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NBError(f'{LIB_PATH} is missing: run `make -C {CSRC}` (there is no CPU fallback)')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib

"""ctypes binding of libfusionfpn.so (include/ffpn.h).  There is no fallback: if the library is
missing or a CUDA device is not an sm_100 part, every op raises."""
import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# FFPN_LIB=debug loads the -DFFPN_DEBUG build (make DEBUG=1): only that build honours the result-invalidating
# ablation switches; bench.py refuses to report numbers from it
DEBUG_LIB = os.environ.get('FFPN_LIB', '') == 'debug'
LIB_PATH = os.path.join(_HERE, 'libfusionfpn_dbg.so' if DEBUG_LIB else 'libfusionfpn.so')

F32, BF16 = 0, 1
STAT_ROWS = 1184


class ConvDesc(C.Structure):
    """Mirror of ffpn_conv_desc (include/ffpn.h)."""
    _fields_ = [('B', C.c_int64), ('S', C.c_int64), ('W', C.c_int64), ('H', C.c_int64),
                ('oS', C.c_int64), ('oW', C.c_int64), ('oH', C.c_int64),
                ('Cin', C.c_int32), ('Cout', C.c_int32),
                ('kS', C.c_int32), ('kW', C.c_int32), ('kH', C.c_int32),
                ('sS', C.c_int32), ('sW', C.c_int32), ('sH', C.c_int32),
                ('pS', C.c_int32), ('pW', C.c_int32), ('pH', C.c_int32),
                ('dtype', C.c_int32), ('impl', C.c_int32)]


_P, _I, _L, _F, _D, _Z = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t
_IP = C.POINTER(C.c_int)
_DP = C.POINTER(ConvDesc)

# name -> argtypes (after the leading ctx*); restype is int unless listed in _RESTYPE
SIGNATURES = {
    'ffpn_conv_fwd': [_DP, _P, _P, _P, _I, _P, _P, _P, _IP, _P, _Z, _P],
    'ffpn_conv_fwd_bn': [_DP, _P, _P, _P, _I, _P, _P, _P, _IP, _D, _P, _P, _P, _P, _F, _F, _I, _P, _P, _P, _P, _P, _Z, _P],
    'ffpn_conv_dgrad': [_DP, _P, _P, _P, _P, _P, _Z, _P],
    'ffpn_conv_dgrad_bnr': [_DP, _P, _P, _P, _P, _P, _P, _P, _IP, _P, _Z, _P],
    'ffpn_conv_wgrad': [_DP, _P, _P, _P, _I, _P, _P, _P, _Z, _P],
    'ffpn_bn_finalize': [_P, _I, _I, _D, _P, _P, _P, _P, _F, _F, _I, _P, _P, _P, _P, _P],
    'ffpn_bn_bwd_reduce': [_I, _L, _I, _P, _P, _P, _P, _I, _P, _IP, _P],
    'ffpn_bn_bwd_finalize': [_P, _I, _I, _I, _I, _D, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    'ffpn_bn_bwd_apply': [_I, _L, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P],
    'ffpn_bn_bwd_apply2': [_I, _L, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    'ffpn_block_end_fwd': [_I, _L, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    'ffpn_block_end_bwd': [_I, _L, _L, _L, _L, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _IP, _P],
    'ffpn_maxpool_fwd': [_I, _L, _L, _L, _L, _I, _I, _I, _I, _P, _P, _P, _P],
    'ffpn_maxpool_bwd': [_I, _L, _L, _L, _L, _I, _I, _I, _I, _P, _P, _P, _P],
    'ffpn_proj_tail_fwd': [_I, _L, _L, _I, _P, _P, _P, _P, _I, _I, _P],
    'ffpn_proj_tail_bwd': [_I, _L, _L, _I, _P, _I, _I, _P, _P],
    'ffpn_resize2d_fwd': [_I, _I, _L, _L, _L, _L, _L, _I, _P, _P, _I, _I, _P, _P],
    'ffpn_resize2d_bwd': [_I, _I, _L, _L, _L, _L, _L, _I, _P, _I, _I, _P, _P, _P],
    'ffpn_upsample_fwd': [_I, _L, _L, _L, _I, _I, _I, _P, _P, _I, _I, _P],
    'ffpn_upsample_bwd': [_I, _L, _L, _L, _I, _I, _I, _P, _I, _I, _P, _P],
    'ffpn_slice_copy': [_I, _L, _I, _P, _I, _I, _P, _I, _I, _P],
    'ffpn_head_fwd': [_I, _L, _L, _I, _I, _I, _P, _P, _P, _P, _P],
    'ffpn_head_bwd': [_I, _L, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P],
    'ffpn_mix_loss_fwd': [_L, _I, _L, _P, _P, _P, _Z, _P, _P],
    'ffpn_mix_loss_bwd': [_L, _I, _L, _P, _P, _P, _P, _P, _P],
    'ffpn_pack_volume': [_I, _L, _L, _L, _P, _P, _P],
    'ffpn_cast': [_I, _L, _P, _P, _P],
    'ffpn_sgd_step': [_L, _P, _P, _P, _F, _F, _F, _F, _I, _P],
    'ffpn_zscore_slices': [_L, _L, _P, _F, _P, _P, _P],
    'ffpn_dice_metric': [_L, _I, _L, _I, _F, _F, _P, _P, _P, _P],
    'ffpn_weight_arena_begin': [_P, _Z],
    'ffpn_weight_arena_seal': [],
    'ffpn_weight_arena_pack': [_P],
    'ffpn_weight_arena_end': [],
    'ffpn_weight_arena_enable': [_I],
    'ffpn_route_counts': [C.POINTER(C.c_int64)],
}
NO_CTX = {
    'ffpn_abi_version': ([], C.c_int),
    'ffpn_build_info': ([], C.c_int),
    'ffpn_create': ([C.POINTER(C.c_void_p), _I], C.c_int),
    'ffpn_destroy': ([_P], None),
    'ffpn_last_error': ([_P], C.c_char_p),
    'ffpn_launch_count': ([_P], C.c_int64),
    'ffpn_conv_workspace_bytes': ([_DP], C.c_size_t),
    'ffpn_conv_plan_info': ([_DP, _I, C.c_char_p, _Z], C.c_int),
    'ffpn_head_bwd_workspace_bytes': ([_I, _I], C.c_size_t),
    'ffpn_mix_loss_workspace_bytes': ([_I], C.c_size_t),
}
EXPORTS = sorted(list(SIGNATURES) + list(NO_CTX))

_lib = None
_ctx = {}
_lock = threading.Lock()


class FfpnError(RuntimeError):
    pass


def load():
    """Load libfusionfpn.so (once).  Raises if it has not been built -- there is no other path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise FfpnError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                            '(or `make -C multimodal-fusion-fpn_b200/csrc`). There is no CPU/PyTorch fallback.')
        lib = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = [_P] + args
            fn.restype = C.c_int
        for name, (args, res) in NO_CTX.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = res
        if lib.ffpn_abi_version() != 3:
            raise FfpnError('libfusionfpn.so ABI version mismatch')
        _lib = lib
    return _lib


def ctx(device_index: int):
    """Per-device ffpn_ctx*, created on first use."""
    lib = load()
    h = _ctx.get(device_index)
    if h is None:
        with _lock:
            h = _ctx.get(device_index)
            if h is None:
                if not torch.cuda.is_available():
                    raise FfpnError('fusion FPN kernels need a CUDA device (sm_100a); there is no CPU fallback')
                out = C.c_void_p()
                rc = lib.ffpn_create(C.byref(out), device_index)
                if rc != 0:
                    raise FfpnError(f'ffpn_create(device={device_index}) failed with code {rc} '
                                    '(4 = device is not an sm_100 part)')
                h = out
                _ctx[device_index] = h
    return h


# timing experiments only (results invalid): honoured by the debug build alone
_SKIP = set(filter(None, os.environ.get('FFPN_TIMING_SKIP', '').split(','))) if DEBUG_LIB else set()


def call(name: str, device_index: int, *args):
    if _SKIP and name in _SKIP:
        return
    lib = load()
    h = ctx(device_index)
    rc = getattr(lib, name)(h, *args)
    if rc != 0:
        raise FfpnError(f'{name}: {lib.ffpn_last_error(h).decode()}')


def launch_count(device_index: int = 0) -> int:
    return int(load().ffpn_launch_count(ctx(device_index)))


ROUTES = ('tcgen05_ws', 'stem', 'tcgen05_gen1', 'cuda_core')


def route_counts(device_index: int = 0) -> dict:
    """Conv calls per kernel family since the ctx was created (ffpn_route_counts)."""
    out = (C.c_int64 * 4)()
    call('ffpn_route_counts', device_index, out)
    return dict(zip(ROUTES, [int(v) for v in out]))


def is_debug_build() -> bool:
    return bool(load().ffpn_build_info() & 1)


def dtype_code(t: torch.dtype) -> int:
    if t == torch.float32:
        return F32
    if t == torch.bfloat16:
        return BF16
    raise FfpnError(f'unsupported activation dtype {t}')

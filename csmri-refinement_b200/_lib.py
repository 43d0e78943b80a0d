"""ctypes binding of libcsmri_dc.so (the C ABI of include/csmri_dc.h).

There is no fallback: if the shared library has not been built, importing any
op raises immediately with the build command.
"""
import ctypes
import glob
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIB_PATH = os.path.join(CSRC, 'libcsmri_dc.so')
MAIN_SOURCE = 'csmri_dc.cu'
NVCC_FLAGS = ['-std=c++17', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a',
              '-lineinfo', '-Xcompiler', '-fPIC', '-shared']

_c_float_p = ctypes.c_void_p   # device pointers travel as integers
_lib = None


def sources():
    """Everything the library is compiled from: every .cu / .cuh under csrc/
    (csmri_dc.cu includes all the headers) plus the public ABI header."""
    srcs = sorted(glob.glob(os.path.join(CSRC, '*.cu')) + glob.glob(os.path.join(CSRC, '*.cuh')))
    srcs.append(os.path.join(os.path.dirname(_HERE), 'include', 'csmri_dc.h'))
    return srcs


def is_stale():
    """True when the .so is missing or older than any of its sources."""
    if not os.path.exists(LIB_PATH):
        return True
    return os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in sources())


def build(force=False, verbose=False):
    """Compile csrc/csmri_dc.cu for sm_100a into csrc/libcsmri_dc.so."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', LIB_PATH, os.path.join(CSRC, MAIN_SOURCE)]
    if verbose:
        cmd += ['-Xptxas', '-v']
        print(' '.join(cmd), file=sys.stderr)
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB_PATH


_SIGNATURES = {
    'csmri_version': (ctypes.c_int, []),
    'csmri_last_error': (ctypes.c_char_p, []),
    'csmri_init': (ctypes.c_int, []),
    'csmri_dc_workspace_bytes': (ctypes.c_size_t, [ctypes.c_int] * 3),
    'csmri_dc_prepare': (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_float, _c_float_p, _c_float_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'csmri_dc_prepare_lines': (ctypes.c_int, [_c_float_p, ctypes.c_void_p] + [ctypes.c_int] * 4 +
                               [ctypes.c_float, _c_float_p, _c_float_p, ctypes.c_void_p,
                                ctypes.c_void_p]),
    'csmri_dc_forward_cartesian': (ctypes.c_int, [_c_float_p] * 5 + [ctypes.c_int] * 3 +
                                   [ctypes.c_void_p]),
    'csmri_dc_adjoint_cartesian': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int] * 3 +
                                   [ctypes.c_void_p]),
    'csmri_dc_forward_general': (ctypes.c_int, [_c_float_p] * 5 + [ctypes.c_int] * 3 +
                                 [ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    'csmri_dc_adjoint_general': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int] * 3 +
                                 [ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    'csmri_dc_forward': (ctypes.c_int, [_c_float_p] * 7 + [ctypes.c_int] * 3 +
                         [ctypes.c_float, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    'csmri_dc_adjoint': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_int] * 3 +
                         [ctypes.c_float, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    'csmri_undersample': (ctypes.c_int, [_c_float_p, ctypes.c_void_p] + [_c_float_p] * 6 +
                          [ctypes.c_int] * 3 + [ctypes.c_void_p, ctypes.c_void_p]),
    'csmri_fft2': (ctypes.c_int, [_c_float_p, _c_float_p] + [ctypes.c_int] * 4 +
                   [ctypes.c_void_p, ctypes.c_void_p]),
    'csmri_magnitude_clamp': (ctypes.c_int, [_c_float_p, _c_float_p] + [ctypes.c_int] * 3 +
                              [ctypes.c_float, ctypes.c_float, ctypes.c_void_p]),
    'csmri_psnr_sum': (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_void_p] + [ctypes.c_int] * 3 +
                       [ctypes.c_float, ctypes.c_float, ctypes.c_void_p]),
    'csmri_shift_crop': (ctypes.c_int, [_c_float_p, _c_float_p] + [ctypes.c_int] * 13 +
                         [_c_float_p, ctypes.c_void_p]),
    'csmri_plane_absmax': (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p]),
    'csmri_plane_divide': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int, ctypes.c_int,
                                                             ctypes.c_void_p]),
    'csmri_plane_minmax': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int, ctypes.c_int,
                                                             ctypes.c_longlong, ctypes.c_void_p]),
    'csmri_plane_scale': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_int, ctypes.c_int,
                                                            ctypes.c_longlong, ctypes.c_longlong,
                                                            ctypes.c_int, ctypes.c_void_p]),
    'csmri_refine_real_penalty_add': (ctypes.c_int, [_c_float_p] * 6 + [ctypes.c_int] * 3 +
                                      [ctypes.c_void_p]),
    'csmri_refine_partials': (ctypes.c_int, []),
    'csmri_refine_real_penalty_add_backward': (ctypes.c_int, [_c_float_p] * 6 + [ctypes.c_int] * 3 +
                                               [ctypes.c_void_p]),
    'csmri_conv3x3_wgrad_workspace_bytes': (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    'csmri_conv3x3_wgrad': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_void_p] + [ctypes.c_int] * 6 +
                            [ctypes.c_void_p]),
    'csmri_conv3x3_wgrad_thin_bias': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_void_p] + [ctypes.c_int] * 3 +
                                      [ctypes.c_void_p]),
    'csmri_conv3x3_wgrad_thin_in_bias': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_void_p] + [ctypes.c_int] * 3 +
                                      [ctypes.c_void_p]),
    'csmri_conv3x3_thin_dgrad': (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_void_p, _c_float_p] +
                                 [ctypes.c_int] * 5 + [ctypes.c_float, ctypes.c_void_p]),
    'csmri_conv3x3_thin_masked': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_int] * 3 +
                                  [ctypes.c_float, ctypes.c_void_p]),
    'csmri_conv3x3_wgrad_bias': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_void_p] + [ctypes.c_int] * 3 +
                                 [ctypes.c_void_p]),
    'csmri_conv3x3_thin': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_int] * 5 +
                           [ctypes.c_float, ctypes.c_void_p]),
    'csmri_conv3x3_tc': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_int] * 4 +
                         [ctypes.c_float, ctypes.c_int, ctypes.c_void_p]),
    'csmri_conv3x3_tc_signs': (ctypes.c_int, [_c_float_p] * 6 + [ctypes.c_int] * 4 +
                               [ctypes.c_float, ctypes.c_void_p]),
    'csmri_conv3x3_tc_masked': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_int] * 4 +
                                [ctypes.c_float, ctypes.c_int, ctypes.c_void_p]),
    'csmri_bias_lrelu': (ctypes.c_int, [_c_float_p] * 2 + [ctypes.c_int] * 4 +
                         [ctypes.c_float, ctypes.c_void_p]),
    'csmri_bias_lrelu_backward': (ctypes.c_int, [_c_float_p] * 5 + [ctypes.c_int] * 4 +
                                  [ctypes.c_float, ctypes.c_void_p]),
    # tuning knob used by bench.py only (not declared in include/csmri_dc.h)
    'csmri_set_variant': (ctypes.c_int, [ctypes.c_int]),
    'csmri_set_tuning': (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    'csmri_set_trace': (ctypes.c_int, [ctypes.c_void_p]),
}

# every symbol include/csmri_dc.h declares
ABI_SYMBOLS = [k for k in _SIGNATURES if k not in ('csmri_set_variant', 'csmri_set_tuning', 'csmri_set_trace')]


def lib():
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'libcsmri_dc.so is missing (%s). Build it with '
                '`python -c "import __graft_entry__ as g; g.build()"`. There is no '
                'CPU/PyTorch fallback for the DC path.' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().csmri_last_error().decode('utf-8', 'replace')
        raise RuntimeError('libcsmri_dc error %d: %s' % (rc, msg))

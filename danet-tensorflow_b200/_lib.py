"""
ctypes binding of libdanet_sm100.so (include/danet.h).  There is NO fallback: if the
library is missing the import of any compute entry point raises, loudly.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'lib', 'libdanet_sm100.so')

c_f = C.c_void_p      # device float*
c_i = C.c_int
c_ll = C.c_longlong
c_sz = C.c_size_t
c_v = C.c_void_p

# name -> (restype, argtypes); mirrors include/danet.h declaration by declaration
PROTOTYPES = {
    'danet_version': (c_i, []),
    'danet_last_error_string': (C.c_char_p, []),
    'danet_check_device': (c_i, []),
    'danet_timestamp': (c_i, [c_v, c_v]),
    'danet_crc32c': (C.c_uint, [c_v, c_sz, C.c_uint]),
    'danet_stft_num_frames': (c_i, [c_i]),
    'danet_stft_fwd': (c_i, [c_f, c_i, c_i, c_f, c_f, c_v]),
    'danet_mix_features_fwd': (c_i, [c_f, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_v]),
    'danet_center_workspace_bytes': (c_sz, [c_i]),
    'danet_center_fwd': (c_i, [c_f, c_i, c_ll, c_f, c_f, c_v]),
    'danet_leaky_relu_bwd': (c_i, [c_f, c_f, c_f, c_ll, C.c_float, c_v]),
    'danet_leaky_relu_fwd': (c_i, [c_f, c_f, c_ll, C.c_float, c_v]),
    'danet_linear_workspace_bytes': (c_sz, [c_i, c_i, c_i, c_i]),
    'danet_linear_fwd': (c_i, [c_f, c_ll, c_f, c_ll, c_f, c_f, c_i, c_i, c_i, c_i, c_v, c_sz, c_i, c_v]),
    'danet_gemm_workspace_bytes': (c_sz, [c_i, c_i, c_i]),
    'danet_gemm': (c_i, [c_f, c_ll, c_i, c_i, c_i, c_f, c_ll, c_i, c_f, c_f, c_ll, c_i, c_i, c_i, c_i, c_i,
                         c_v, c_sz, c_v]),
    'danet_lstm_seq_workspace_bytes': (c_sz, [c_i, c_i, c_i]),
    'danet_lstm_seq_fwd': (c_i, [c_f, c_ll, c_ll, C.POINTER(C.c_void_p), c_ll, c_f, c_f, c_f, c_v, c_i,
                                 c_i, c_i, c_i, c_i, c_v, c_sz, c_i, c_v]),
    'danet_lstm_pack_wh_bytes': (c_sz, [c_i, c_i]),
    'danet_lstm_pack_wh': (c_i, [C.POINTER(C.c_void_p), c_ll, c_i, c_i, c_v, c_sz, c_v]),
    'danet_lstm_seq_fwd_packed': (c_i, [c_f, c_ll, c_ll, C.POINTER(C.c_void_p), c_ll, c_v, c_f, c_f, c_f, c_v, c_i,
                                        c_i, c_i, c_i, c_i, c_v, c_sz, c_i, c_v]),
    'danet_split_operand_bytes': (c_sz, [c_i, c_i]),
    'danet_split_operand': (c_i, [c_f, c_ll, c_i, c_i, c_i, c_v, c_i, c_i, c_v]),
    'danet_split_operand_paired': (c_i, [c_f, c_ll, c_i, c_i, c_i, c_i, c_v, c_i, c_i, c_v]),
    'danet_gemm_split': (c_i, [c_v, c_v, c_f, c_f, c_f, c_i, c_f, c_ll, c_i, c_i, c_i, c_i, c_i, c_v]),
    'danet_zero_async': (c_i, [c_v, c_sz, c_v]),
    'danet_split_operand_time_major': (c_i, [c_f, c_ll, c_i, c_i, c_i, c_v, c_v]),
    'danet_gemm_split_pipelined': (c_i, [c_v, c_v, c_f, c_f, c_ll, c_i, c_i, c_i, c_i, c_i, c_v, C.POINTER(C.c_int), c_v]),
    'danet_lstm_seq_fwd_pipelined': (c_i, [c_f, c_ll, c_ll, C.POINTER(C.c_void_p), c_ll, c_v, c_f, c_v, c_i, c_i, c_i, c_i, c_i,
                                           c_v, c_i, c_i, c_i, c_i, c_v, c_sz, c_i, c_v]),
    'danet_proj_anchor_workspace_bytes': (c_sz, [c_i, c_i, c_i, c_i]),
    'danet_proj_anchor_fwd': (c_i, [c_v, c_v, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_v, c_i, c_i, c_i, c_i, c_i, c_i,
                                    c_v, c_sz, c_v]),
    'danet_mean_fwd': (c_i, [c_f, c_i, c_ll, c_f, c_f, c_v]),
    'danet_lstm_seq_bwd_workspace_bytes': (c_sz, [c_i, c_i, c_i]),
    'danet_lstm_seq_bwd': (c_i, [c_f, c_f, c_f, C.POINTER(C.c_void_p), c_ll, c_i, c_i, c_i, c_i, c_v, c_sz, c_i, c_v]),
    'danet_colsum_workspace_bytes': (c_sz, [c_i]),
    'danet_colsum': (c_i, [c_f, c_ll, c_ll, c_i, c_f, c_i, c_v, c_sz, c_v]),
    'danet_clip_sgd': (c_i, [c_f, c_f, c_ll, C.c_float, C.c_float, C.c_float, c_v]),
    'danet_clip_adam': (c_i, [c_f, c_f, c_f, c_f, c_ll, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                              C.c_float, c_i, c_v]),
    'danet_attractor_workspace_bytes': (c_sz, [c_i, c_i, c_i]),
    'danet_attractor_truth_fwd': (c_i, [c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_v, c_sz, c_v]),
    'danet_anchor_num_subsets': (c_i, [c_i, c_i]),
    'danet_attractor_anchor_fwd': (c_i, [c_f, c_f, c_f, c_f, c_f, c_v, c_f, c_i, c_i, c_i, c_i, c_i,
                                         c_v, c_sz, c_v]),
    'danet_head_bwd_workspace_bytes': (c_sz, [c_i, c_i, c_i]),
    'danet_head_bwd_attractors': (c_i, [c_f, c_f, c_f, c_f, c_v, c_f, c_i, c_i, c_i, c_i, c_i, c_v, c_sz, c_v]),
    'danet_head_bwd_embed': (c_i, [c_f, c_f, c_f, c_f, c_v, c_f, c_i, c_f, c_f, c_f, c_v, c_f, c_f, c_f,
                                   c_i, c_i, c_i, c_i, c_i, c_i, c_v, c_sz, c_v]),
    'danet_attractor_kmeans_fwd': (c_i, [c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_v, c_sz, c_v]),
    'danet_mask_cmul_fwd': (c_i, [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_v]),
    'danet_istft_fwd': (c_i, [c_f, c_i, c_i, c_f, c_v]),
    'danet_conv2d_fwd': (c_i, [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, C.c_float, c_v]),
    'danet_maxpool2x2_fwd': (c_i, [c_f, c_f, c_ll, c_i, c_i, c_v]),
    'danet_conv2d_bwd_data': (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_v]),
    'danet_conv2d_bwd_weights': (c_i, [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_v]),
    'danet_maxpool2x2_bwd': (c_i, [c_f, c_f, c_f, c_ll, c_i, c_i, c_v]),
    'danet_add_fwd': (c_i, [c_f, c_f, c_f, c_ll, c_v]),
    'danet_mask_cmul_istft_fwd': (c_i, [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_v]),
    'danet_pit_workspace_bytes': (c_sz, [c_i, c_i]),
    'danet_pit_mse_fwd': (c_i, [c_f, c_f, c_i, c_i, c_i, c_i, c_f, c_f, c_v, c_f, c_f, c_v, c_sz, c_v]),
}

ERROR_NAMES = {-1: 'DANET_E_SHAPE', -2: 'DANET_E_ALIGN', -3: 'DANET_E_ARCH', -4: 'DANET_E_CUDA',
               -5: 'DANET_E_ARG', -6: 'DANET_E_WORKSPACE'}

_lib = None


def load():
    """dlopen the library once and attach the prototypes"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'libdanet_sm100.so is missing (%s). Build it with `python -c "import __graft_entry__ as g; '
            'g.build()"`; there is no CPU or PyTorch fallback for the DANet hot path.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    """Map the C-ABI's error convention onto the reference's Python one:
    shape / argument problems -> ValueError, everything else -> RuntimeError"""
    if rc == 0:
        return
    msg = '%s failed: %s (%s)' % (what, load().danet_last_error_string().decode(), ERROR_NAMES.get(rc, rc))
    if rc in (-1, -2, -5, -6):
        raise ValueError(msg)
    raise RuntimeError(msg)

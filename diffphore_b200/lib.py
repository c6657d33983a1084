"""ctypes binding of libdiffphore_sm100.so (include/diffphore_b200.h).

The library is built in-tree by __graft_entry__.build() / `python -m diffphore_b200.build`.  There is NO fallback:
if the shared object is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DIFFPHORE_LIB: another build of the same library (A/B runs of kernel variants on one box); default: the in-tree build
LIB_PATH = os.environ.get('DIFFPHORE_LIB') or os.path.join(_HERE, 'libdiffphore_sm100.so')

c_fp = C.c_void_p       # device pointers travel as integers
i32 = C.c_int32


class DpMlp(C.Structure):
    _fields_ = [('w0', C.c_void_p), ('b0', C.c_void_p), ('w3', C.c_void_p), ('b3', C.c_void_p)]


class DpSmallWeights(C.Structure):
    _fields_ = [(n, DpMlp) for n in ('lig_edge', 'pp_edge', 'cross_edge', 'cdt', 'pdt', 'pmt', 'center_edge',
                                     'final_edge', 'tr_final', 'rot_final')] + \
               [(n, C.c_void_p) for n in ('boarder_tables', 'boarder_w', 'boarder_b', 'tor_w0', 'tor_w3')]


class DpConstants(C.Structure):
    _fields_ = [('rbf_mu', (C.c_float * 20) * 4), ('rbf_coeff', C.c_float * 4), ('clash_cutoff', C.c_float * 5),
                ('lig_radius', C.c_float), ('scaler', C.c_float), ('max_neighbors', i32), ('no_clamp', i32)]


# offsets into the per-noise-level constant block (enum in the header)
SC = dict(SEMB=0, LIG_NODE=20, PH_NODE=40, LIG_EDGE=60, PP_EDGE=80, CROSS_EDGE=100, CENTER=120, TR=140, ROT=160,
          INV_TR_SIGMA=180, SO3_NORM=181, SQRT_TORUS=182, TR_A=183, TR_B=184, ROT_A=185, ROT_B=186, TOR_A=187,
          TOR_B=188, SIZE=256)
TP_L0, TP_L1, TP_L2, TP_L3, TP_FINAL, TP_TOR = range(6)

P = C.c_void_p
_SIGNATURES = {
    'dp_version': ([], i32),
    'dp_set_constants': ([C.POINTER(DpConstants)], i32),
    'dp_lig_graph': ([P, P, P, P, P, i32, i32, i32, C.POINTER(DpSmallWeights), P, P, P, P, P, P, P, P, P, P, P, P], i32),
    'dp_pp_setup': ([P, P, P, i32, C.POINTER(DpSmallWeights), P, P, P], i32),
    'dp_pp_step': ([P, i32, C.POINTER(DpSmallWeights), P, P, P], i32),
    'dp_cross_setup': ([P, P, i32, P, P, C.POINTER(DpSmallWeights), P, P, P], i32),
    'dp_cross_step': ([P, P, P, P, P, P, P, i32, i32, P, P, P, P, P, P, C.POINTER(DpSmallWeights), P, P, P, P, P, P], i32),
    'dp_node_embed': ([P, P, P, P, P, P, P, i32, i32, C.POINTER(DpSmallWeights), P, P, P, P], i32),
    'dp_edge_mlp': ([P, P, P, P, i32, P, P, P, i32, P, P, P, i32, i32, i32, P, i32, P, P], i32),
    'dp_edge_mlp_tc': ([P, P, P, P, i32, P, P, P, i32, P, P, P, C.c_float, i32, i32, i32, P, i32, P, P, P], i32),
    'dp_build_tiles': ([P, P, i32, P, P, P, P, P], i32),
    'dp_conv_fused': ([i32, P, P, P, P, i32, P, P, P, i32, P, C.c_float, P, C.c_float, P, P, P, i32, P, P, P, i32, P, P, P, P, i32, i32, P], i32),
    'dp_conv_fused_flat': ([i32, P, P, P, P, i32, P, P, P, i32, P, C.c_float, P, C.c_float, P, P, P, i32, P, P, P, i32, P, P, P, P, i32, i32, P], i32),
    'dp_conv_fused2': ([i32, P, P, P, P, i32, P, P, P, i32, P, C.c_float, P, C.c_float, P, P, P, i32, P, P, P, i32, P, P, P, P, i32, i32, P], i32),
    'dp_tp_scatter': ([i32, P, P, P, P, i32, P, P, P, P, P, P, i32, i32, i32, P], i32),
    'dp_center_step': ([P, P, i32, C.POINTER(DpSmallWeights), P, P, P, P], i32),
    'dp_score_head': ([P, i32, C.POINTER(DpSmallWeights), P, P, P, P], i32),
    'dp_tor_graph': ([P, P, P, P, P, i32, i32, C.POINTER(DpSmallWeights), P, P, P, P, P, P, P, P, P, P, P], i32),
    'dp_tor_head': ([P, i32, C.POINTER(DpSmallWeights), P, P, P], i32),
    'dp_conformer_update': ([P, P, P, P, P, P, P, P, i32, i32, i32, P, P, P, P, P, P, P, i32, P], i32),
    'dp_randomize_position': ([P, P, P, P, P, P, P, P, i32, i32, i32, P, P, P, i32, P], i32),
}
EXPORTS = ['dp_last_error'] + list(_SIGNATURES)

_lib = None


def load():
    """Load the CUDA library; raise (never fall back) if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                           '(there is no CPU fallback for the denoising path)')
    lib = C.CDLL(LIB_PATH)
    lib.dp_last_error.restype = C.c_char_p
    lib.dp_last_error.argtypes = []
    for name, (args, res) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, res
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        raise RuntimeError(f'{what} failed: {load().dp_last_error().decode()}')


def ptr(t):
    """Device (or host) address of a torch tensor, None -> NULL."""
    if t is None:
        return None
    assert t.is_contiguous(), 'C ABI needs contiguous tensors'
    return t.data_ptr()

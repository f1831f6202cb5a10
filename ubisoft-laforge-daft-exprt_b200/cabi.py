"""ctypes binding of the C-ABI declared in include/daft_exprt_b200.h.

The prototypes are parsed from the header itself, so the binding cannot drift from the declared ABI; the CPU test-suite
checks that the built shared library exports every declared symbol.  There is NO fallback: if the library is missing or a
call fails, a RuntimeError is raised (the product path never routes through PyTorch math or the oracle).
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), 'include', 'daft_exprt_b200.h')
LIB_PATH = os.path.join(_HERE, 'libdaftexprt_b200.so')

DX_GEMM_FP32_CUDA_CORES = 0
DX_GEMM_TCGEN05_TF32 = 1
DX_GEMM_TCGEN05_BF16X3 = 2
DX_ATTENTION_MMA_SYNC = 0
DX_ATTENTION_TCGEN05 = 1

_lib = None
_protos = None


def _ctype_of(decl):
    decl = decl.strip()
    if '*' in decl:
        return ctypes.c_void_p
    base = decl.rsplit(' ', 1)[0].replace('const', '').strip() if ' ' in decl else decl
    return {'int': ctypes.c_int, 'float': ctypes.c_float, 'size_t': ctypes.c_size_t, 'int64_t': ctypes.c_int64,
            'uint64_t': ctypes.c_uint64}[base]


def parse_header(path=HEADER):
    """Return {name: (restype, [argtypes])} for every `dx_*` prototype in the header."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    protos = {}
    for m in re.finditer(r'(const char\*|int|size_t|uint64_t)\s+(dx_\w+)\s*\(([^;{]*?)\)\s*;', text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), ' '.join(m.group(3).split())
        restype = {'const char*': ctypes.c_char_p, 'int': ctypes.c_int, 'size_t': ctypes.c_size_t, 'uint64_t': ctypes.c_uint64}[ret]
        argtypes = [] if args in ('', 'void') else [_ctype_of(a) for a in args.split(',')]
        protos[name] = (restype, argtypes)
    return protos


def load(path=LIB_PATH):
    """Load libdaftexprt_b200.so and attach prototypes.  Raises loudly when the library has not been built."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(f'{path} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                           f'(or `python ubisoft-laforge-daft-exprt_b200/build.py`). There is no CPU/PyTorch fallback.')
    lib = ctypes.CDLL(path)
    _protos = parse_header()
    for name, (restype, argtypes) in _protos.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.dx_abi_version() != 1:
        raise RuntimeError('libdaftexprt_b200.so ABI version mismatch')
    _lib = lib
    return lib


def last_error():
    return load().dx_last_error().decode()


def check(rc, what=''):
    if rc != 0:
        raise RuntimeError(f'daft_exprt_b200 C-ABI call {what} failed (rc={rc}): {last_error()}')

"""CPU-side checks of the drop-in boundary (no compute calls): the shared library loads, exports every function that
include/ctgan_sm100.h declares, the ctypes prototypes of ctgan_b200/_lib.py cover exactly that set, and the product fails
loudly when the library is missing (there is no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
HEADER = os.path.join(ROOT, 'include', 'ctgan_sm100.h')
LIB = os.path.join(ROOT, 'ctgan_b200', 'libctgan_sm100.so')


def _declared():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)           # comments mention entry points too
    return sorted(set(re.findall(r'\b(ctgan_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 50
    lib = ctypes.CDLL(LIB)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_prototypes_match_header():
    from ctgan_b200 import _lib
    declared = set(_declared())
    bound = set(_lib.EXPORTS)
    # every bound prototype is declared in the header; every declared function is reachable from Python
    assert not (bound - declared), sorted(bound - declared)
    assert not (declared - bound), sorted(declared - bound)
    for n in declared:
        assert hasattr(_lib.lib, n), n


def test_error_reporting_without_a_gpu():
    lib = ctypes.CDLL(LIB)
    lib.ctgan_version.restype = ctypes.c_int
    lib.ctgan_last_error.restype = ctypes.c_char_p
    lib.ctgan_tc_available.restype = ctypes.c_int
    assert lib.ctgan_version() > 0
    assert lib.ctgan_tc_available() in (0, 1)
    # a descriptor-level error (null descriptor) is reported through the return code + ctgan_last_error, no launch involved
    lib.ctgan_conv_fprop.restype = ctypes.c_int
    rc = lib.ctgan_conv_fprop(None, None, None, None, None, 0, None)
    assert rc != 0 and len(lib.ctgan_last_error()) > 0


def test_missing_library_fails_loudly():
    env = dict(os.environ, CTGAN_SM100_LIB='/nonexistent/libctgan_sm100.so', PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, '-c', 'import ctgan_b200.kernels'], env=env, capture_output=True, text=True)
    assert r.returncode != 0
    assert 'libctgan_sm100' in (r.stderr + r.stdout)

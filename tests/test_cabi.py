"""CPU-side checks of the drop-in boundary: the library loads and exports every symbol the header declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'uppasd_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    names = re.findall(r'^\s*(?:const\s+)?(?:void|int|long|char\s*\*|const char\s*\*|asd_engine\s*\*)\s*\*?\s*(\w+)\s*\(', src, flags=re.M)
    return sorted(set(n for n in names if not n.startswith('asd_cb_')))


def test_library_exports_every_declared_symbol():
    from uppasd_b200 import build, capi
    build.build()
    lib = capi.load()
    decl = _declared_symbols()
    assert len(decl) >= 35
    for name in decl:
        assert hasattr(lib, name), name
    # the bindings cover exactly the header
    assert sorted(capi.SYMBOLS) == decl


def test_no_cpu_fallback_without_a_gpu():
    """Without a CUDA device every compute entry must fail loudly (asd_create refuses to build an engine)."""
    import ctypes as C
    from uppasd_b200 import capi
    lib = capi.load()
    if lib.asd_device_count() > 0:
        pytest.skip('a GPU is present')
    h = C.c_void_p()
    assert lib.asd_create(C.byref(h), -1) != 0
    assert b'no CPU fallback' in lib.asd_last_error()


def test_product_never_touches_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, 'uppasd_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.inl', '.h')):
                txt = open(os.path.join(base, f)).read()
                assert 'oracle' not in txt, os.path.join(base, f)

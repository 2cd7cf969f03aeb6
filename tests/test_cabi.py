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


def _build_example(tmp_path):
    import subprocess
    exe = os.path.join(str(tmp_path), 'sd_minimal')
    lib = os.path.join(ROOT, 'uppasd_b200')
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-I', os.path.join(ROOT, 'include'),
                    os.path.join(ROOT, 'examples', 'sd_minimal.c'), '-L', lib, '-luppasd_b200', '-Wl,-rpath,' + lib, '-o', exe],
                   check=True, capture_output=True)
    return exe


def test_header_is_plain_c_and_the_c_example_links(tmp_path):
    """include/uppasd_b200.h is what a C or Fortran host binds: it must compile as strict C99, and a C program using only
    that header must link against the library; without a GPU it stops with the library's own error, no CPU fallback."""
    import subprocess
    from uppasd_b200 import build, capi
    build.build()
    exe = _build_example(tmp_path)
    if capi.load().asd_device_count() > 0:
        pytest.skip('a GPU is present: the GPU test runs the example')
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and 'no CPU fallback' in r.stderr


@pytest.mark.gpu
def test_c_example_matches_the_python_host(tmp_path):
    import subprocess
    import numpy as np
    from uppasd_b200 import host
    out = subprocess.run([_build_example(tmp_path)], capture_output=True, text=True, check=True).stdout
    nums = [float(x) for x in re.findall(r'-?\d+\.\d+', out)]
    e = host.Engine()
    e.set_constants(1.760859644e11, 1.38064852e-23, 9.274009994e-24, 2.179872325e-21)
    e.set_system(2, 1, 2, np.array([1, 2], dtype=np.int32))
    e.set_exchange(np.array([[2, 1]], dtype=np.int32, order='F'), np.array([1, 1], dtype=np.int32), np.array([[10.0, 10.0]], order='F'))
    e.set_llg(1, 1e-16, landeg=1.0, lambda1=0.1, temp=0.0, seed=1)
    emom = np.zeros((3, 2, 1), order='F'); emom[0, 0, 0] = 1.0; emom[1, 1, 0] = 1.0
    e.set_moments(emom, np.ones((2, 1), order='F'))
    e.commit()
    e.sd_steps(100)
    em = e.get_moments()[0]
    assert np.abs(np.array(nums[:3]) - em[:, 0, 0]).max() <= 1e-12
    assert abs(np.linalg.norm(nums[:3]) - 1.0) <= 1e-11 and abs(nums[0] - 1.0) > 1e-6     # it precessed
    # asd_sd_run: 10 samples, the last one is the sum of the final moments (= asd_measure's, printed before it)
    assert 'samples = 10' in out
    assert np.abs(np.array(nums[-3:]) - np.array(nums[3:6])).max() <= 1e-12

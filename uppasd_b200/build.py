"""Builds libuppasd_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'asd_engine.cu')
DEPS = [os.path.join(HERE, 'csrc', f) for f in os.listdir(os.path.join(HERE, 'csrc'))] + \
       [os.path.join(HERE, '..', 'include', 'uppasd_b200.h')]
OUT = os.environ.get('ASD_LIB_OUT') or os.path.join(HERE, 'libuppasd_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '--use_fast_math=false',
         '-Xcompiler', '-fPIC', '-shared', '-ccbin', 'g++']   # no -split-compile: its parallel back end is not deterministic (register allocation differs between builds; the run kernel then times 0.47 or 0.55 ms per step)


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in DEPS):
        return OUT
    flags = [f for f in FLAGS if f != '--use_fast_math=false']
    for k in ('ASD_MINB', 'ASD_CHUNK', 'ASD_MINB_STAGED', 'ASD_NPF', 'ASD_MC_MINB', 'ASD_MC_CHUNK', 'ASD_RUN_UNROLL', 'ASD_MC_PROF', 'ASD_INT_UNROLL', 'ASD_NO_TFIELD', 'ASD_ABL', 'ASD_NO_OWNPF', 'ASD_WALK_U', 'ASD_WALK_X'):
        if os.environ.get(k):
            flags.append('-D%s=%s' % (k, os.environ[k]))
    cmd = [NVCC] + flags + (['-Xptxas', '-v'] if verbose else []) + ['-o', OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed')
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == '__main__':
    build(force=True, verbose='-v' in sys.argv)
    print('built', OUT)

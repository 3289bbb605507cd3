"""Builds the CUDA library (csrc/astro_b200.cu -> libastro_b200.so) in-tree for sm_100a.

`python -m astro_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.
-fmad=false: the reference's float64 operations are one rounding each; intended FMAs are
written explicitly (__fmaf_rn) in the fp32 paths.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'astro_b200.cu')
DEPS = [SRC, os.path.join(HERE, 'csrc', 'astro_device.cuh'), os.path.join(HERE, 'csrc', 'tick_f32.cuh'),
        os.path.join(os.path.dirname(HERE), 'include', 'astro_b200.h')]
LIB = os.path.join(HERE, 'libastro_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-fmad=false',
              '-std=c++17', '-shared', '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=/path/to/nvcc)')


def is_stale():
    return not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS)


def build_native(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + ['-o', LIB, SRC]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n%s\n%s' % (' '.join(cmd), r.stdout))
    if verbose:
        print(r.stdout)
    return LIB


if __name__ == '__main__':
    build_native(force='--force' in sys.argv, verbose=True)
    print(LIB)

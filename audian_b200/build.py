"""Build libaudian_b200.so in-tree with nvcc for sm_100a.

    python -m audian_b200.build [--force] [--verbose]

The shared library is plain CUDA C++ behind a C ABI (include/audian_b200.h);
it does not link against torch.  nvcc cross-compiles without a GPU.
"""

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SOURCES = ['api.cu', 'misc.cu', 'minmax.cu', 'sosfilt.cu', 'zerophase.cu', 'sosfwd.cu', 'ingest.cu', 'spectrogram.cu']
HEADERS = [os.path.join(CSRC, 'common.cuh'), os.path.join(CSRC, 'sos_common.cuh'),
           os.path.join(os.path.dirname(HERE), 'include', 'audian_b200.h')]
LIB = os.path.join(HERE, 'libaudian_b200.so')
OBJDIR = os.path.join(HERE, 'build')

def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isfile(cand) or cand == 'nvcc'):
            return cand
    return 'nvcc'


def _stale(target, deps):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    flags = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
             '-std=c++17', '-Xcompiler', '-fPIC']
    if verbose:
        flags += ['-Xptxas', '-v']
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJDIR, s.replace('.cu', '.o'))
        objs.append(obj)
        if force or _stale(obj, [src] + HEADERS):
            cmd = [nvcc] + flags + ['-c', src, '-o', obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                              stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f'nvcc failed on {s}\n')
    if failed:
        raise RuntimeError('building libaudian_b200.so failed')
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError('linking libaudian_b200.so failed')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))

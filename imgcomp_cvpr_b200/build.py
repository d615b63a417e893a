"""Builds libimgcomp_b200.so in-tree with nvcc for sm_100a (no GPU needed).

    python -m imgcomp_cvpr_b200.build [--force]

The shared object is linked with the static CUDA runtime and without libcuda, so
it loads (and exports every symbol of include/imgcomp_b200.h) on a CPU-only box;
driver entry points needed for TMA descriptors are resolved at run time.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(PKG, 'build')
LIB = os.path.join(PKG, 'libimgcomp_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
CFLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include'), '-I', CSRC,
          '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(('.cu', '.cpp')))


def _digest(path):
    h = hashlib.sha1()
    for dep in [path] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))) + \
            [os.path.join(ROOT, 'include', 'imgcomp_b200.h')]:
        with open(dep, 'rb') as f:
            h.update(f.read())
    h.update(' '.join(ARCH + CFLAGS).encode())
    return h.hexdigest()


def _compile(src):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src + '.o')
    stamp = obj + '.sha1'
    dig = _digest(path)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False, ''
    cmd = [NVCC] + ARCH + CFLAGS + ['-c', path, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    with open(stamp, 'w') as f:
        f.write(dig)
    return obj, True, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(_compile, _sources()))
    objs = [r[0] for r in results]
    if verbose:
        for r in results:
            if r[1]:
                print(r[2])
    if any(r[1] for r in results) or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ['-shared', '-Xcompiler', '-fPIC', '-o', LIB] + objs + ['-cudart', 'static']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))

"""Builds librcgan_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the
library is a plain C-ABI shared object loaded with ctypes, see _C.py)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'librcgan_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
         '--expt-relaxed-constexpr', '-Xptxas', '-v']


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def sources_present():
    return os.path.isdir(CSRC) and any(f.endswith('.cu') for f in os.listdir(CSRC))


def source_hash():
    """sha256 over every source the library is built from (content, not mtime: the snapshot that travels to the GPU box
    does not have to preserve timestamps)"""
    import hashlib
    h = hashlib.sha256()
    deps = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h')))
    deps.append(os.path.join(HERE, '..', 'include', 'rcgan_b200.h'))
    for d in deps:
        h.update(os.path.basename(d).encode() + b'\0')
        h.update(open(d, 'rb').read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


STAMP = OUT + '.stamp'


def needs_build():
    if not os.path.exists(OUT) or not os.path.exists(STAMP):
        return True
    return open(STAMP).read().strip() != source_hash()


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    log = []

    def cc(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        r = subprocess.run([NVCC] + FLAGS + ['-c', src, '-o', obj], capture_output=True, text=True)
        log.append((src, r.stdout + r.stderr))
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (src, r.stdout + r.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(cc, sources()))
    r = subprocess.run([NVCC, '-shared', '-o', OUT] + objs + ['-lcudart'], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    with open(STAMP, 'w') as f:
        f.write(source_hash() + '\n')
    with open(os.path.join(objdir, 'ptxas.log'), 'w') as f:
        for src, out in log:
            f.write('==== %s\n%s\n' % (src, out))
    if verbose:
        print('built', OUT)
    return OUT


if __name__ == '__main__':
    build_library(force='--force' in sys.argv, verbose=True)

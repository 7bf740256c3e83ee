"""Builds gator_b200/_C/libgator_b200.so from csrc/*.cu with nvcc for sm_100a (no torch headers: the
boundary is a plain C ABI, see include/gator_b200.h).  Incremental: one object per source file."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT_DIR = os.path.join(HERE, '_C')
LIB_PATH = os.path.join(OUT_DIR, 'libgator_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; cannot build libgator_b200.so')


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, 'rb') as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(os.path.join(OUT_DIR, 'obj'), exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(HERE, '..', 'include', 'gator_b200.h'))
    hdr_digest = _digest(hdrs)
    nvcc = _nvcc()
    objs, rebuilt = [], False
    procs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OUT_DIR, 'obj', s[:-3] + '.o')
        stamp = obj + '.stamp'
        want = _digest([src]) + hdr_digest
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
            continue
        cmd = [nvcc] + NVCC_FLAGS + ['-c', src, '-o', obj]
        if verbose:
            print(' '.join(cmd), file=sys.stderr)
        procs.append((subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT), s, stamp, want))
    for p, s, stamp, want in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {s}:\n{out.decode()}')
        with open(stamp, 'w') as f:
            f.write(want)
        rebuilt = True
    if rebuilt or force or not os.path.exists(LIB_PATH):
        cmd = [nvcc, '-shared', '-o', LIB_PATH] + objs
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n' + r.stdout.decode())
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))

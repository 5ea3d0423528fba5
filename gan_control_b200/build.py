"""Build libb200gan.so (sm_100a only) in-tree with nvcc.

    python -m gan_control_b200.build [--force] [--verbose]

The shared library links cudart statically and resolves driver entry points at run
time, so it loads (without running) on a machine that has no GPU driver.
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libb200gan.so')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
NVCC_FLAGS = ['-O3', '-std=c++17', '-lineinfo', '--use_fast_math', '-Xcompiler', '-fPIC',
              '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']
# exported symbols are marked by extern "C" in the sources; keep them visible
NVCC_FLAGS[NVCC_FLAGS.index('-fvisibility=hidden')] = '-fvisibility=default'
# IEEE arithmetic (no approximate division / sqrt, no flush-to-zero): the optimiser must match torch.optim.Adam
NO_FAST_MATH = {'optim.cu'}


def nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; libb200gan.so cannot be built')
    return exe


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, 'rb') as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(' '.join(ARCH + NVCC_FLAGS + sorted(NO_FAST_MATH)).encode())
    return h.hexdigest()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def headers():
    inc = os.path.join(os.path.dirname(HERE), 'include', 'b200gan.h')
    return sorted([os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))] + [inc])


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    hdr_digest = _digest(headers())
    objs, jobs = [], []
    for src in sources():
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + '.o')
        stamp = obj + '.sha'
        want = _digest([src]) + hdr_digest
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
            continue
        jobs.append((src, obj, stamp, want))

    def compile_one(job):
        src, obj, stamp, want = job
        flags = [f for f in NVCC_FLAGS if not (f == '--use_fast_math' and os.path.basename(src) in NO_FAST_MATH)]
        cmd = [nvcc()] + ARCH + flags + ['-Xptxas', '-v'] + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        with open(stamp, 'w') as f:
            f.write(want)
        with open(obj + '.ptxas.log', 'w') as f:
            f.write(r.stderr)
        return src, r.stderr

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, log in ex.map(compile_one, jobs):
                if verbose:
                    print(f'[nvcc] {os.path.basename(src)}')
                    print(log)
    if jobs or not os.path.exists(LIB) or force:
        cmd = [nvcc()] + ARCH + ['-shared', '-o', LIB] + objs + ['-cudart', 'static', '-Xcompiler', '-fPIC']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return LIB


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(path)

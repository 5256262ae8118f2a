"""Build libbsig_b200.so (sm_100a) in-tree with nvcc.

    python -m bayes_sim_ig_b200.build [--force]

Every ``csrc/*.cu`` is compiled with
``-gencode arch=compute_100a,code=sm_100a -lineinfo -O3`` (nvcc cross-compiles
without a GPU) and linked into ``bayes_sim_ig_b200/libbsig_b200.so``; objects
live in ``bayes_sim_ig_b200/_build/``.  Files are rebuilt only when a source or
header is newer than its object.
"""
import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, 'csrc')
BUILD_DIR = os.path.join(PKG_DIR, '_build')
LIB_PATH = os.path.join(PKG_DIR, 'libbsig_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC',
    '-I', os.path.join(ROOT, 'include'), '-I', CSRC,
] + os.environ.get('BSIG_NVCC_EXTRA', '').split()
# (BSIG_NVCC_EXTRA: instrumented builds for profiles/*_marks.py; the staleness check looks at
# mtimes only, so touch the source -- and again before the clean rebuild)


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: cannot build libbsig_b200.so')
    return exe


def _newest_header():
    hdrs = glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(ROOT, 'include', '*.h'))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(nvcc, src, obj, verbose):
    cmd = [nvcc] + NVCC_FLAGS + ['-c', src, '-o', obj]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, res.stdout, res.stderr))
    return res.stderr


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(BUILD_DIR, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    hdr_time = _newest_header()
    jobs, objs = [], []
    for src in sources:
        obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        stale = force or not os.path.exists(obj) or \
            os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time)
        if stale:
            jobs.append((src, obj))
    logs = []
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            futs = [pool.submit(_compile, nvcc, s, o, verbose) for s, o in jobs]
            for f in futs:
                logs.append(f.result())
    if jobs or not os.path.exists(LIB_PATH):
        cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a',
               '-Xcompiler', '-fPIC', '-o', LIB_PATH] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (res.stdout, res.stderr))
    return LIB_PATH, ''.join(logs)


if __name__ == '__main__':
    path, log = build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    if log:
        print(log)
    print(path)

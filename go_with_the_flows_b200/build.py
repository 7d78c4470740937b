"""Build libgwtf.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

One translation unit per kernel family (csrc/gwtf_api.cu + csrc/gwtf_l_*.cu), compiled in parallel into
build/*.o and linked into go_with_the_flows_b200/libgwtf.so.  Only stale objects are rebuilt."""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJDIR = os.path.join(HERE, 'build')
OUT = os.path.join(HERE, 'libgwtf.so')
HEADER = os.path.join(os.path.dirname(HERE), 'include', 'gwtf.h')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
         '-diag-suppress', '177']


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _headers():
    return glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(CSRC, '*.h')) + [HEADER]


def _compile(nvcc, src, obj, verbose):
    cmd = [nvcc] + FLAGS + os.environ.get('GWTF_NVCC_EXTRA', '').split() + (['-Xptxas', '-v'] if verbose else []) + \
        ['-c', src, '-o', obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res.returncode, res.stdout + res.stderr


def build(force=False, verbose=False):
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    os.makedirs(OBJDIR, exist_ok=True)
    hdr_time = max(os.path.getmtime(h) for h in _headers())
    jobs = []
    objs = []
    for src in _sources():
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            jobs.append((src, obj))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
            for src, rc, log in pool.map(lambda j: _compile(nvcc, j[0], j[1], verbose), jobs):
                if rc != 0:
                    sys.stderr.write(log)
                    raise RuntimeError('nvcc failed on %s' % os.path.basename(src))
                if verbose:
                    sys.stderr.write(log)
    if jobs or not os.path.exists(OUT) or any(os.path.getmtime(o) > os.path.getmtime(OUT) for o in objs):
        res = subprocess.run([nvcc, '-shared', '-o', OUT] + objs, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError('linking libgwtf.so failed')
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

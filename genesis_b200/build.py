"""Build libgenesis_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m genesis_b200.build [--force]

Objects go to genesis_b200/csrc/build/, the library to genesis_b200/lib/libgenesis_b200.so (git-ignored,
but shipped to the GPU box by gpurun)."""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, 'build')
LIB_DIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIB_DIR, 'libgenesis_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
    cmd = [NVCC] + FLAGS + ['-I', CSRC, '-c', src, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    return obj


def build(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    hdrs = sorted(glob.glob(os.path.join(CSRC, '*.cuh')))
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    todo = [s for s in srcs
            if force or _stale(os.path.join(OBJ, os.path.basename(s)[:-3] + '.o'), [s] + hdrs)]
    if verbose and todo:
        print('nvcc:', ' '.join(os.path.basename(s) for s in todo))
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        list(ex.map(_compile, todo))
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + '.o') for s in srcs]
    if force or todo or _stale(LIB, objs):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-lcudart', '-lcuda']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


PROBE_SRC = os.path.join(os.path.dirname(HERE), 'tests', 'probe', 'debug_umma.cu')
PROBE_LIB = os.path.join(os.path.dirname(HERE), 'tests', 'probe', 'libgenesis_b200_probe.so')


def build_probe(force=False):
    """TEST INFRASTRUCTURE: the tcgen05 probes (descriptor self-test, instruction-rate probe) as their own library."""
    if force or _stale(PROBE_LIB, [PROBE_SRC] + sorted(glob.glob(os.path.join(CSRC, '*.cuh')))):
        cmd = [NVCC] + FLAGS + ['-I', CSRC, '-shared', PROBE_SRC, '-o', PROBE_LIB, '-lcudart', '-lcuda']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for the probe library:\n%s\n%s' % (r.stdout, r.stderr))
    return PROBE_LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
    print(build_probe(force='--force' in sys.argv))

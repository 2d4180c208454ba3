"""Builds the CUDA library in-tree: hybrid-drt_b200/_lib/libhybdrt_b200.so (sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, '_lib')
LIB_PATH = os.path.join(LIB_DIR, 'libhybdrt_b200.so')
SOURCES = ['capi.cu', 'matrix_kernels.cu', 'chrono_kernels.cu', 'qphb_kernel.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v', '-I', os.path.join(ROOT, 'include'), '-I', CSRC]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, 'include', 'hybdrt_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every .cu for sm_100a and link the shared library.  Returns the library path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append(f'== {src}\n{out}')
        if pr.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
    link = [nvcc, '-shared', '-o', LIB_PATH] + objs + ['-lcudart']
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'link failed:\n{res.stdout}')
    with open(os.path.join(LIB_DIR, 'build.log'), 'w') as fh:
        fh.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose=True))

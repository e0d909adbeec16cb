"""Builds libctgan_sm100.so in-tree (ctgan_b200/libctgan_sm100.so) with nvcc for sm_100a.

Run as `python -m ctgan_b200.build` or through `__graft_entry__.build()`.  The .so is
git-ignored but travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, 'build')
LIB = os.path.join(HERE, 'libctgan_sm100.so')
SOURCES = ['api.cu', 'conv_simt.cu', 'conv_thin.cu', 'conv_head.cu', 'conv_tc.cu', 'conv_wgrad_multi.cu', 'conv_splitk.cu', 'conv_tf32.cu', 'conv_s2d.cu', 'elementwise.cu', 'norm.cu', 'layernorm.cu', 'loss.cu', 'optim.cu', 'peer.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, 'rb') as f:
            h.update(f.read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build_extension(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    headers.append(os.path.join(HERE, '..', 'include', 'ctgan_sm100.h'))
    objs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src.replace('.cu', '.o'))
        stamp = obj + '.sha'
        dig = _digest([path] + headers)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', path, '-o', obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or r.returncode != 0:
            sys.stdout.write(r.stdout)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed on %s' % src)
        with open(stamp, 'w') as f:
            f.write(dig)
    link_stamp = LIB + '.sha'
    dig = _digest(objs)
    if force or not os.path.exists(LIB) or not os.path.exists(link_stamp) or open(link_stamp).read() != dig:
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stdout.write(r.stdout)
            raise RuntimeError('link failed')
        with open(link_stamp, 'w') as f:
            f.write(dig)
    return LIB


if __name__ == '__main__':
    print(build_extension(verbose='-v' in sys.argv, force='-f' in sys.argv))

"""Builds `xtrack_b200/libxtb200.so` (the C-ABI library, include/xtb200.h) with
nvcc for sm_100a.  Run `python -m xtrack_b200.build` (or `__graft_entry__.build()`).

The tracking kernel translation unit is compiled twice: with FMA contraction
(default, `xtb_launch_track_fast`) and without (`-fmad=false`,
`xtb_launch_track_exact`), see csrc/xtb_kernel_inst.cu.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
# Tuning experiments: XTB_LIB_SUFFIX=_x XTB_EXTRA_DEFINES="-DXTB_NPT_THIN=1" builds
# libxtb200_x.so beside the product library; XTB_LIB_SUFFIX alone selects it at load time.
SUFFIX = os.environ.get('XTB_LIB_SUFFIX', '')
EXTRA = os.environ.get('XTB_EXTRA_DEFINES', '').split()
LIB = os.path.join(HERE, f'libxtb200{SUFFIX}.so')
OBJ = os.path.join(HERE, 'csrc', '_build' + SUFFIX)

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
HEAVY = ['-DXTB_WITH_HEAVY']
COMMON = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', *HEAVY, *EXTRA]


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found')
    return exe


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
            if f.endswith(('.cu', '.cuh', '.h'))] + [
        os.path.join(HERE, '..', 'include', 'xtb200.h'), __file__]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    units = [
        ('xtb_kernel_fast.o', 'xtb_kernel_inst.cu', ['-DXTB_EXACT=0', '-fmad=true']),
        ('xtb_kernel_exact.o', 'xtb_kernel_inst.cu', ['-DXTB_EXACT=1', '-fmad=false']),
        ('xtb_api.o', 'xtb_api.cu', []),
    ]
    procs = []
    for obj, src, extra in units:
        cmd = [nvcc, *ARCH, *COMMON, *extra, '-Xptxas', '-v', '-c',
               os.path.join(CSRC, src), '-o', os.path.join(OBJ, obj)]
        if verbose:
            print(' '.join(cmd))
        procs.append((obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                            text=True)))
    log = []
    for obj, pp in procs:
        out, _ = pp.communicate()
        log.append(f'==== {obj}\n{out}')
        if pp.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f'nvcc failed on {obj}')
    with open(os.path.join(OBJ, 'ptxas.log'), 'w') as fid:
        fid.write('\n'.join(log))
    cmd = [nvcc, *ARCH, '-shared', '-o', LIB, *[os.path.join(OBJ, o) for o, _, _ in units],
           '-lcudart']
    if verbose:
        print(' '.join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build(force='-f' in sys.argv, verbose=True))

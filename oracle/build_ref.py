"""ORACLE build recipe (test infrastructure).

Compiles the reference's own physics headers, in place from /root/reference,
through the generated shim into shared libraries under `oracle/_ref/`
(git-ignored; they travel to the GPU box with the gpurun snapshot):

  libxt_ref_serial.so       -DXO_CONTEXT_CPU_SERIAL              bit reference
  libxt_ref_omp.so          -DXO_CONTEXT_CPU_OPENMP -fopenmp     timing baseline
  libxt_ref_noise.so        serial + +-1 ulp noise on libm results  libm-sensitivity yardstick
  libxt_ref_synrad*.so      the same three without -DXTRACK_MULTIPOLE_NO_SYNRAD: synchrotron
                            radiation compiled in (mean / quantum models)

Flags mirror xobjects' CPU context as far as it is known (`-O3`, no
`-march=native`, no `-ffast-math`): baseline x86-64 has no FMA, so the
arithmetic is plain IEEE double without contraction (`-ffp-contract=off` makes
that explicit).  No reference source is copied: the only inputs from the
reference are `-I/root/reference` include paths.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('XTB_REFERENCE_ROOT', '/root/reference')
OUT = os.path.join(HERE, '_ref')

FROZEN_VARS = ('zeta', 'delta', 'ptau', 'rpp', 'rvv', 's')   # set by build()

VARIANTS = {
    'serial': ['-DXO_CONTEXT_CPU_SERIAL', '-DXTRACK_MULTIPOLE_NO_SYNRAD'],
    'omp': ['-DXO_CONTEXT_CPU_OPENMP', '-fopenmp', '-DXTRACK_MULTIPOLE_NO_SYNRAD'],
    # the clean serial build with +-1 ulp noise on every transcendental libm result
    # (shim/xobjects/headers/ulp_noise.h): measures the reference's sensitivity to its libm
    'noise': ['-DXO_CONTEXT_CPU_SERIAL', '-DXTRACK_MULTIPOLE_NO_SYNRAD', '-DXTB_ORACLE_ULP_NOISE'],
    # radiation compiled in (line.config XTRACK_MULTIPOLE_NO_SYNRAD = False after
    # configure_radiation, line.py:4744-4837): mean / quantum models, per-particle RNG
    'synrad': ['-DXO_CONTEXT_CPU_SERIAL'],
    'synrad_omp': ['-DXO_CONTEXT_CPU_OPENMP', '-fopenmp'],
    'synrad_noise': ['-DXO_CONTEXT_CPU_SERIAL', '-DXTB_ORACLE_ULP_NOISE'],
    # OpenMP builds of the noise variants (the noise is a hash of the argument: stateless),
    # so that the GPU tests do not spend the GPU box's time on a serial CPU run
    'noise_omp': ['-DXO_CONTEXT_CPU_OPENMP', '-fopenmp', '-DXTRACK_MULTIPOLE_NO_SYNRAD',
                  '-DXTB_ORACLE_ULP_NOISE'],
    'synrad_noise_omp': ['-DXO_CONTEXT_CPU_OPENMP', '-fopenmp', '-DXTB_ORACLE_ULP_NOISE'],
}


def available():
    return os.path.isdir(os.path.join(REF, 'xtrack', 'beam_elements', 'elements_src'))


def lib_path(variant):
    return os.path.join(OUT, f'libxt_ref_{variant}.so')


def build(variants=None, force=False, verbose=False):
    if not available():
        raise RuntimeError(f'reference tree not found at {REF}')
    os.makedirs(os.path.join(OUT, 'gen'), exist_ok=True)
    sys.path.insert(0, HERE)
    import gen_shim
    src = gen_shim.generate()
    gen_h = os.path.join(OUT, 'gen', 'xt_generated.h')
    if not os.path.exists(gen_h) or open(gen_h).read() != src:
        with open(gen_h, 'w') as fid:
            fid.write(src)
    built = []
    for vv in (variants or VARIANTS):
        out = lib_path(vv)
        deps = [gen_h, os.path.join(HERE, 'track_line.c'), __file__,
                os.path.join(HERE, 'shim', 'xobjects', 'headers', 'common.h'),
                os.path.join(HERE, 'shim', 'xobjects', 'headers', 'ulp_noise.h'),
                os.path.join(HERE, 'shim', 'xtrack', 'headers', 'synrad_total_energy_tables.h')]
        if (not force and os.path.exists(out)
                and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps)):
            continue
        cmd = ['gcc', '-std=gnu99', '-O3', '-ffp-contract=off', '-fPIC', '-shared',
               '-Wno-unused-function', '-Wno-unused-variable',
               *VARIANTS[vv],
               '-I', os.path.join(OUT, 'gen'), '-I', os.path.join(HERE, 'shim'),
               '-I', REF,
               os.path.join(HERE, 'track_line.c'), '-o', out, '-lm']
        if verbose:
            print(' '.join(cmd))
        subprocess.run(cmd, check=True)
        built.append(out)
    return built


if __name__ == '__main__':
    print(build(force='-f' in sys.argv, verbose=True))

"""Builds lib/libdanet_sm100.so (the C-ABI of include/danet.h) with nvcc for sm_100a."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, 'lib', 'libdanet_sm100.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']
# experiment builds (e.g. DANET_NVCC_EXTRA=-DDANET_LSTM_PRE_DEPTH=3): part of the recorded hash, so a change rebuilds
FLAGS += [f for f in os.environ.get('DANET_NVCC_EXTRA', '').split() if f]


def sources():
    return sorted(glob.glob(os.path.join(HERE, 'csrc', '*.cu')))


INFO = os.path.join(HERE, 'lib', 'build_info.json')


def source_hash():
    """sha256 over every source the library is built from (+ the compiler flags): what `build()` records next to the
    .so, so a shipped binary can be told apart from one that no longer matches the tree"""
    import hashlib
    h = hashlib.sha256(' '.join(FLAGS).encode())
    deps = sources() + sorted(glob.glob(os.path.join(HERE, 'csrc', '*.cuh'))) + \
        [os.path.join(os.path.dirname(HERE), 'include', 'danet.h')]
    for d in deps:
        h.update(os.path.basename(d).encode())
        h.update(open(d, 'rb').read())
    return h.hexdigest()


def built_hash():
    import json
    try:
        return json.load(open(INFO)).get('source_sha256')
    except Exception:
        return None


def stale():
    """True when there is no library or it was built from other sources than the tree holds now (content hash, not
    modification times: a snapshot copied to the GPU box has fresh mtimes on every file)"""
    return not os.path.exists(LIB) or built_hash() != source_hash()


def build(force=False, verbose=False):
    """Compile every .cu (one nvcc process per file, in parallel) and link the shared library."""
    if not force and not stale():
        return LIB
    os.makedirs(os.path.join(HERE, 'lib'), exist_ok=True)
    objdir = os.path.join(HERE, 'lib', 'obj')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    objs = []
    for src, obj, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, out))
        if verbose:
            sys.stderr.write(out)
        objs.append(obj)
    cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs + ['-lcuda']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s' % r.stdout.decode())
    import json
    import time
    with open(INFO, 'w') as f:
        json.dump({'source_sha256': source_hash(), 'flags': FLAGS, 'nvcc': NVCC, 'built_at': time.strftime('%Y-%m-%dT%H:%M:%SZ', time.gmtime()),
                   'sources': [os.path.basename(x) for x in sources()]}, f, indent=1)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

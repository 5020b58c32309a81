"""TEST / BENCH INFRASTRUCTURE ONLY -- installs the UNMODIFIED reference (pyRDDLGym, pure Python) from
/root/reference into baseline/_ref with pip, so that it travels to the GPU box with the tree
(baseline/_ref is git-ignored, not gpurun-ignored). There the parity tests step the CUDA kernels
against the live reference RDDLSimulator and ``bench.py --impl reference`` times it on the host cores.

    python oracle/install_reference.py        (also run by __graft_entry__.build())

/root/reference is read-only and the build writes into the source tree, so the install runs from a
copy under /tmp; --no-deps because gymnasium / ply / pygame are not in the offline wheelhouse (the
simulator and compiler need only NumPy; oracle/ref_import.py stubs the absent modules)."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, 'baseline', '_ref')
SOURCE = os.environ.get('RDDL_REFERENCE_SOURCE', '/root/reference')


def install(force=False):
    """Returns the install directory, or None when the reference source is not present (GPU box)."""
    if os.path.isdir(os.path.join(TARGET, 'pyRDDLGym')) and not force:
        return TARGET
    if not os.path.isdir(os.path.join(SOURCE, 'pyRDDLGym')):
        return None
    tmp = tempfile.mkdtemp(prefix='refsrc_')
    try:
        src = os.path.join(tmp, 'reference')
        shutil.copytree(SOURCE, src, symlinks=True)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        os.makedirs(TARGET, exist_ok=True)
        cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--find-links',
               '/opt/wheelhouse', '--no-deps', '--quiet', '--target', TARGET, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError('pip install of the reference failed')
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return TARGET


if __name__ == '__main__':
    print(install(force='--force' in sys.argv))

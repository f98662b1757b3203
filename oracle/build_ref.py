"""Recipe for oracle/_ref: the UNMODIFIED reference, importable on the GPU box.

The reference (fgnt/pb_bss) is pure Python, so "building" it is a verbatim copy of the modules the hot path needs
from the read-only checkout (/root/reference, build container only) into the git-ignored directory oracle/_ref/
(which is NOT gpurun-ignored, so it travels to the GPU box like the built libpbb.so).  Nothing is edited; the import
shims of oracle/ref_shim.py (stub top-level package, cached_property) are applied at import time.  bench.py's
reference arm then times the real reference (`cpu_baseline.kind: "reference"`); without oracle/_ref it falls back
to the NumPy port (`"port"`).  No file of the reference is ever committed (.gitignore: oracle/_ref/).

    python -m oracle.build_ref          # called by __graft_entry__.build() when /root/reference exists
"""
import os
import shutil

SRC = os.environ.get('PB_BSS_REFERENCE', '/root/reference')
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
# modules of the hot path and what they import
PARTS = ['distribution', 'extraction', 'math', 'permutation_alignment.py', 'utils.py', 'testing', 'initializer']


def build():
    if not os.path.isdir(os.path.join(SRC, 'pb_bss')):
        return False
    out = os.path.join(DST, 'pb_bss')
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(out)
    for p in PARTS:
        s = os.path.join(SRC, 'pb_bss', p)
        if os.path.isdir(s):
            shutil.copytree(s, os.path.join(out, p), ignore=shutil.ignore_patterns('__pycache__', '*.pyc', '*.so', '*.c'))
        elif os.path.isfile(s):
            shutil.copy(s, os.path.join(out, p))
    with open(os.path.join(DST, 'README'), 'w') as f:
        f.write('verbatim copy of the hot-path modules of the reference (oracle/build_ref.py); not part of the repository\n')
    return True


if __name__ == '__main__':
    print('oracle/_ref built' if build() else f'no reference checkout at {SRC}: oracle/_ref not built')

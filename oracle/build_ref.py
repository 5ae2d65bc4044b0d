"""Build ``oracle/_ref/`` from the reference where it lies under /root/reference -- TEST / BENCH INFRASTRUCTURE.

The reference is pure Python (no native sources), so "compiling the path from its own source files"
means byte-compiling it: every ``pretorched/**/*.py`` of the UNMODIFIED tree is compiled with ``py_compile`` to a
sourceless ``.pyc`` at the same relative path under ``oracle/_ref/pretorched/`` (legacy layout, importable without the
``.py``).  No reference source text enters the repository: ``oracle/_ref/`` is git-ignored (it ships to the GPU box with the
snapshot like any other built artefact), and the GPU box -- same image, same interpreter -- imports the compiled package, so
``bench.py --impl reference`` and the ``cpu_baseline`` leg time the reference's own ``forward`` there (``kind: "reference"``).

    python -m oracle.build_ref          # needs /root/reference; idempotent
"""
import os
import py_compile
import shutil
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("B2_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")


def build(force=False):
    src_pkg = os.path.join(SRC, "pretorched")
    if not os.path.isdir(src_pkg):
        return None                                  # GPU box: use whatever the snapshot brought along
    stamp = os.path.join(DST, "BUILT_FROM")
    tag = "%s python %d.%d" % (SRC, sys.version_info[0], sys.version_info[1])
    if not force and os.path.exists(stamp) and open(stamp).read().strip() == tag:
        return DST
    shutil.rmtree(DST, ignore_errors=True)
    n = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")              # dpn.py:255-262 has invalid escape sequences (SyntaxWarning)
        for dirpath, _, files in os.walk(src_pkg):
            rel = os.path.relpath(dirpath, SRC)
            for f in files:
                if f.endswith(".py"):
                    out = os.path.join(DST, rel, f + "c")
                    os.makedirs(os.path.dirname(out), exist_ok=True)
                    py_compile.compile(os.path.join(dirpath, f), cfile=out, dfile=os.path.join(rel, f), doraise=True)
                    n += 1
    with open(stamp, "w") as fh:
        fh.write(tag + "\n")
    return DST


if __name__ == "__main__":
    path = build(force="--force" in sys.argv)
    print(path if path else "no reference tree at %s: nothing built" % SRC)

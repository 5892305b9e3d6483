#!/usr/bin/env python
"""Compile the reference's own Cython hot path into ``oracle/_ref`` (TEST INFRASTRUCTURE ONLY).

What it does
------------
* reads ``holodeck/cyutils.pyx``, ``holodeck/cyutils.pxd`` and ``holodeck/sams/sam_cyutils.pyx``
  where they lie under ``/root/reference`` (nothing is copied into the repository),
* stages them in a temporary directory, applying two mechanical edits to the *staged copy*:
    1. ``long(normal_threshold)`` -> ``int(normal_threshold)`` at ``cyutils.pyx:855``
       (Python-2 builtin; the reference pins Cython<3, this image has Cython 3.3);
    2. ``PCG64()`` -> ``PCG64(_oracle_seed())`` at the seven RNG construction sites
       (``cyutils.pyx:674,875,979,1120,1315,1472,1698``) where ``_oracle_seed()`` returns the
       module global ``ORACLE_SEED`` (default ``None`` == the reference's unseeded behaviour).
       This makes realizations reproducible so supplied-count parity can be bit-exact;
* cythonizes + compiles with the reference's ``setup.py:32-66`` flags (npyrandom, npymath),
* writes ONLY the two ``.so`` files plus the stub package from ``oracle/ref_stub`` into
  ``oracle/_ref/holodeck`` (git-ignored, travels to the GPU box with the snapshot).

Run:  ``python oracle/build_ref.py``  (needs ``/root/reference``; a no-op on the GPU box).
"""
import os
import shutil
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path(os.environ.get("HOLO_REFERENCE", "/root/reference"))
OUT = HERE / "_ref"

SEED_HELPER = '''
ORACLE_SEED = None

def _oracle_seed():
    return ORACLE_SEED

'''


def stage_sources(tmp):
    pkg = tmp / "holodeck"
    (pkg / "sams").mkdir(parents=True)
    src = (REF / "holodeck" / "cyutils.pyx").read_text()
    n_long = src.count("long(normal_threshold)")
    assert n_long == 1, n_long
    src = src.replace("long(normal_threshold)", "int(normal_threshold)")
    n_rng = src.count("PCG64()")
    assert n_rng == 7, n_rng
    src = src.replace("PCG64()", "PCG64(_oracle_seed())")
    marker = "# ---- Define Parameters"
    assert marker in src
    src = src.replace(marker, SEED_HELPER + marker, 1)
    (pkg / "cyutils.pyx").write_text(src)
    shutil.copy(REF / "holodeck" / "cyutils.pxd", pkg / "cyutils.pxd")
    shutil.copy(REF / "holodeck" / "sams" / "sam_cyutils.pyx", pkg / "sams" / "sam_cyutils.pyx")
    # package markers so that `from holodeck.cyutils cimport ...` resolves during cythonize
    (pkg / "__init__.py").write_text("")
    (pkg / "sams" / "__init__.py").write_text("")
    return pkg


def build():
    if not REF.exists():
        print(f"[oracle] {REF} not present: keeping prebuilt oracle/_ref as is")
        return False
    import numpy as np
    from setuptools import Extension
    from setuptools.dist import Distribution
    from Cython.Build import cythonize

    tmp = Path(tempfile.mkdtemp(prefix="holo_ref_build_"))
    try:
        stage_sources(tmp)
        cwd = os.getcwd()
        os.chdir(tmp)
        inc = np.get_include()
        libdirs = [os.path.abspath(os.path.join(inc, "..", "..", "random", "lib")),
                   os.path.abspath(os.path.join(inc, "..", "lib"))]
        common = dict(include_dirs=[inc], library_dirs=libdirs, libraries=["npyrandom", "npymath"],
                      define_macros=[("NPY_NO_DEPRECATED_API", 0)],
                      extra_compile_args=["-O2", "-Wno-unused-function", "-w"])
        exts = [
            Extension("holodeck.cyutils", sources=[os.path.join("holodeck", "cyutils.pyx")], **common),
            Extension("holodeck.sams.sam_cyutils",
                      sources=[os.path.join("holodeck", "sams", "sam_cyutils.pyx")], **common),
        ]
        exts = cythonize(exts, compiler_directives={"language_level": "3"}, quiet=True)
        dist = Distribution({"name": "holodeck_ref_oracle", "ext_modules": exts})
        cmd = dist.get_command_obj("build_ext")
        cmd.inplace = True
        cmd.ensure_finalized()
        cmd.run()
        os.chdir(cwd)

        if OUT.exists():
            shutil.rmtree(OUT)
        shutil.copytree(HERE / "ref_stub", OUT)
        n = 0
        for so in (tmp / "holodeck").rglob("*.so"):
            rel = so.relative_to(tmp)
            shutil.copy(so, OUT / rel)
            n += 1
        assert n == 2, n
        print(f"[oracle] built reference Cython -> {OUT}")
        return True
    finally:
        os.chdir(HERE)
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    ok = build()
    sys.exit(0 if ok or OUT.exists() else 1)

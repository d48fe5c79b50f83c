"""Build recipe for oracle/_ref: the reference's ONLY native component, compiled unchanged.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product path
(rvgp_b200/); see oracle/README.md.

What it does
------------
Cythonises ``/root/reference/RVGP/lib/ptu_dijkstra.pyx`` *where it lies* (the same
recipe as the reference's ``setup.py:25-29``: one Extension named ``ptu_dijkstra`` with
numpy's include dir) and writes ONLY the compiled extension module into ``oracle/_ref/``.
The generated C file is a build intermediate kept in a temp dir and deleted.  No reference
source is copied into the repository.

With ``--instrumented`` it also builds ``ptu_dijkstra_instr`` from a temp copy of the .pyx
patched (in /tmp, 3 inserted lines) to export the popped geodesic-neighbourhood index
sequence of every source right before ``pyx:396``; that build is used once by
``tests/golden/make_golden.py`` to pin oracle/geodesic.c's heap emulation by *sequence*
equality.  It is never shipped.

The reference needs int32 CSR indices (``pyx:26-30``) while networkx 3.6 / SciPy 1.18 hand
back int64; ``oracle/ref_harness.py`` handles that by wrapping ``networkx.adjacency_matrix``
(no source edit).
"""
import os
import subprocess
import sys
import sysconfig
import tempfile
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/RVGP/lib/ptu_dijkstra.pyx"
OUT = os.path.join(HERE, "_ref")


def _compile(pyx_path, modname, out_dir):
    import numpy
    tmp = tempfile.mkdtemp(prefix="rvgp_ref_build_")
    try:
        c_file = os.path.join(tmp, modname + ".c")
        subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx_path, "-o", c_file,
                               "--module-name", modname],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ext = sysconfig.get_config_var("EXT_SUFFIX")
        so = os.path.join(out_dir, modname + ext)
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-w", "-fno-strict-aliasing",
               "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include(),
               c_file, "-o", so, "-lm"]
        subprocess.check_call(cmd)
        return so
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def build(instrumented=False):
    """Compile the reference extension; returns path(s) or None when /root/reference is absent."""
    if not os.path.exists(REF_PYX):
        return None
    os.makedirs(OUT, exist_ok=True)
    so = _compile(REF_PYX, "ptu_dijkstra", OUT)
    if instrumented:
        tmp = tempfile.mkdtemp(prefix="rvgp_ref_instr_")
        try:
            src = open(REF_PYX).read()
            # (1) a module-level sink the harness can read back
            src = src.replace("ITYPE = np.int32\n", "ITYPE = np.int32\nGEO_SEQ = []\n", 1)
            # (2) record the popped sequence of every source right before the centring loop
            marker = "        # construct and center geodesic neighborhood from indices\n"
            assert marker in src
            src = src.replace(marker, "        GEO_SEQ.append(geoNbh_indices.copy())\n" + marker, 1)
            p = os.path.join(tmp, "ptu_dijkstra_instr.pyx")
            open(p, "w").write(src)
            _compile(p, "ptu_dijkstra_instr", OUT)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    return so


if __name__ == "__main__":
    r = build(instrumented="--instrumented" in sys.argv)
    print("built" if r else "reference absent: nothing built", r or "")

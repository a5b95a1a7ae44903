"""Compile the UNMODIFIED reference Cython hot-path modules into oracle/_ref/.

Test infrastructure only.  The sources are read where they lie under /root/reference
(read-only); only generated C files and shared objects are written, and only into
oracle/_ref/ (git-ignored, but shipped to the GPU box with the snapshot).  The reference's
own setup.py cannot be used (distutils is gone in Python 3.12 and it passes no NumPy
include directory), so this script restates its intent: cython (language_level=2, the
reference is Python-2 era) followed by the gcc line quoted at the top of each .pyx
(e.g. template_speech_rec/code_parts.pyx:2-3: gcc -shared -pthread -fPIC -fwrapv -O2
-Wall -fno-strict-aliasing).

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
Returns exit code 0 and prints "skipped" when the reference tree is absent (GPU box).
"""
import argparse
import os
import subprocess
import sys
import sysconfig

MODULES = ("code_parts", "spread_waliji_patches", "compute_likelihood_linear_filter", "cluster_times")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def build(reference="/root/reference", force=False, verbose=True):
    src_dir = os.path.join(reference, "template_speech_rec")
    if not os.path.isdir(src_dir):
        if verbose:
            print("oracle/_ref: skipped (no reference tree at %s)" % reference)
        return False
    import numpy as np
    os.makedirs(OUT, exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    inc = [sysconfig.get_paths()["include"], np.get_include()]
    for mod in MODULES:
        pyx = os.path.join(src_dir, mod + ".pyx")
        c_file = os.path.join(OUT, mod + ".c")
        so = os.path.join(OUT, mod + ext)
        if not force and os.path.exists(so) and os.path.getmtime(so) >= os.path.getmtime(pyx):
            continue
        subprocess.check_call([sys.executable, "-m", "cython", "-2", "-o", c_file, pyx])
        cmd = ["gcc", "-shared", "-pthread", "-fPIC", "-fwrapv", "-O2", "-Wall", "-Wno-unused-function",
               "-Wno-cpp", "-fno-strict-aliasing", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"]
        cmd += ["-I" + i for i in inc] + ["-o", so, c_file]
        subprocess.check_call(cmd)
        if verbose:
            print("oracle/_ref: built", os.path.basename(so))
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    build(a.reference, a.force)

"""Build the reference's own MeshUDF marching cubes (the only native component of LIA-DiTella/DiffUDF)
as the topology oracle of north_star / SURVEY 8d config 3 ("marching-cubes vertex and face topology is identical on the
same field").  TEST INFRASTRUCTURE ONLY: nothing under diffudf_b200/ imports it.

    python oracle/build_ref_mc.py            # needs /root/reference (build container); outputs only under oracle/_ref/

The three modules are compiled from the sources WHERE THEY LIE (nothing is copied into the repo):
    src/marching_cubes/_marching_cubes_lewiner_cy.pyx   (Cython -> C++: marching_cubes_udf, LutProvider)
    src/marching_cubes/_marching_cubes_lewiner_luts.py  (lookup tables)   } pure-Python modules, compiled by Cython too so
    src/marching_cubes/_marching_cubes_lewiner.py       (udf_mc_lewiner)  } that they travel as binaries, not as source
The reference's setup.py is not used (its CFLAGS hack is ignored for C++ by current setuptools); this is the short recipe:
cython -> g++ -shared, Python and numpy include directories passed explicitly.  oracle/_ref/ is git-ignored and travels to
the GPU box with the snapshot like the package's own .so."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "/root/reference/src/marching_cubes"
MODULES = [("_marching_cubes_lewiner_cy", "pyx"), ("_marching_cubes_lewiner_luts", "py"), ("_marching_cubes_lewiner", "py")]


def built():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(OUT, name + suffix)) for name, _ in MODULES)


def build(force=False):
    """Returns True if the oracle modules exist afterwards."""
    if built() and not force:
        return True
    if not os.path.isdir(SRC):
        return False
    import numpy as np
    os.makedirs(os.path.join(OUT, "build"), exist_ok=True)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + np.get_include()]
    for name, ext in MODULES:
        cpp = os.path.join(OUT, "build", name + ".cpp")
        subprocess.run([sys.executable, "-m", "cython", "-3", "--cplus", "-o", cpp, os.path.join(SRC, f"{name}.{ext}")], check=True)
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-w", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION", *inc, cpp, "-o",
                        os.path.join(OUT, name + suffix)], check=True)
    import shutil
    shutil.rmtree(os.path.join(OUT, "build"), ignore_errors=True)       # generated C++ does not travel, the binaries do
    return built()


def load():
    """Imports the built modules (oracle/_ref on sys.path) and returns udf_mc_lewiner, or None if they are not built."""
    if not built():
        return None
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import _marching_cubes_lewiner
    return _marching_cubes_lewiner.udf_mc_lewiner


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref marching cubes:", "built" if ok else "NOT built (no /root/reference)")
    sys.exit(0 if ok else 1)

"""TEST INFRASTRUCTURE.  Builds the reference's UNMODIFIED Cython wrapper (recometrics/wrapper.pyx) against
librecometrics_b200.so -- the drop-in the north-star describes: "a thin C-ABI that the existing Cython ... wrappers call".

    python oracle/build_ref_cython.py [/root/reference]   ->  oracle/_ref/cy_b200/cpp_funs.<abi>.so

wrapper.pyx is cythonized where it lies; its `recometrics_signatures.hpp` resolves to oracle/ref_cython_signatures/, which
includes the reference's own header (declarations only) and then include/recometrics_b200_shim.hpp, whose definitions of
calc_metrics_float/_double, get_has_openmp and the six split_data_* entry points all go to the C-ABI.  No reference source
file is compiled into the module: every native call of the reference's Python package lands in librecometrics_b200.so.
Everything generated goes to oracle/_ref/ (git-ignored, travels to the GPU box); no reference source is copied into the
repository.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def build(reference="/root/reference", verbose=False):
    pyx = os.path.join(reference, "recometrics", "wrapper.pyx")
    src = os.path.join(reference, "src")
    lib = os.path.join(ROOT, "recometrics_b200", "librecometrics_b200.so")
    if not (os.path.exists(pyx) and os.path.exists(lib)):
        return None
    import numpy
    out = os.path.join(HERE, "_ref", "cy_b200")
    tmp = os.path.join(out, "build")
    os.makedirs(tmp, exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    target = os.path.join(out, "cpp_funs" + ext)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    run = lambda cmd: subprocess.run(cmd, check=True, capture_output=not verbose, text=True)
    wrapper_cpp = os.path.join(tmp, "wrapper.cpp")
    run([sys.executable, "-m", "cython", "--cplus", "-3", "--module-name", "cpp_funs", pyx, "-o", wrapper_cpp])   # (setup.py: Extension "recometrics.cpp_funs")
    common = [cxx, "-std=c++11", "-O2", "-fPIC", "-fopenmp", "-w", "-c"]
    inc_py = ["-I", sysconfig.get_paths()["include"], "-I", numpy.get_include()]
    run(common + ["-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
                  "-DRMB200_REFERENCE_SIGNATURES_HPP=\"%s\"" % os.path.join(src, "recometrics_signatures.hpp"),
                  "-I", os.path.join(HERE, "ref_cython_signatures"), "-I", os.path.join(ROOT, "include")] + inc_py +
        [wrapper_cpp, "-o", os.path.join(tmp, "wrapper.o")])
    run([cxx, "-shared", "-fopenmp", os.path.join(tmp, "wrapper.o"), "-o", target,
         "-L", os.path.dirname(lib), "-lrecometrics_b200", "-Wl,-rpath,$ORIGIN/../../../recometrics_b200"])
    return target


if __name__ == "__main__":
    t = build(sys.argv[1] if len(sys.argv) > 1 else "/root/reference", verbose=True)
    print("built", t) if t else print("nothing built (reference or librecometrics_b200.so missing)")

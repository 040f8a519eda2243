"""
Build the B200 (sm_100a) backend library of meteotools_b200.

Drop-in for the reference's ``meteotools/build.py`` (:24-111), which shells out to
``gfortran -shared -fPIC -fopenmp -pthread -lgomp -o lib/fastcompute.so fastcompute.f90 -O3``.
This one compiles the hand-written CUDA sources under ``csrc/`` with nvcc for
``sm_100a`` into the SAME output name, ``<package>/lib/fastcompute.so``; the library
exports the 8 legacy symbols of ``fastcompute.f90`` plus the ``mt_*`` device API
declared in ``include/meteotools_b200.h``.

Execute:  python build.py            (or ``python -m meteotools_b200.build``)

Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only, no other targets, no PTX fallback
  -fmad=false                               no FMA contraction: +,-,*,/ round like the reference
  -lineinfo                                 so that ncu's source page maps to csrc/
"""
import os
import platform
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

path_of_meteotools = os.path.dirname(os.path.abspath(__file__))
platform_name = platform.system()

SOURCES = ["api.cu", "interp.cu", "legacy.cu", "interplevel.cu", "fd.cu", "thermo.cu", "diag.cu",
           "util.cu", "setdata_tma.cu", "reduce.cu", "spline.cu", "diag_tma.cu", "halo.cu", "peer.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-fmad=false", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


class task:
    """Same shape as the reference's ``build.task`` (build.py:24-98): a named build unit whose
    ``run_compile`` produces ``lib/<source_name>.so``."""

    def __init__(self, build_name="fastcompute", source_name="fastcompute", optimization="-O3",
                 verbose=False):
        self.build_name = build_name
        self.source_name = source_name
        self.optimization = optimization
        self.platform = platform_name
        self.lib_extension = ".dll" if self.platform == "Windows" else ".so"
        self.verbose = verbose
        self.csrc = os.path.join(path_of_meteotools, "csrc")
        self.include = os.path.join(os.path.dirname(path_of_meteotools), "include")
        self.objdir = os.path.join(path_of_meteotools, "lib", "obj")
        self.output = os.path.join(path_of_meteotools, "lib", self.source_name + self.lib_extension)

    @staticmethod
    def find_nvcc():
        for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
            if cand and os.path.exists(cand):
                return cand
        raise RuntimeError("nvcc not found: the meteotools_b200 backend cannot be built "
                           "(there is no CPU fallback)")

    def _stale(self, src, obj):
        if not os.path.exists(obj):
            return True
        t = os.path.getmtime(obj)
        deps = [src] + [str(p) for p in Path(self.csrc).glob("*.cuh")] + \
               [str(p) for p in Path(self.include).glob("*.h")]
        return any(os.path.getmtime(d) > t for d in deps)

    def _compile_one(self, nvcc, name):
        src = os.path.join(self.csrc, name)
        obj = os.path.join(self.objdir, name.replace(".cu", ".o"))
        if not self._stale(src, obj):
            return obj, ""
        cmd = [nvcc] + NVCC_FLAGS + ["-I", self.include, "-I", self.csrc, "-c", src, "-o", obj]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {name}:\n{proc.stdout}\n{proc.stderr}")
        return obj, proc.stderr

    def run_compile(self, force=False):
        nvcc = self.find_nvcc()
        Path(self.objdir).mkdir(parents=True, exist_ok=True)
        if force:
            for p in Path(self.objdir).glob("*.o"):
                p.unlink()
        with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
            results = list(ex.map(lambda n: self._compile_one(nvcc, n), SOURCES))
        objs = [r[0] for r in results]
        self.ptxas_log = "".join(r[1] for r in results)
        newest = max(os.path.getmtime(o) for o in objs)
        if force or not os.path.exists(self.output) or os.path.getmtime(self.output) < newest:
            cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", self.output] + objs + ["-ldl"]
            proc = subprocess.run(cmd, capture_output=True, text=True)
            if proc.returncode != 0:
                raise RuntimeError(f"link failed:\n{proc.stdout}\n{proc.stderr}")
        if self.verbose and self.ptxas_log:
            print(self.ptxas_log)
        return self.output


def build(force=False, verbose=False):
    return task(verbose=verbose).run_compile(force=force)


if __name__ == "__main__":
    Path(os.path.join(path_of_meteotools, "lib")).mkdir(parents=True, exist_ok=True)
    out = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(out)

"""Build recipe for programs written against the C++ front end (include/onika/**) -- the way a plugin is compiled:
nvcc, C++20, extended lambdas, sm_100a, linked against libonika_b200.so."""
from __future__ import annotations

import os
import subprocess
import sys

from . import _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")
CXX_TESTS = os.path.join(ROOT, "tests", "cxx")
PLUGIN_FLAGS = [
    "-std=c++20", "--extended-lambda", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O2",
    "--fmad=false", "-Xcompiler", "-fopenmp", "-Xcompiler", "-Wno-deprecated-declarations", "-I" + INCLUDE,
]
PROGRAMS = {"frontend_test": (["frontend_test.cu"], []), "frontend_test_generic": (["frontend_test.cu"], ["-DONIKA_B200_DISABLE_TYPED_FAST_PATHS"]),
            "multi_gpu_test": (["multi_gpu_test.cu"], [])}


def compile_program(sources, output, extra_flags=(), verbose=False):
    lib_dir = os.path.dirname(_build.build())
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    cmd = [_build.nvcc(), *PLUGIN_FLAGS, *extra_flags, "-o", output, *sources, "-L" + lib_dir, "-lonika_b200",
           "-Xlinker", "-rpath=" + lib_dir, "-Xlinker", "-rpath=$ORIGIN/../../onika_b200"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd, env=env)
    return output


def program_path(name: str) -> str:
    return os.path.join(CXX_TESTS, name)


def needs_build(name: str) -> bool:
    out = program_path(name)
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = [os.path.join(CXX_TESTS, f) for f in os.listdir(CXX_TESTS) if f.endswith((".cu", ".h"))]
    for base, _, files in os.walk(INCLUDE):
        deps += [os.path.join(base, f) for f in files]
    deps.append(_build.LIB)
    return any(os.path.getmtime(d) > t for d in deps)


REFERENCE_TUTORIAL = "/root/reference/plugins/onikaParallelTutorial"
REFPLUGIN_BIN = os.path.join(CXX_TESTS, "_refplugins", "run_reference_tutorial")


REFERENCE_MSP_DIR = os.path.join(os.path.dirname(REFPLUGIN_BIN), "data")
REFERENCE_TESTS = os.path.join(os.path.dirname(os.path.dirname(REFERENCE_TUTORIAL)), "tests")
REFERENCE_GPU_TESTS = ("gpu_uvm_benchmark", "gpu_reduce_min_fp64", "parallel_queue_lane")
REFERENCE_SOATL_TESTS = ("soatupletest", "soatlserializetest")


def reference_test_path(name: str) -> str:
    return os.path.join(os.path.dirname(REFPLUGIN_BIN), "ref_" + name)


def build_reference_plugins(force: bool = False, verbose: bool = False):
    """Compile the reference's OWN tutorial plugin sources, unchanged and from where they lie (never copied), against
    this repo's include/onika + the operator runner.  Only possible where /root/reference exists; the binary lands in the
    git-ignored tests/cxx/_refplugins/ and travels to the GPU box like oracle/_ref."""
    if not os.path.isdir(REFERENCE_TUTORIAL):
        return None
    srcs = [os.path.join(REFERENCE_TUTORIAL, f) for f in sorted(os.listdir(REFERENCE_TUTORIAL)) if f.endswith(".cu")]
    newest = max(os.path.getmtime(p) for p in srcs + [os.path.join(CXX_TESTS, "run_operator.cu"), _build.LIB])
    if not force and os.path.exists(REFPLUGIN_BIN) and os.path.getmtime(REFPLUGIN_BIN) > newest and not needs_build("frontend_test"):
        return REFPLUGIN_BIN
    os.makedirs(os.path.dirname(REFPLUGIN_BIN), exist_ok=True)
    # the reference's own stand-alone GPU tests of this path (tests/gpu_uvm_benchmark.cu, tests/gpu_reduce_min_fp64.cu,
    # tests/parallel_queue_lane.cu -- the latter through the include/onika/app/api.h boundary shim):
    # each is one translation unit with its own main(), compiled unchanged against include/onika
    for name in REFERENCE_GPU_TESTS:
        src = os.path.join(REFERENCE_TESTS, name + ".cu")
        if os.path.exists(src):
            compile_program([src], reference_test_path(name), verbose=verbose)
    # the reference's SOATL correctness tests (tests/soatupletest.cpp, tests/soatlserializetest.cpp; asserts enabled), with
    # the reference's own tests/declare_fields.h
    for name in REFERENCE_SOATL_TESTS:
        src = os.path.join(REFERENCE_TESTS, name + ".cpp")
        if os.path.exists(src):
            compile_program([src], reference_test_path(name), extra_flags=["-x", "cu", "-I" + REFERENCE_TESTS], verbose=verbose)
    # the reference's own application input files (data/config/main-config.msp, data/exemples/*.msp), unchanged: like the
    # binaries above they land in the git-ignored directory and travel to the GPU box with it
    import shutil
    ref_data = os.path.join(os.path.dirname(os.path.dirname(REFERENCE_TUTORIAL)), "data")
    for sub in ("config", "exemples"):
        dst = os.path.join(REFERENCE_MSP_DIR, sub)
        os.makedirs(dst, exist_ok=True)
        for f in sorted(os.listdir(os.path.join(ref_data, sub))):
            if f.endswith(".msp"):
                shutil.copyfile(os.path.join(ref_data, sub, f), os.path.join(dst, f))
    return compile_program([os.path.join(CXX_TESTS, "run_operator.cu")] + srcs, REFPLUGIN_BIN, verbose=verbose)


MIXED_TU_BIN = os.path.join(CXX_TESTS, "mixed_tu_test")


def build_mixed_tu_test(force: bool = False, verbose: bool = False):
    """one program from a g++-compiled and an nvcc-compiled translation unit that share the default CUDA context
    (tests/cxx/mixed_tu_host.cpp + mixed_tu_device.cu): the layouts of the shared types must not depend on the compiler"""
    if not force and not needs_build("mixed_tu_test"):
        return MIXED_TU_BIN
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    obj = os.path.join(CXX_TESTS, "mixed_tu_host.o")
    cmd = ["g++", "-std=c++20", "-O2", "-fopenmp", "-c", "-I" + INCLUDE, os.path.join(CXX_TESTS, "mixed_tu_host.cpp"), "-o", obj]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd, env=env)
    return compile_program([os.path.join(CXX_TESTS, "mixed_tu_device.cu"), obj], MIXED_TU_BIN, verbose=verbose)


def build_all(force: bool = False, verbose: bool = False):
    outs = []
    build_reference_plugins(force, verbose)
    for name, (srcs, flags) in PROGRAMS.items():
        if force or needs_build(name):
            compile_program([os.path.join(CXX_TESTS, s) for s in srcs], program_path(name), extra_flags=flags, verbose=verbose)
        outs.append(program_path(name))
    outs.append(build_mixed_tu_test(force, verbose))
    return outs


if __name__ == "__main__":
    print(build_all(force=True, verbose=True))

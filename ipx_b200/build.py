"""Build recipes for the native parts of ipx_b200 (in-tree, sm_100a only).

* ``ipx_b200/_build/libipxgpu.so``   CUDA kernels + C ABI (include/ipxgpu.h); needs nvcc only.
* ``ipx_b200/_build/libipx_gpu.so``  IPX with seven TUs (the hot path and Maxvolume) replaced by the GPU
  drop-ins of ipx_b200/host; needs the reference tree (compiled against its
  UNMODIFIED headers), so it is built where /root/reference exists and travels
  to the GPU box as a prebuilt file.
"""

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG)
OUT = os.path.join(PKG, "_build")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
REF = os.environ.get("IPX_REFERENCE", "/root/reference")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

LIBIPXGPU = os.path.join(OUT, "libipxgpu.so")
LIBIPX_GPU = os.path.join(OUT, "libipx_gpu.so")

# Reference TUs replaced by ipx_b200/host/*_gpu.cc.
REPLACED = ["normal_matrix", "diagonal_precond", "conjugate_residuals", "splitted_normal_matrix",
            "kkt_solver_diag", "kkt_solver_basis", "maxvolume"]
ABSENT = ["basiclu_wrapper", "basiclu_kernel"]  # need the un-vendored BASICLU
SHIMS = ["lu_provider", "sparse_lu", "lapack_min"]
# Reference translation units compiled unchanged but with one free function renamed, so that
# the drop-in build can put its own definition in front of it (multiply_add_gpu.cc).
RENAMED = {"sparse_matrix": ["-DMultiplyAdd=MultiplyAdd_reference"]}
EXTRA_HOST = ["multiply_add_gpu"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout)
        raise RuntimeError("build step failed: " + cmd[0])
    return r.stdout


def nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build_libipxgpu(force=False, verbose=False):
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    srcs.append(os.path.join(REPO, "include", "ipxgpu.h"))
    if not force and not _newer(LIBIPXGPU, srcs):
        return LIBIPXGPU
    cmd = [nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
           "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-ccbin", CXX,
           "-I", "/usr/include", "--fmad=false",
           os.path.join(CSRC, "ipxgpu.cu"), "-o", LIBIPXGPU, "-lcudart", "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    out = _run(cmd)
    if verbose:
        print(out)
    return LIBIPXGPU


def have_reference():
    return os.path.exists(os.path.join(REF, "src", "normal_matrix.cc"))


def build_libipx_gpu(force=False):
    """IPX with the GPU drop-ins; requires the reference tree for headers and
    for the TUs that are NOT on the hot path (compiled in place, never copied)."""
    if not have_reference():
        if os.path.exists(LIBIPX_GPU):
            return LIBIPX_GPU
        raise RuntimeError("reference tree absent and no prebuilt libipx_gpu.so")
    build_libipxgpu()
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    flags = ["-std=c++11", "-O2", "-fPIC", f"-I{REF}/include", f"-I{REF}/src",
             f"-I{REPO}/include", f"-I{HOST}"]
    objs = []
    jobs = []
    for f in sorted(os.listdir(os.path.join(REF, "src"))):
        if not f.endswith(".cc"):
            continue
        stem = f[:-3]
        if stem in REPLACED or stem in ABSENT:
            continue
        if stem in RENAMED:
            jobs.append((os.path.join(REF, "src", f),
                         os.path.join(OUT, "obj", "ref_" + stem + "_renamed.o"), RENAMED[stem]))
        else:
            jobs.append((os.path.join(REF, "src", f),
                         os.path.join(OUT, "obj", "ref_" + stem + ".o"), []))
    for stem in SHIMS + [r + "_gpu" for r in REPLACED] + EXTRA_HOST + ["gpu_bridge"]:
        jobs.append((os.path.join(HOST, stem + ".cc"), os.path.join(OUT, "obj", stem + ".o"), []))
    headers = [os.path.join(HOST, h) for h in os.listdir(HOST) if h.endswith(".h")]
    headers.append(os.path.join(REPO, "include", "ipxgpu.h"))
    procs = []
    for src, obj, extra in jobs:
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            procs.append((src, subprocess.Popen([CXX] + flags + extra + ["-c", src, "-o", obj],
                                                stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                                text=True)))
            if len(procs) >= 8:
                _drain(procs)
    _drain(procs)
    if force or _newer(LIBIPX_GPU, objs + [LIBIPXGPU]):
        _run([CXX, "-shared", "-Wl,-Bsymbolic", "-o", LIBIPX_GPU] + objs +
             [f"-L{OUT}", "-lipxgpu", "-Wl,-rpath,$ORIGIN"])
    # Drop-in proof: the reference's own example and Catch suite, compiled from the
    # reference tree unmodified, linked against the GPU build.
    link = [f"-L{OUT}", "-lipx_gpu", "-lipxgpu", "-Wl,-rpath,$ORIGIN"]
    afiro = os.path.join(OUT, "afiro_gpu")
    if force or _newer(afiro, [LIBIPX_GPU]):
        _run([CXX] + flags + [os.path.join(REF, "example", "afiro.cc"), "-o", afiro] + link)
    check = os.path.join(OUT, "ipx_check_gpu")
    if force or _newer(check, [LIBIPX_GPU]):
        srcs = [os.path.join(REF, "check", f) for f in sorted(os.listdir(os.path.join(REF, "check")))
                if f.endswith(".cc")]
        _run([CXX] + flags + [f"-I{REF}/third_party"] + srcs + ["-o", check] + link)
    return LIBIPX_GPU


def _drain(procs):
    while procs:
        src, p = procs.pop(0)
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("compile failed: " + src)


if __name__ == "__main__":
    build_libipxgpu(force="--force" in sys.argv, verbose="-v" in sys.argv)
    if have_reference() and "--kernels-only" not in sys.argv:
        build_libipx_gpu(force="--force" in sys.argv)

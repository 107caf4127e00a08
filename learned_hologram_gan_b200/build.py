"""Build libasm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""

from __future__ import annotations

import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libasm_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = glob.glob(os.path.join(CSRC, "*")) + [os.path.join(HERE, "..", "include", "asm_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def _compile_objects(defines=(), tag="", verbose=False):
    """One nvcc -c per translation unit, in parallel; an object is rebuilt only when its source, a header of
    csrc/ or the public headers are newer (the objects live in git-ignored build/)."""
    from concurrent.futures import ThreadPoolExecutor

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    objdir = os.path.join(HERE, "build" + tag)
    os.makedirs(objdir, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    newest_header = max(os.path.getmtime(h) for h in headers)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + [f"-D{d}" for d in defines]
    jobs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if os.path.isfile(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), newest_header):
            jobs.append((obj, None))
        else:
            jobs.append((obj, [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]))

    def run(job):
        obj, cmd = job
        if cmd is None:
            return obj, ""
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        done = list(ex.map(run, jobs))
    if verbose:
        for _, err in done:
            print(err)
    return [o for o, _ in done]


def _link(objs, out):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", out] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    if force:
        shutil.rmtree(os.path.join(HERE, "build"), ignore_errors=True)
    return _link(_compile_objects(verbose=verbose), LIB)


def build_variant(name: str, defines) -> str:
    """Experiment builds (tools/): the same sources with extra -D flags into lib/libasm_b200_<name>.so."""
    out = os.path.join(HERE, "lib", f"libasm_b200_{name}.so")
    return _link(_compile_objects(defines=tuple(defines), tag="_" + name), out)


if __name__ == "__main__":
    print(build(force=True, verbose=True))

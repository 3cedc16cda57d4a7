"""TEST INFRASTRUCTURE: builds oracle/liboracle_kmat.so (gcc + OpenMP) from oracle/kmat.c and binds it with ctypes.
__graft_entry__.build() calls build(); the library is git-ignored (*.so) but travels to the GPU box with the tree."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "kmat.c")
LIB = os.path.join(HERE, "liboracle_kmat.so")
_lib = None


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    tmp = "%s.%d.tmp" % (LIB, os.getpid())
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-o", tmp, SRC, "-lm"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + res.stdout)
    os.replace(tmp, LIB)
    return LIB


def load():
    global _lib
    if _lib is None:
        try:
            build()
        except Exception:
            if not os.path.exists(LIB):
                raise
        lib = C.CDLL(LIB)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        lib.oracle_kmat.argtypes = [C.c_int, C.c_double, C.c_int, ip, ip, dp, dp, C.c_long, dp, C.c_long, C.c_int, dp]
        lib.oracle_kmat.restype = None
        lib.oracle_colsumsq.argtypes = [dp, C.c_long, C.c_long, dp]
        lib.oracle_colsumsq.restype = None
        lib.oracle_grad_trace.argtypes = [C.c_int, C.c_double, C.c_int, ip, ip, dp, dp, C.c_long, C.c_int, dp, dp]
        lib.oracle_grad_trace.restype = None
        _lib = lib
    return _lib


if __name__ == "__main__":
    print(build(force=True))

"""ctypes binding of tests/native/libcvb_test_hooks.so (test-only tcgen05 self-test / micro-benchmarks; see cvb_test_hooks.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_vp, _i = C.c_void_p, C.c_int


def load():
    import cyclevae_vc_b200._lib  # noqa: F401  (the hooks link against the product library: load it first, RTLD_GLOBAL)
    lib = C.CDLL(os.path.join(_HERE, "libcvb_test_hooks.so"))
    lib.cvb_selftest_umma.restype = _i
    lib.cvb_selftest_umma.argtypes = [_i, _i, _i, _vp, _vp, _vp, _vp]
    lib.cvb_bench_ingest.restype = _i
    lib.cvb_bench_ingest.argtypes = [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp]
    lib.cvb_bench_allgather.restype = _i
    lib.cvb_bench_allgather.argtypes = [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]
    return lib

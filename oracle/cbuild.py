"""Build / load the C restatement of the reference CPU path (oracle/lr_oracle.c).  Test and baseline
infrastructure only: imported by tests/, __graft_entry__ and bench.py, never by luaradio_b200/."""
import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_float, c_int, c_long, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))


def build(native=False):
    target = "liblr_oracle_native.so" if native else "liblr_oracle.so"
    path = os.path.join(HERE, target)
    src = os.path.join(HERE, "lr_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, target], check=True, capture_output=True)
    return path


def load(native=False):
    try:
        path = build(native)
    except Exception:
        if not native:
            raise
        path = build(False)
    lib = ctypes.CDLL(path)
    lib.lro_wbfm_chain.restype = c_long
    lib.lro_wbfm_chain.argtypes = [c_void_p, c_long, c_uint64, c_void_p, c_int, c_int, c_double, c_float,
                                   c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_long]
    lib.lro_fir_crcf_parallel.restype = None
    lib.lro_fir_crcf_parallel.argtypes = [c_void_p, c_long, c_void_p, c_int, c_void_p, c_int]
    lib.lro_max_threads.restype = c_int
    return lib

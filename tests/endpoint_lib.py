"""ctypes loader for tests/native (endpoint driver + the oracle pair-ops table).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
NATIVE = os.path.join(HERE, "native")


def load(pkg, need_oracle):
    """Returns (driver, (oracle_lib, oracle_ops_ptr) or None).  Builds tests/native on first use."""
    if not os.path.exists(pkg.ENDPOINT_LIB_PATH):
        pkg.build()
    if need_oracle and not os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(os.path.dirname(HERE), "oracle"), "oracle"])
    targets = ["libendpoint_driver.so"] + (["liboracle_pair_ops.so"] if need_oracle else [])
    subprocess.check_call(["make", "-s", "-C", NATIVE] + targets)
    C.CDLL(pkg.LIB_PATH, mode=C.RTLD_GLOBAL)
    C.CDLL(pkg.ENDPOINT_LIB_PATH, mode=C.RTLD_GLOBAL)
    D = C.CDLL(os.path.join(NATIVE, "libendpoint_driver.so"))
    u64 = C.c_uint64
    D.drv_read_and_write.restype = C.c_int
    D.drv_read_and_write.argtypes = [C.c_void_p, u64, u64, u64, C.c_int, C.c_int, C.c_int, C.POINTER(u64)]
    D.drv_shutdown_sequence.restype = C.c_int
    D.drv_shutdown_sequence.argtypes = [C.c_void_p, C.c_int]
    D.drv_peer_close.restype = C.c_int
    D.drv_peer_close.argtypes = [C.c_void_p, C.c_int, C.c_int]
    D.drv_echo.restype = C.c_int
    D.drv_echo.argtypes = [C.c_void_p, C.c_int, u64, u64, C.c_int, C.c_int, C.c_int, C.POINTER(u64)]
    D.drv_two_threads.restype = C.c_int
    D.drv_two_threads.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(u64)]
    D.drv_multi_echo.restype = C.c_int
    D.drv_multi_echo.argtypes = [C.c_void_p, C.c_int, C.c_int, u64, u64, C.c_int, C.c_int, C.c_int, C.POINTER(u64)]
    ops = None
    if need_oracle:
        O = C.CDLL(os.path.join(NATIVE, "liboracle_pair_ops.so"))
        O.oracle_pair_ops.restype = C.c_void_p
        O.oracle_ops_config.argtypes = [u64, C.c_int]
        O.oracle_pair_ops_batch.restype = C.c_void_p
        O.oracle_pair_ops_async.restype = C.c_void_p
        ops = (O, O.oracle_pair_ops())
    return D, ops

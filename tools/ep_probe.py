"""Experiment helper: the endpoint streaming driver (tools/native/ep_stream.cc) at chosen sizes.
   python tools/ep_probe.py CONNS THREAD_PAIRS MSGS POOL [MSG_BYTES]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

pkg = ge.load_package()
L = pkg.lib()
pkg.init(0)
conns, threads, msgs, pool = (int(x) for x in sys.argv[1:5])
msg = int(sys.argv[5]) if len(sys.argv) > 5 else 4 << 20
pkg.config_set("GRPC_RDMA_RING_BUFFER_SIZE_KB", 16384)
C.CDLL(pkg.ENDPOINT_LIB_PATH, mode=C.RTLD_GLOBAL)
ES = C.CDLL(os.path.join(os.path.dirname(pkg.LIB_PATH), "libb200_epstream.so"))
ES.ep_stream_run.restype = C.c_double
ES.ep_stream_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
assert L.b200_service_start(pool) == 0, pkg.last_error()
print("service up", flush=True)
o = (C.c_uint64 * 4)()
t0 = time.time()
t = ES.ep_stream_run(None, conns, threads, msgs, 2, msg, 0, o)
print("conns %d threads %d msgs %d pool %d: t=%.4f s -> %.2f GB/s payload, bad=%d, submits c/s %d/%d (wall %.1f s)"
      % (conns, threads, msgs, pool, t, (o[0] / t / 1e9) if t > 0 else 0, o[1], o[2], o[3], time.time() - t0), flush=True)
L.b200_service_stop()
print("stopped", flush=True)

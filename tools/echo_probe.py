"""Why is the endpoint echo slow?  launches and seconds per echo run (experiment helper)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import __graft_entry__ as ge
import endpoint_lib

pkg = ge.load_package()
pkg.init(0)
L = pkg.lib()
D, _ = endpoint_lib.load(pkg, need_oracle=False)
pkg.config_set("B200_RING_BUFFER_SIZE_BYTES", 1 << 20)
for n, mx, busy, poller in ((4, 200000, 200, 0), (4, 2000000, 200, 0), (4, 2000000, 0, 1)):
    l0 = L.b200_launch_count()
    nb = C.c_uint64(0)
    t0 = time.time()
    rc = D.drv_echo(None, n, mx, 4242, busy, poller, 1, C.byref(nb))
    dt = time.time() - t0
    nl = L.b200_launch_count() - l0
    print("echo n=%d max=%d busy=%d poller=%d rc=%d: %.2f s, %d bytes, %d launches, %.1f us/launch, %.0f B/launch"
          % (n, mx, busy, poller, rc, dt, nb.value, nl, dt / max(nl, 1) * 1e6, 2 * nb.value / max(nl, 1)), flush=True)

"""Host-only benchmark of the FASTQ reader (tdg_fastq_next): reads/s for a given file and thread count."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tagdust_b200 import _capi

path = sys.argv[1]
lib = _capi.load_library()
for threads in [int(x) for x in (sys.argv[2:] or ["1", "4", "8"])]:
    h = C.c_void_p()
    lib.tdg_fastq_open(path.encode(), -1, C.byref(h))
    ch = _capi.FastqChunkC()
    t = time.time(); n = 0
    while True:
        lib.tdg_fastq_next(h, 303104, threads, C.byref(ch))
        if ch.n == 0:
            break
        n += ch.n
    dt = time.time() - t
    print(threads, "threads", n, "reads", round(dt, 3), "s", round(n / dt / 1e6, 2), "M reads/s")
    lib.tdg_fastq_close(h)

"""CPU test of integration/emit_fast.c: the calibration read emitters (emit_read_sequence /
emit_random_sequence, barcode_hmm.c:2599-3046) with precomputed thresholds must emit exactly the reads
the reference's emitters do for the same srand() seed.  Each side runs in its own process: the fast
emitters are interposed by loading integration/_build/libemit_fast.so ahead of the reference library."""
import os
import subprocess
import sys

import numpy as np
import pytest

from refharness import have_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAST = os.path.join(ROOT, "integration", "_build", "libemit_fast.so")

CHILD = r"""
import ctypes, sys, os, zlib
import numpy as np
root, fast, name, seed, use_fast = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
if use_fast:
    ctypes.CDLL(fast, mode=ctypes.RTLD_GLOBAL)      # its emit_* / free_model_bag now precede the reference's
from refharness import RefHarness
from cases import CASES, build_ref_model
R = RefHarness()
avg = CASES[name]["read_len"]
p, mb, desc = build_ref_model(R, name, avg_len=avg, max_len=400)
R.model_calibration_edit(mb, p)
codes, lens = R.emit(mb, 1500, 1500, avg, seed, 4096)
R.model_free(mb)
# a second model right away: the table cache must follow the model_bag, not its address
p2, mb2, _ = build_ref_model(R, "b4_r", avg_len=30, max_len=400)
R.model_calibration_edit(mb2, p2)
c2, l2 = R.emit(mb2, 300, 300, 30, seed + 1, 4096)
print(zlib.crc32(codes.tobytes()), zlib.crc32(lens.tobytes()), int(lens.sum()), int(lens.max()),
      zlib.crc32(c2.tobytes()), zlib.crc32(l2.tobytes()))
"""


def run_child(name, seed, use_fast):
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT, FAST, name, str(seed), str(int(use_fast))], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r.stdout.strip().split()


@pytest.mark.parametrize("name,seed", [("b48_r", 42), ("f_s_b_r", 7), ("p_b_r_p", 1234), ("o_b_s_r", 99), ("g_b2_r", 5), ("s20_b_r", 2024)])
def test_fast_emitters_emit_the_same_reads(name, seed):
    if not have_ref() or not os.path.exists(FAST):
        pytest.skip("oracle/_ref or integration/_build not built")
    want = run_child(name, seed, False)
    got = run_child(name, seed, True)
    assert got == want
    assert int(want[2]) > 0


def test_rand_replica_matches_libc():
    """integration/rand_glibc.c: lock-free srand()/rand() with glibc's TYPE_3 generator -- same sequence as libc."""
    import ctypes
    if not os.path.exists(FAST):
        pytest.skip("integration/_build not built")
    mine = ctypes.CDLL(FAST)                       # RTLD_LOCAL: only looked up through this handle
    libc = ctypes.CDLL("libc.so.6")
    for seed in (0, 1, 42, 1234, 2**31 - 1, 2**31, 2**32 - 1, 987654321):
        mine.srand(ctypes.c_uint(seed)); libc.srand(ctypes.c_uint(seed))
        a = [mine.rand() for _ in range(20000)]
        b = [libc.rand() for _ in range(20000)]
        assert a == b, seed

"""CPU tests of the C-ABI library: it loads, exports every declared symbol, and its host-side
pieces (logsum table, architecture compiler, model derivation) agree with the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cases import CASES, GOLDEN, bits, build_ref_model
from refharness import background_logp
from tagdust_b200 import _capi
from tagdust_b200.api import Context, TagdustError, compile_architecture
from test_oracle import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    lib = _capi.load_library()
    hdr = open(os.path.join(ROOT, "include", "tagdust_b200.h")).read()
    declared = set(re.findall(r"\b(tdg_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no prototypes found in the header"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/tagdust_b200.h but not exported"
    assert declared == set(_capi.PROTOTYPES), declared ^ set(_capi.PROTOTYPES)


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(TagdustError) as e:
        Context(1)
    assert e.value.code == _capi.TDG_ENODEV
    assert "no CPU fallback" in str(e.value)


def test_logsum_table_bits(ref):
    lib = _capi.load_library()
    t = np.zeros(16000, np.float32)
    lib.tdg_logsum_table(t.ctypes.data_as(_capi.c_float_p))
    assert np.array_equal(t.view(np.uint32), ref.logsum_table().view(np.uint32))


@pytest.mark.parametrize("name", list(CASES))
def test_arch_compile_matches_golden_model(name):
    """tdg_arch_compile vs the model the reference's init_model_bag built (stored in the golden file)."""
    z, gold = load_golden(name)
    c = CASES[name]
    L = c["read_len"]
    bg = background_logp((2501.0, 2480.0, 2510.0, 2492.0, 21.0))
    mine = compile_architecture(c["segments"], bg, float(L), L + 8, five=c.get("five", (0, 0, 0)), three=c.get("three", (0, -1, -1)))
    assert gold.same_bits(mine) == []


@pytest.mark.parametrize("name", list(CASES))
def test_arch_compile_matches_live_reference(ref, name):
    c = CASES[name]
    L = c["read_len"]
    bg = background_logp((3100.0, 1900.0, 2200.0, 2800.0, 7.0))
    for e, i in ((0.05, 0.1), (0.02, 0.25)):
        p = ref.param_new(c["segments"], e=e, i=i)
        mb = ref.model_new(p, background=bg, average_length=float(L + 3), max_seq_len=L + 8,
                           five=c.get("five", (0, 0, 0)), three=c.get("three", (0, -1, -1)))
        want = ref.flatten(mb, p)
        mine = compile_architecture(c["segments"], bg, float(L + 3), L + 8, e=e, i=i, five=c.get("five", (0, 0, 0)),
                                    three=c.get("three", (0, -1, -1)))
        assert want.same_bits(mine) == []
        ref.model_calibration_edit(mb, p)
        want = ref.flatten(mb, p)
        mine = compile_architecture(c["segments"], bg, float(L + 3), L + 8, e=e, i=i, five=c.get("five", (0, 0, 0)),
                                    three=c.get("three", (0, -1, -1)), calibration_edit=True)
        assert want.same_bits(mine) == []
        ref.model_free(mb); ref.param_free(p)


def test_arch_compile_rejects_bad_segment():
    with pytest.raises(TagdustError):
        compile_architecture(["X:ACGT", "R:N"], background_logp(), 50.0, 60)
    with pytest.raises(TagdustError):
        compile_architecture(["B:ACGT,AC", "R:N"], background_logp(), 50.0, 60)


def test_model_validate_reports_structure():
    lib = _capi.load_library()
    z, gold = load_golden("b48_r")
    buf = C.create_string_buffer(300)
    assert lib.tdg_model_validate(C.byref(gold.c), buf, 300) == 0
    msg = buf.value.decode()
    assert "H=50" in msg and "C=295" in msg and "std_segments=1" in msg and "dp_structured=1" in msg


def test_shard_plan():
    lib = _capi.load_library()
    for n, nd in ((0, 1), (1, 1), (100, 3), (75776 * 5 + 17, 8), (64, 8), (33, 2)):
        first = (C.c_int32 * nd)(); cnt = (C.c_int32 * nd)()
        assert lib.tdg_plan_shards(n, nd, first, cnt) == 0
        assert sum(cnt) == n
        pos = 0
        for k in range(nd):
            if cnt[k]:
                assert first[k] == pos and first[k] % 32 == 0
                pos += cnt[k]


def test_live_ops_count_what_the_kernels_execute():
    """tdg_desc_live_ops (bench.py's roofline numerator): terms whose transition is log(0) are not counted.  cfg2 shape:
    SURVEY 8d's algorithmic count is 8 + 10 logsums per (column, position) = 108 per 6-column HMM and position; the
    standard B-segment pattern leaves 47 live."""
    import bench
    from tagdust_b200.api import compile_architecture, live_ops
    segs, _ = bench.architecture()
    desc = compile_architecture(segs, bench.background(), 150.0, 150)
    lo = live_ops(desc)
    H, Cn = desc.total_hmms, desc.total_columns
    per_hmm = (lo["ls_bwd"] + lo["ls_fwd"]) / H
    assert 44 < per_hmm < 48, per_hmm
    assert lo["ls_bwd"] < 8 * Cn and lo["ls_fwd"] < 10 * Cn
    assert lo["add_bwd"] > 0 and lo["add_fwd"] > 0
    # a model with a skippable segment and partial adapters has more live terms per column than the standard pattern
    desc2 = compile_architecture(["P:GGGGGGG", "B:ACGT,TTGA", "O:N", "R:N"], bench.background(), 60.0, 60, five=(7.0, 6.2, 1.1))
    lo2 = live_ops(desc2)
    assert lo2["ls_fwd"] / desc2.total_columns > lo["ls_fwd"] / Cn

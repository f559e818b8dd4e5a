"""CPU tests (-m "not gpu"): the oracle against the reference's golden vectors, and against the
live reference library where oracle/_ref exists."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from cases import CASES, GOLDEN, bits, build_ref_model, make_case_reads
from tagdust_b200._capi import MODE_ARCH_COMP, MODE_GET_LABEL, MODE_GET_PROB, ModelDesc


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"vectors_{name}.npz"))
    desc = ModelDesc(z["seg_type"].tobytes(), z["model_seg_num_hmms"], z["model_seg_num_cols"], z["model_seg_skip"],
                     z["model_background"], z["model_transition"], z["model_m_emit"], z["model_i_emit"],
                     z["model_silent_to_M"], z["model_silent_to_I"], z["model_label"], z["model_transition_matrix"],
                     int(z["average_raw_length"]))
    return z, desc


def test_logsum_table_matches_reference(oracle, ref):
    assert np.array_equal(oracle.logsum_table().view(np.uint32), ref.logsum_table().view(np.uint32))


def test_logsum_pointwise(oracle, ref):
    rng = np.random.default_rng(0)
    a = (rng.random(4000) * -40).astype(np.float32)
    b = (a - rng.random(4000).astype(np.float32) * 17).astype(np.float32)
    a[:5] = -np.inf
    b[3:9] = -np.inf
    for x, y in zip(a, b):
        r, o = np.float32(ref.logsum(float(x), float(y))), np.float32(oracle.logsum(float(x), float(y)))
        assert r.view(np.uint32) == o.view(np.uint32)
        r, o = np.float32(ref.logsum(float(y), float(x))), np.float32(oracle.logsum(float(y), float(x)))
        assert r.view(np.uint32) == o.view(np.uint32)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_golden_vectors(oracle, name):
    """The committed vectors were produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
    z, desc = load_golden(name)
    out = oracle.run(desc, MODE_GET_LABEL, z["codes"], z["lens"], threshold=float(z["threshold"]), minlen=16, dust=100, threads=4)
    for k in ("f_score", "b_score", "r_score", "bar_prob", "mapq"):
        assert np.array_equal(bits(out[k]), bits(z[k])), k
    for k in ("read_type", "barcode", "fingerprint"):
        assert np.array_equal(out[k], z[k]), k
    lens = z["lens"]
    for r in range(len(lens)):
        assert np.array_equal(out["labels"][r, : lens[r] + 1], z["labels"][r, : lens[r] + 1])
        assert np.array_equal(out["seq"][r, : lens[r]], z["seq_out"][r, : lens[r]])
    assert np.array_equal(out["len"], z["len_out"])


@pytest.mark.parametrize("name", ["b4_r", "o_b_s_r", "p_b_r_p", "g_b2_r"])
def test_oracle_matches_live_reference(oracle, ref, name):
    codes, lens, _ = make_case_reads(name, 400, seed=77, len_jitter=6, n_frac=0.03)
    p, mb, desc = build_ref_model(ref, name, threshold=2.0, minlen=10, dust=20, max_len=int(lens.max()) + 2)
    r = ref.run_phmm(mb, p, 1, codes, lens)
    s = ref.decode_scores(mb, codes, lens)
    o = oracle.run(desc, MODE_GET_LABEL, codes, lens, threshold=2.0, minlen=10, dust=20, threads=3)
    assert np.array_equal(bits(o["mapq"]), bits(r["mapq"]))
    for k in ("f_score", "b_score", "r_score"):
        assert np.array_equal(bits(o[k]), bits(s[k]))
    for k in ("read_type", "barcode", "fingerprint", "len"):
        assert np.array_equal(o[k], r[k]), k
    assert np.array_equal(o["seq"], r["seq"])
    ref.model_free(mb); ref.param_free(p)


def test_oracle_get_prob_mode(oracle, ref):
    codes, lens, _ = make_case_reads("b4_r", 200, seed=5)
    p, mb, desc = build_ref_model(ref, "b4_r")
    r = ref.run_phmm(mb, p, 4, codes, lens)
    o = oracle.run(desc, MODE_GET_PROB, codes, lens)
    assert np.array_equal(bits(o["mapq"]), bits(r["mapq"]))
    assert np.array_equal(bits(o["bar_prob"]), bits(r["bar_prob"].astype(np.float32)))
    ref.model_free(mb); ref.param_free(p)


@pytest.mark.parametrize("threads", [1, 3])
def test_oracle_arch_compare(oracle, ref, threads):
    """MODE_ARCH_COMP: float sums depend on the thread slicing; same slicing -> same bits."""
    names = ["b4_r", "p_b_r_p", "o_b_s_r"]
    codes, lens, _ = make_case_reads("b4_r", 240, seed=9, read_len=40)
    built = [build_ref_model(ref, n, avg_len=40, max_len=48, threads=threads) for n in names]
    post_ref = ref.run_arch_comp([b[1] for b in built], built[0][0], codes, lens)
    bs, post = oracle.arch_compare([b[2] for b in built], codes, lens, threads=threads)
    assert np.array_equal(bits(post), bits(post_ref))
    for k, b in enumerate(built):
        assert np.array_equal(bits(bs[k]), bits(ref.backward_scores(b[1], codes, lens)))
        ref.model_free(b[1]); ref.param_free(b[0])


def test_windowed_matchstart_matchend(oracle, ref):
    """-start/-end: the HMM sees seq+matchstart for matchend-matchstart residues (barcode_hmm.c:2292-2296)."""
    codes, lens, _ = make_case_reads("b4_r", 150, seed=3, len_jitter=0)
    c = CASES["b4_r"]
    p = ref.param_new(c["segments"], threshold=1.0, minlen=5, dust=100, matchstart=2, matchend=22)
    mb = ref.model_new(p, average_length=20.0, max_seq_len=30)
    desc = ref.flatten(mb, p)
    # shift the reads right by two bases so the barcode starts at matchstart
    sh = np.zeros_like(codes); sh[:, 2:] = codes[:, :-2]; sh[:, :2] = 3
    lens2 = lens + 2
    r = ref.run_phmm(mb, p, 1, sh, lens2)
    o = oracle.run(desc, MODE_GET_LABEL, sh, lens2, threshold=1.0, minlen=5, dust=100, matchstart=2, matchend=22)
    assert np.array_equal(bits(o["mapq"]), bits(r["mapq"]))
    for k in ("read_type", "barcode", "fingerprint"):
        assert np.array_equal(o[k], r[k]), k
    ref.model_free(mb); ref.param_free(p)

"""-ref artifact filter: the oracle's restatement of match_to_reference (barcode_hmm.c:2478-2583) and of
the two Myers bit-vector variants it calls (misc.c:572-636, :718-796) against the unmodified reference (CPU tests), and the
CUDA path against the oracle (-m gpu)."""
import numpy as np
import pytest

from cases import CASES, build_ref_model, make_case_reads
from tagdust_b200.api import MODE_GET_LABEL


def make_reference_set(rng, codes, lens, n_from_reads=6, n_random=5, skip=6):
    """Reference sequences: a few windows cut out of the reads' bodies (forward or reverse-complemented, some with an
    edit), plus random ones.  Returns (flat codes, s_index)."""
    seqs = []
    for k in range(n_from_reads):
        r = int(rng.integers(0, len(lens)))
        body = codes[r, skip:lens[r]].copy()
        if k % 2:
            body = (3 - body[::-1]) % 4          # reverse complement (bases only)
        if k % 3 == 0 and len(body) > 10:
            body[len(body) // 2] = (body[len(body) // 2] + 1) % 4
        pad = rng.integers(0, 4, size=int(rng.integers(0, 30))).astype(np.uint8)
        seqs.append(np.concatenate([pad, body.astype(np.uint8), pad[::-1]]))
    for _ in range(n_random):
        seqs.append(rng.integers(0, 4, size=int(rng.integers(20, 300))).astype(np.uint8))
    s_index = np.zeros(len(seqs) + 1, np.int32)
    s_index[1:] = np.cumsum([len(x) for x in seqs])
    return np.concatenate(seqs).astype(np.uint8), s_index


@pytest.mark.parametrize("which", ["bmp_single", "bpm_check_error"])
def test_myers_variants_match_reference(oracle, ref, which):
    rng = np.random.default_rng(17)
    for trial in range(400):
        n = int(rng.integers(1, 260))
        m = int(rng.integers(1, 64 if which == "bpm_check_error" else 140))   # bpm_check_error: indices >= 64 are UB in the reference
        t = rng.integers(0, 5, size=n).astype(np.uint8)
        p = rng.integers(0, 5, size=m).astype(np.uint8)
        if trial % 3 == 0:
            k = int(rng.integers(0, m))
            p[:k] = 65                                   # spacer prefix, as make_extracted_read leaves it
        if trial % 5 == 0 and n > m:
            s = int(rng.integers(0, n - m + 1))
            t[s:s + m] = np.where(p == 65, t[s:s + m], p)  # plant the pattern
        assert oracle.myers(which, t, p) == ref.myers(which, t, p), (which, trial, n, m)


R_FIRST = ["R:N", "B:" + ",".join(CASES["b4_r"]["barcodes"])]   # the read comes first: the pattern starts with bases, not spacers


def r_first_reads(n, seed, body=70):
    """body + 3' barcode.  With a 5' barcode the rewritten read starts with spacers, which no reference character matches,
    so the reference's filter can never fire there: every spacer inside the first 63 residues costs one edit.  For the
    same reason the body is longer than 63 nt here (the 3' barcode's spacers fall outside the pattern)."""
    from tagdust_b200 import synth
    rng = np.random.default_rng(seed)
    tags = [synth.encode(t) for t in CASES["b4_r"]["barcodes"]]
    L = body + 6
    codes = np.zeros((n, 96), np.uint8)
    lens = np.full(n, L, np.int32)
    for r in range(n):
        codes[r, :body] = rng.integers(0, 4, size=body)
        codes[r, body:L] = tags[int(rng.integers(0, len(tags)))]
        if r % 9 == 0:
            codes[r, :L] = rng.integers(0, 4, size=L)
        if r % 13 == 0:
            lens[r] = L - int(rng.integers(1, 5))
            codes[r, lens[r]:] = 0
        if r % 41 == 7:
            codes[r, :body] = 2                        # low complexity body: dust after the artifact match
    return codes, lens


def r_first_model(ref, threads, threshold=0.5):
    p = ref.param_new(R_FIRST, threshold=threshold, minlen=8, dust=100, threads=threads)
    from refharness import background_logp
    mb = ref.model_new(p, background=background_logp((2501.0, 2480.0, 2510.0, 2492.0, 21.0)), average_length=76.0, max_seq_len=90)
    return p, mb, ref.flatten(mb, p)


@pytest.mark.parametrize("threads", [1, 3, 4])
def test_label_run_with_reference_matches_reference(oracle, ref, threads):
    codes, lens = r_first_reads(203, seed=3)
    rng = np.random.default_rng(5)
    rc, si = make_reference_set(rng, codes, lens - 6, n_from_reads=10, skip=0)
    p, mb, desc = r_first_model(ref, threads)
    ref.set_reference(p, rc, si, 2)
    want = ref.run_phmm_ref(mb, p, codes, lens)
    ref.set_reference(p, [], [], 2)
    oracle.set_reference(rc, si, 2)
    got = oracle.run(desc, MODE_GET_LABEL, codes, lens, threshold=0.5, minlen=8, dust=100, threads=threads)
    oracle.set_reference([], [], 2)
    ref.model_free(mb); ref.param_free(p)
    assert (want["read_type"] >> 8).max() > 0, "test set has no artifact hits"
    assert np.array_equal(got["read_type"], want["read_type"])
    assert np.array_equal(got["barcode"], want["barcode"])


@pytest.mark.parametrize("threads", [1, 5])
def test_rna_dust_with_reference_matches_reference(oracle, ref, threads):
    rng = np.random.default_rng(9)
    n, L = 131, 70
    codes = np.zeros((n, 80), np.uint8)
    lens = rng.integers(20, L + 1, size=n).astype(np.int32)
    for r in range(n):
        codes[r, :lens[r]] = rng.integers(0, 4, size=lens[r])
    codes[5, :lens[5]] = 0                                # low complexity
    rc, si = make_reference_set(rng, codes, lens, n_from_reads=12, skip=0)
    p = ref.param_new(["R:N"], threads=threads, dust=100)
    ref.set_reference(p, rc, si, 2)
    want = ref.run_rna_dust(p, codes, lens)
    ref.set_reference(p, [], [], 2)
    ref.param_free(p)
    oracle.set_reference(rc, si, 2)
    got = oracle.run_rna_dust(codes, lens, dust=100, threads=threads)
    oracle.set_reference([], [], 2)
    assert (want >> 8).max() > 0 and (want == 6).any()
    assert np.array_equal(got, want)


# ---------------------------------------------------------------- CUDA path (k_artifact) vs the oracle
@pytest.mark.gpu
@pytest.mark.parametrize("threads", [1, 3, 4, 7])
def test_gpu_label_run_with_reference(gpu_ctx, oracle, ref, threads):
    codes, lens = r_first_reads(1203, seed=31)
    rng = np.random.default_rng(7)
    rc, si = make_reference_set(rng, codes, lens - 6, n_from_reads=24, skip=0)
    p, mb, desc = r_first_model(ref, threads)
    oracle.set_reference(rc, si, 2)
    want = oracle.run(desc, MODE_GET_LABEL, codes, lens, threshold=0.5, minlen=8, dust=100, threads=threads)
    oracle.set_reference([], [], 2)
    ref.model_free(mb); ref.param_free(p)
    model = gpu_ctx.model(desc, int(lens.max()))
    batch = gpu_ctx.batch(len(lens), int(lens.max()))
    batch.append(codes, lens)
    rs = gpu_ctx.refset(rc, si)
    got = gpu_ctx.run_phmm(model, batch, MODE_GET_LABEL, threshold=0.5, minlen=8, dust=100, refset=rs, filter_error=2,
                           slice_threads=threads)
    rs.close(); batch.close(); model.close()
    assert (want["read_type"] >> 8).max() > 0
    assert np.array_equal(got["read_type"], want["read_type"])
    assert np.array_equal(got["barcode"], want["barcode"])


@pytest.mark.gpu
@pytest.mark.parametrize("threads,cut", [(1, 2), (5, 2), (4, 0), (3, 6)])
def test_gpu_rna_dust_with_reference(gpu_ctx, oracle, threads, cut):
    rng = np.random.default_rng(19)
    n, L = 2031, 140
    codes = np.zeros((n, 144), np.uint8)
    lens = rng.integers(1, L + 1, size=n).astype(np.int32)
    for r in range(n):
        codes[r, :lens[r]] = rng.integers(0, 5 if r % 17 == 0 else 4, size=lens[r])
    codes[5, :lens[5]] = 0
    rc, si = make_reference_set(rng, codes, lens, n_from_reads=40, n_random=20, skip=0)
    oracle.set_reference(rc, si, cut)
    want = oracle.run_rna_dust(codes, lens, dust=100, threads=threads)
    oracle.set_reference([], [], 2)
    batch = gpu_ctx.batch(n, L)
    batch.append(codes, lens)
    rs = gpu_ctx.refset(rc, si)
    got = gpu_ctx.rna_dust(batch, dust=100, refset=rs, filter_error=cut, slice_threads=threads)
    none = gpu_ctx.rna_dust(batch, dust=100)                      # no reference: dust only
    rs.close(); batch.close()
    assert (want >> 8).max() > 0
    assert np.array_equal(got, want)
    oracle.set_reference([], [], 2)
    assert np.array_equal(none, oracle.run_rna_dust(codes, lens, dust=100, threads=threads))

"""GPU parity (the -m gpu tests proper): CUDA path through the C ABI vs the oracle / the reference."""
import numpy as np
import pytest

from cases import CASES, bits, build_ref_model, make_case_reads
from tagdust_b200.api import MODE_GET_LABEL, MODE_GET_PROB

pytestmark = pytest.mark.gpu

SCORE_KEYS = ("b_score", "f_score", "r_score", "bar_prob", "mapq")


def compare(gpu, ora, lens, mode, name):
    n = len(lens)
    report = {}
    for k in SCORE_KEYS:
        report[k] = int((bits(gpu[k]) != bits(ora[k])).sum())
    lab_bad = 0
    for r in range(n):
        if not np.array_equal(gpu["labels"][r, : lens[r] + 1], ora["labels"][r, : lens[r] + 1]):
            lab_bad += 1
    report["labels"] = lab_bad
    if mode == MODE_GET_LABEL:
        for k in ("read_type", "barcode", "fingerprint"):
            report[k] = int((gpu[k] != ora[k]).sum())
    return report


@pytest.mark.parametrize("name", list(CASES))
def test_label_parity_vs_oracle(gpu_ctx, oracle, ref, name):
    n = 1500
    codes, lens, _ = make_case_reads(name, n)
    p, mb, desc = build_ref_model(ref, name)
    max_len = int(lens.max())
    thr = 1.5
    ora = oracle.run(desc, MODE_GET_LABEL, codes, lens, threshold=thr, minlen=16, dust=100, threads=8)
    model = gpu_ctx.model(desc, max_len)
    batch = gpu_ctx.batch(n, max_len)
    batch.append(codes, lens)
    gpu = gpu_ctx.run_phmm(model, batch, MODE_GET_LABEL, threshold=thr, minlen=16, dust=100)
    rep = compare(gpu, ora, lens, MODE_GET_LABEL, name)
    batch.close(); model.close(); ref.model_free(mb); ref.param_free(p)
    assert all(v == 0 for v in rep.values()), f"{name}: mismatches {rep}"


def run_gpu(gpu_ctx, desc, codes, lens, mode, **kw):
    max_len = int(lens.max()) if len(lens) else 1
    model = gpu_ctx.model(desc, max(max_len, 1))
    batch = gpu_ctx.batch(max(len(lens), 1), max(max_len, 1))
    if len(lens):
        batch.append(codes, lens)
    out = gpu_ctx.run_phmm(model, batch, mode, **kw)
    batch.close(); model.close()
    return out


@pytest.mark.parametrize("name", list(CASES))
def test_golden_vectors(gpu_ctx, name):
    """CUDA path vs the committed vectors produced by the UNMODIFIED reference."""
    from test_oracle import load_golden
    z, desc = load_golden(name)
    lens = z["lens"]
    gpu = run_gpu(gpu_ctx, desc, z["codes"], lens, MODE_GET_LABEL, threshold=float(z["threshold"]), minlen=16, dust=100)
    for k in SCORE_KEYS:
        assert np.array_equal(bits(gpu[k]), bits(z[k])), k
    for k in ("read_type", "barcode", "fingerprint"):
        assert np.array_equal(gpu[k], z[k]), k
    for r in range(len(lens)):
        assert np.array_equal(gpu["labels"][r, : lens[r] + 1], z["labels"][r, : lens[r] + 1])
    # extracted flag <-> the reference rewrote the sequence (spacer 65 appears or read kept whole)
    assert np.array_equal(gpu["extracted"].astype(bool), z["read_type"] == 0) or (z["read_type"] == 6).any()


def test_get_prob_mode(gpu_ctx, oracle, ref):
    codes, lens, _ = make_case_reads("f_s_b_r", 700, seed=4)
    p, mb, desc = build_ref_model(ref, "f_s_b_r")
    ora = oracle.run(desc, MODE_GET_PROB, codes, lens, threads=8)
    for want_labels in (True, False):
        gpu = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_PROB, want_labels=want_labels)
        for k in SCORE_KEYS:
            assert np.array_equal(bits(gpu[k]), bits(ora[k])), k
        if want_labels:
            for r in range(len(lens)):
                assert np.array_equal(gpu["labels"][r, : lens[r] + 1], ora["labels"][r, : lens[r] + 1])
    ref.model_free(mb); ref.param_free(p)


def test_calibration_model_and_reads(gpu_ctx, oracle, ref):
    """Threshold calibration path (calibrateQ.c): reads emitted by the reference's emitters from the
    edited model, scored in MODE_GET_PROB.  Emitted reads are longer than the average length."""
    p, mb, desc0 = build_ref_model(ref, "b4_r", avg_len=26, max_len=400)
    ref.model_calibration_edit(mb, p)
    codes, lens = ref.emit(mb, 300, 300, 26, 42, 512)
    ref.model_free(mb)
    mb = ref.model_new(p, average_length=26.0, max_seq_len=int(lens.max()))
    desc = ref.flatten(mb, p)
    want = ref.run_phmm(mb, p, 4, codes, lens)
    gpu = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_PROB, want_labels=False)
    assert np.array_equal(bits(gpu["mapq"]), bits(want["mapq"]))
    assert np.array_equal(bits(gpu["bar_prob"]), bits(want["bar_prob"].astype(np.float32)))
    ref.model_free(mb); ref.param_free(p)


@pytest.mark.parametrize("threads", [1, 8])
def test_arch_compare(gpu_ctx, oracle, ref, threads):
    names = ["b4_r", "p_b_r_p", "o_b_s_r", "g_b2_r"]
    codes, lens, _ = make_case_reads("b4_r", 1000, seed=9, read_len=40)
    built = [build_ref_model(ref, n, avg_len=40, max_len=48, threads=threads) for n in names]
    post_ref = ref.run_arch_comp([b[1] for b in built], built[0][0], codes, lens)
    models = [gpu_ctx.model(b[2], 48) for b in built]
    batch = gpu_ctx.batch(len(lens), 48)
    batch.append(codes, lens)
    bs, post = gpu_ctx.arch_compare(models, batch, num_threads=threads)
    for k, b in enumerate(built):
        assert np.array_equal(bits(bs[k]), bits(ref.backward_scores(b[1], codes, lens))), names[k]
    assert np.array_equal(bits(post), bits(post_ref))
    assert int(np.argmax(post)) == 0
    batch.close()
    for m in models:
        m.close()
    for b in built:
        ref.model_free(b[1]); ref.param_free(b[0])


def test_window_start_end(gpu_ctx, oracle, ref):
    codes, lens, _ = make_case_reads("b4_r", 500, seed=3, len_jitter=0)
    c = CASES["b4_r"]
    p = ref.param_new(c["segments"], threshold=1.0, minlen=5, dust=100, matchstart=2, matchend=22)
    mb = ref.model_new(p, average_length=20.0, max_seq_len=30)
    desc = ref.flatten(mb, p)
    sh = np.zeros_like(codes); sh[:, 2:] = codes[:, :-2]; sh[:, :2] = 3
    lens2 = lens + 2
    ora = oracle.run(desc, MODE_GET_LABEL, sh, lens2, threshold=1.0, minlen=5, dust=100, matchstart=2, matchend=22, threads=4)
    gpu = run_gpu(gpu_ctx, desc, sh, lens2, MODE_GET_LABEL, threshold=1.0, minlen=5, dust=100, matchstart=2, matchend=22)
    for k in SCORE_KEYS:
        assert np.array_equal(bits(gpu[k]), bits(ora[k])), k
    for k in ("read_type", "barcode", "fingerprint"):
        assert np.array_equal(gpu[k], ora[k]), k
    ref.model_free(mb); ref.param_free(p)


def test_ragged_short_and_ambiguous_reads(gpu_ctx, oracle, ref):
    """Edge cases: lengths from the shortest the model can emit up to the maximum, N-rich reads,
    homopolymers (dust), thresholds that reject everything / nothing, minlen larger than the read."""
    rng = np.random.default_rng(8)
    p, mb, desc = build_ref_model(ref, "b4_r", max_len=120)
    n = 1200
    lens = rng.integers(4, 121, size=n).astype(np.int32)
    lens[:8] = [4, 5, 6, 7, 8, 120, 120, 119]
    stride = 128
    codes = np.zeros((n, stride), np.uint8)
    for r in range(n):
        codes[r, : lens[r]] = rng.integers(0, 4, size=lens[r])
    codes[100:200] = np.where(rng.random((100, stride)) < 0.5, 4, codes[100:200])   # N-rich
    for r in range(200, 260):
        codes[r, : lens[r]] = rng.integers(0, 4)                                     # homopolymers
    for r in range(n):
        codes[r, lens[r]:] = 0
    for thr, minlen, dust in ((0.0, 16, 100), (39.0, 16, 100), (1.5, 200, 0), (1.5, 1, 5)):
        ora = oracle.run(desc, MODE_GET_LABEL, codes, lens, threshold=thr, minlen=minlen, dust=dust, threads=8)
        gpu = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_LABEL, threshold=thr, minlen=minlen, dust=dust)
        rep = compare(gpu, ora, lens, MODE_GET_LABEL, "edge")
        assert all(v == 0 for v in rep.values()), (thr, minlen, dust, rep)
    ref.model_free(mb); ref.param_free(p)


def test_empty_batch(gpu_ctx, ref):
    p, mb, desc = build_ref_model(ref, "b4_r")
    out = run_gpu(gpu_ctx, desc, np.zeros((0, 32), np.uint8), np.zeros(0, np.int32), MODE_GET_LABEL, threshold=1.0)
    assert out["mapq"].shape == (0,)
    ref.model_free(mb); ref.param_free(p)


def test_too_long_read_is_rejected(gpu_ctx, ref):
    from tagdust_b200.api import TagdustError
    p, mb, desc = build_ref_model(ref, "b4_r")
    model = gpu_ctx.model(desc, 30)
    batch = gpu_ctx.batch(4, 64)
    codes = np.zeros((1, 64), np.uint8)
    batch.append(codes, np.array([50], np.int32))
    with pytest.raises(TagdustError):
        gpu_ctx.run_phmm(model, batch, MODE_GET_LABEL, threshold=1.0)
    batch.close(); model.close(); ref.model_free(mb); ref.param_free(p)


def test_multi_wave_and_sampled_parity_full_size(gpu_ctx, oracle):
    """BASELINE cfg2 at full shape (150 nt, 48 barcodes), several waves of 75 776 reads:
    size-independent properties + oracle parity on a random sample of the reads."""
    from tagdust_b200 import synth
    from tagdust_b200.api import compile_architecture
    from refharness import background_logp
    from cases import TAGS6_ED3
    tags = TAGS6_ED3[:48]
    desc = compile_architecture(["B:" + ",".join(tags), "R:N"], background_logp((2.5e6, 2.5e6, 2.5e6, 2.5e6, 1.0)), 150.0, 150)
    n = 75776 * 2 + 1234
    codes, lens, truth = synth.make_reads_fast(n, 150, tags, error_rate=0.01, random_frac=0.05, seed=77)
    gpu = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_LABEL, threshold=1.5, minlen=16, dust=100)
    # properties: forward == backward likelihood up to float noise; labels monotone in segment; truth recovered
    assert np.all(np.abs(gpu["f_score"] - gpu["b_score"]) < 2e-2)
    ok = gpu["read_type"] == 0
    assert ok[truth >= 0].mean() > 0.99
    sel = ok & (truth >= 0)
    assert ((gpu["barcode"][sel] & 0xFFFF) == truth[sel]).mean() > 0.995
    lab = gpu["labels"][:, 1:151]
    seg = (np.asarray(desc.label)[lab] & 0xFFFF)
    assert np.all(np.diff(seg.astype(np.int32), axis=1) >= 0)          # segments never go backwards
    # idempotence: same batch again -> identical bits
    gpu2 = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_LABEL, threshold=1.5, minlen=16, dust=100)
    for k in SCORE_KEYS:
        assert np.array_equal(bits(gpu[k]), bits(gpu2[k]))
    # oracle parity on a sample spread over all waves (incl. the ragged last wave)
    idx = np.concatenate([np.arange(0, n, 997), np.arange(n - 40, n)])
    ora = oracle.run(desc, MODE_GET_LABEL, codes[idx], lens[idx], threshold=1.5, minlen=16, dust=100, threads=8)
    sub = {k: v[idx] for k, v in gpu.items()}
    rep = compare(sub, ora, lens[idx], MODE_GET_LABEL, "cfg2")
    assert all(v == 0 for v in rep.values()), rep


def test_two_device_context_matches_single(oracle, ref):
    """Sharding over the GPUs of one context: contiguous tile-aligned shards, output order = input order."""
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tagdust_b200.api import Context
    ctx = Context(2)
    assert ctx.device_count == 2
    codes, lens, _ = make_case_reads("b_b_r", 3000, seed=12)
    p, mb, desc = build_ref_model(ref, "b_b_r")
    ora = oracle.run(desc, MODE_GET_LABEL, codes, lens, threshold=1.5, minlen=16, dust=100, threads=8)
    gpu = run_gpu(ctx, desc, codes, lens, MODE_GET_LABEL, threshold=1.5, minlen=16, dust=100)
    rep = compare(gpu, ora, lens, MODE_GET_LABEL, "2gpu")
    assert all(v == 0 for v in rep.values()), rep
    ctx.close(); ref.model_free(mb); ref.param_free(p)


def test_small_waves_and_long_reads(gpu_ctx, oracle, ref, monkeypatch):
    """Wave size is an execution detail: forcing 3-CTA waves (the path taken when the scratch of a
    long-read model does not fit in HBM, e.g. threshold-calibration reads) gives identical bits; reads of
    several hundred nt use the same kernels."""
    rng = np.random.default_rng(21)
    p, mb, desc = build_ref_model(ref, "b4_r", avg_len=300, max_len=700)
    n = 4000
    lens = rng.integers(200, 651, size=n).astype(np.int32)
    codes = np.zeros((n, 656), np.uint8)
    tags = [synth_encode(t) for t in CASES["b4_r"]["barcodes"]]
    for r in range(n):
        codes[r, : lens[r]] = rng.integers(0, 4, size=lens[r])
        if r % 10:
            codes[r, :6] = tags[r % 4]
    base = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_LABEL, threshold=1.0, minlen=16, dust=100)
    monkeypatch.setenv("TDG_WAVE_CTAS", "3")
    small = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_LABEL, threshold=1.0, minlen=16, dust=100)
    monkeypatch.delenv("TDG_WAVE_CTAS")
    for k in SCORE_KEYS + ("read_type", "barcode", "fingerprint"):
        assert np.array_equal(bits(base[k]), bits(small[k])), k
    assert np.array_equal(base["labels"], small["labels"])
    idx = np.arange(0, n, 40)
    ora = oracle.run(desc, MODE_GET_LABEL, codes[idx], lens[idx], threshold=1.0, minlen=16, dust=100, threads=8)
    rep = compare({k: v[idx] for k, v in base.items()}, ora, lens[idx], MODE_GET_LABEL, "long")
    assert all(v == 0 for v in rep.values()), rep
    ref.model_free(mb); ref.param_free(p)


def synth_encode(s):
    from tagdust_b200.synth import encode
    return encode(s)


@pytest.mark.parametrize("env", ["TDG_GENERIC_LABEL_DP", "TDG_NO_STDU", "TDG_NO_SMEM_STATE"])
@pytest.mark.parametrize("name", ["b48_r", "f_s_b_r", "b_b_r", "o_b_s_r", "s20_b_r", "f12_b12_r", "p18_b_r_p14"])
def test_alternative_kernel_paths_give_the_same_bits(gpu_ctx, oracle, ref, name, env, monkeypatch):
    """Every read must come out the same whichever code path the host picks: the verbatim O(L*H^2) label DP
    instead of the structured one, the generic run-time-mask columns instead of the standard-pattern ones,
    thread-local instead of shared-memory profile state."""
    n = 600
    codes, lens, _ = make_case_reads(name, n, seed=5)
    p, mb, desc = build_ref_model(ref, name)
    kw = dict(threshold=1.5, minlen=16, dust=100)
    ora = oracle.run(desc, MODE_GET_LABEL, codes, lens, threads=8, **kw)
    monkeypatch.setenv(env, "1")
    gpu = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_LABEL, **kw)
    monkeypatch.delenv(env)
    rep = compare(gpu, ora, lens, MODE_GET_LABEL, name)
    ref.model_free(mb); ref.param_free(p)
    assert all(v == 0 for v in rep.values()), f"{name} with {env}: mismatches {rep}"


def spans_from_labels(labels, lens, desc, extracted, stride):
    """make_extracted_read (barcode_hmm.c:3325-3356) restated on the label rows: the runs of residues j whose
    labels[j + 1] lies in an R segment."""
    is_r = np.array([desc.seg_type[int(lab) & 0xFFFF:(int(lab) & 0xFFFF) + 1] == b"R" for lab in desc.label])
    out = np.zeros((len(lens), stride, 2), np.uint16)
    for r in range(len(lens)):
        if not extracted[r]:
            continue
        m = is_r[labels[r, 1:lens[r] + 1]]
        k, j = 0, 0
        while j < lens[r]:
            if not m[j]:
                j += 1
                continue
            s = j
            while j < lens[r] and m[j]:
                j += 1
            out[r, k] = (s, j - s)
            k += 1
    return out


@pytest.mark.parametrize("name", ["b48_r", "p_b_r_p", "o_b_s_r", "f_s_b_r", "p18_b_r_p14"])
def test_spans_equal_label_rows(gpu_ctx, ref, name):
    """want_spans: the R-run table the demux writer works from is exactly what the label rows say; and with
    want_labels=0 (rows stay on the device) every other output keeps its bits."""
    codes, lens, _ = make_case_reads(name, 1200, seed=23, read_len=60 if name == "b48_r" else None)
    p, mb, desc = build_ref_model(ref, name)
    full = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_LABEL, threshold=1.0, minlen=8, dust=100, want_spans=True)
    stride = full["spans"].shape[1]
    assert stride == 1 + sum(1 for t in desc.seg_type if t == ord("R"))
    want = spans_from_labels(full["labels"], lens, desc, full["extracted"], stride)
    assert full["extracted"].sum() > 100
    assert np.array_equal(full["spans"], want)
    lean = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_LABEL, threshold=1.0, minlen=8, dust=100, want_labels=False, want_spans=True)
    assert "labels" not in lean
    for k in SCORE_KEYS:
        assert np.array_equal(bits(lean[k]), bits(full[k])), k
    for k in ("read_type", "barcode", "fingerprint", "extracted", "spans"):
        assert np.array_equal(lean[k], full[k]), k
    ref.model_free(mb); ref.param_free(p)


def test_spans_with_window(gpu_ctx, ref):
    """-start/-end: labels behind the window are 0, so an R segment at HMM 0's segment adds a trailing run."""
    name = "b4_r"
    codes, lens, _ = make_case_reads(name, 600, seed=29, read_len=40, len_jitter=0)
    p, mb, desc = build_ref_model(ref, name, max_len=48)
    out = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_LABEL, threshold=0.5, minlen=8, dust=100, matchstart=0, matchend=30, want_spans=True)
    want = spans_from_labels(out["labels"], lens, desc, out["extracted"], out["spans"].shape[1])
    assert np.array_equal(out["spans"], want)
    ref.model_free(mb); ref.param_free(p)


def _rand_tags(rng, n, L):
    seen, out = set(), []
    while len(out) < n:
        t = "".join(rng.choice(list("ACGT"), size=L))
        if t not in seen:
            seen.add(t); out.append(t)
    return out


def test_long_linker_segment_90_columns(gpu_ctx, oracle):
    """A 90-nt S: linker (90-column HMMs): beyond the unrolled and the shared-memory-state paths, the column loop keeps
    its profile state in thread-local memory (the reference allocates any length, barcode_hmm.c:4591-4674)."""
    from refharness import background_logp
    from tagdust_b200 import synth
    from tagdust_b200.api import compile_architecture
    rng = np.random.default_rng(77)
    linker = "".join(rng.choice(list("ACGT"), size=90))
    tags = ["ACGTAC", "TTGACC", "GGCATT", "CATGCA", "TACGGA", "AGTCTG"]
    segs = ["S:" + linker, "B:" + ",".join(tags), "R:N"]
    desc = compile_architecture(segs, background_logp((2501.0, 2480.0, 2510.0, 2492.0, 21.0)), 150.0, 160)
    assert max(desc.seg_num_cols) >= 90
    codes, lens, _ = synth.make_reads(700, 150, [linker + t for t in tags], error_rate=0.02, random_frac=0.1, seed=5, len_jitter=4, n_frac=0.005)
    ora = oracle.run(desc, MODE_GET_LABEL, codes, lens, threshold=1.0, minlen=16, dust=100, threads=8)
    gpu = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_LABEL, threshold=1.0, minlen=16, dust=100)
    rep = compare(gpu, ora, lens, MODE_GET_LABEL, "s90")
    assert all(v == 0 for v in rep.values()), rep
    assert (gpu["read_type"] == 0).mean() > 0.7


@pytest.mark.parametrize("forced", [False, True])
def test_model_larger_than_shared_memory(gpu_ctx, oracle, monkeypatch, forced):
    """130 sixteen-nt barcodes: 131 x 16 + 1 = 2 097 columns, more than fit beside the logsum table in shared memory
    (~1 880): the kernels read the tables from global memory.  H = 132 is beyond the reference's `total_prob[100]`, so the
    check is against the port.  forced=True sends a small model down the same kernels (TDG_MODEL_IN_GLOBAL)."""
    from refharness import background_logp
    from tagdust_b200 import synth
    from tagdust_b200.api import compile_architecture
    rng = np.random.default_rng(78)
    if forced:
        monkeypatch.setenv("TDG_MODEL_IN_GLOBAL", "1")
        tags = _rand_tags(rng, 6, 8)
        n, L = 900, 60
    else:
        tags = _rand_tags(rng, 130, 16)
        n, L = 400, 80
    segs = ["B:" + ",".join(tags), "R:N"]
    desc = compile_architecture(segs, background_logp((2501.0, 2480.0, 2510.0, 2492.0, 21.0)), float(L), L + 8)
    if not forced:
        assert desc.total_columns > 2000 and desc.total_hmms == 132
    codes, lens, truth = synth.make_reads(n, L, tags, error_rate=0.02, random_frac=0.1, seed=6, len_jitter=3)
    ora = oracle.run(desc, MODE_GET_LABEL, codes, lens, threshold=1.0, minlen=16, dust=100, threads=8)
    gpu = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_LABEL, threshold=1.0, minlen=16, dust=100)
    rep = compare(gpu, ora, lens, MODE_GET_LABEL, "bigmodel")
    assert all(v == 0 for v in rep.values()), rep
    sel = (gpu["read_type"] == 0) & (truth >= 0)
    assert sel.sum() > 0.6 * (truth >= 0).sum() and ((gpu["barcode"][sel] & 0xFFFF) == truth[sel]).mean() > 0.98

"""BASELINE.json configs 3, 4 and 5 at their real architecture shapes, through the C ABI:
several waves of reads on the GPU, oracle parity on a sample spread over the waves, plus
size-independent properties.  (Config 1 = tests/test_gold_dropin.py, config 2 = test_gpu_parity.py.)

Reference limits honoured: config 3 uses 95 barcodes (1 + 2 + 96 + 1 = 100 HMMs,
`float total_prob[100]`, barcode_hmm.c:4186); config 4 puts both index segments on read 1."""
import numpy as np
import pytest

from cases import TAGS6_ED3, bits
from refharness import background_logp
from tagdust_b200 import synth
from tagdust_b200.api import MODE_GET_LABEL, compile_architecture
from test_gpu_parity import SCORE_KEYS, compare, run_gpu

pytestmark = pytest.mark.gpu

BG = background_logp((2.5e6, 2.5e6, 2.5e6, 2.5e6, 1.0))
WAVE = 148 * 512
LINKER = "ACGTTGCAGTCA"


def sample_parity(oracle, desc, codes, lens, gpu, n, step, **kw):
    idx = np.concatenate([np.arange(0, n, step), np.arange(n - 24, n)])
    ora = oracle.run(desc, MODE_GET_LABEL, codes[idx], lens[idx], threads=8, **kw)
    sub = {k: v[idx] for k, v in gpu.items()}
    rep = compare(sub, ora, lens[idx], MODE_GET_LABEL, "sample")
    assert all(v == 0 for v in rep.values()), rep


def test_cfg3_umi_linker_95_barcodes(gpu_ctx, oracle):
    """-1 F:NNNNNNNN -2 S:<12 nt linker> -3 B:<95 barcodes> -4 R:N on 150 nt reads (H = 100, C = 8 + 24 + 576 + 1)."""
    tags = TAGS6_ED3[:95]
    segs = ["F:NNNNNNNN", "S:" + LINKER, "B:" + ",".join(tags), "R:N"]
    desc = compile_architecture(segs, BG, 150.0, 150)
    assert desc.total_hmms == 100
    n = WAVE + 3001
    codes, lens, truth = synth.make_reads(n, 150, [LINKER + t for t in tags], umi_len=8, error_rate=0.01, random_frac=0.05, seed=31)
    kw = dict(threshold=1.5, minlen=16, dust=100)
    gpu = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_LABEL, **kw)
    ok = gpu["read_type"] == 0
    model = truth >= 0
    assert ok[model].mean() > 0.98
    sel = ok & model
    assert ((gpu["barcode"][sel] & 0xFFFF) == truth[sel]).mean() > 0.995
    assert (gpu["barcode"][sel] >> 16 == 2).all()                       # segment index of the B segment
    # UMI: 8 bases, 2 bits each, length byte 8 (extract_reads :3212-3215, :3262-3275)
    fp = gpu["fingerprint"][sel]
    assert ((fp & 0xFF) == 8).all()
    umi = np.zeros(n, np.int64)
    for k in range(8):
        umi = (umi << 2) | (codes[:, k] & 3)
    assert ((fp >> 8) == umi[sel]).mean() > 0.97                         # 1 % errors / indel-shifted labels allowed
    assert np.all(np.abs(gpu["f_score"] - gpu["b_score"]) < 5e-2)
    sample_parity(oracle, desc, codes, lens, gpu, n, 331, **kw)


def test_cfg4_dual_index(gpu_ctx, oracle):
    """-1 B:<24 I7 tags> -2 B:<16 I5 tags> -3 R:N on 150 nt reads (H = 43, C = 253); 384 combinations,
    compared per segment through the labels (ri->barcode keeps only the last B segment, :3216-3225)."""
    i7, i5 = TAGS6_ED3[:24], TAGS6_ED3[24:40]
    segs = ["B:" + ",".join(i7), "B:" + ",".join(i5), "R:N"]
    desc = compile_architecture(segs, BG, 150.0, 150)
    assert desc.total_hmms == 43 and desc.total_columns == 253
    n = WAVE + 2000
    codes, lens, truth = synth.make_reads(n, 150, i7, second_barcodes=i5, error_rate=0.01, random_frac=0.05, seed=41)
    kw = dict(threshold=1.5, minlen=16, dust=100)
    gpu = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_LABEL, **kw)
    ok = (gpu["read_type"] == 0) & (truth >= 0)
    assert ok[truth >= 0].mean() > 0.98
    lab = gpu["labels"][:, 1:151].astype(np.int32)
    hmm_seg = np.asarray(desc.label) & 0xFFFF
    hmm_idx = (np.asarray(desc.label) >> 16) & 0x7FFF
    first = np.full(n, -1); second = np.full(n, -1)
    for r in np.nonzero(ok)[0][:4000]:
        s = hmm_seg[lab[r]]
        a = lab[r][s == 0]; b = lab[r][s == 1]
        if len(a): first[r] = hmm_idx[a[0]]
        if len(b): second[r] = hmm_idx[b[0]]
    chk = np.nonzero(ok)[0][:4000]
    combo = first[chk] * 16 + second[chk]
    assert (combo == truth[chk]).mean() > 0.99
    assert ((gpu["barcode"][chk] & 0xFFFF) == second[chk]).all() and ((gpu["barcode"][chk] >> 16) == 1).all()
    sample_parity(oracle, desc, codes, lens, gpu, n, 257, **kw)


def candidate_architectures(n_arch):
    return synth.candidate_architectures(TAGS6_ED3, n_arch)


def test_cfg5_architecture_detection(gpu_ctx, oracle):
    """test_architectures: 64 candidate architectures scored by backward() over a read sample,
    posteriors = per-thread float sums normalised with logsum (do_arch_comparison :2111-2148, merge :1994-2017)."""
    archs = candidate_architectures(64)
    descs = [compile_architecture(s, BG, 150.0, 150) for s in archs]
    tags = TAGS6_ED3[:48]
    n = 20000
    codes, lens, _ = synth.make_reads_fast(n, 150, tags, error_rate=0.01, random_frac=0.05, seed=51)
    models = [gpu_ctx.model(d, 150) for d in descs]
    batch = gpu_ctx.batch(n, 150)
    batch.append(codes, lens)
    bs, post = gpu_ctx.arch_compare(models, batch, num_threads=8)
    assert int(np.argmax(post)) == 5                                   # the generating architecture wins
    assert np.isclose(np.exp(post.astype(np.float64)).sum(), 1.0, atol=1e-3)
    # oracle parity on a prefix (b_score per (architecture, read) and the posteriors of that prefix)
    m = 160
    small = gpu_ctx.batch(m, 150)
    small.append(codes[:m], lens[:m])
    bs_s, post_s = gpu_ctx.arch_compare(models, small, num_threads=3)
    want_bs, want_post = oracle.arch_compare(descs, codes[:m], lens[:m], threads=3)
    assert np.array_equal(bits(bs_s), bits(want_bs))
    assert np.array_equal(bits(post_s), bits(want_post))
    assert np.array_equal(bits(bs[:, :m]), bits(want_bs))               # independent of the batch it was scored in
    small.close(); batch.close()
    for mo in models:
        mo.close()


def test_cfg2_full_size_properties(gpu_ctx):
    """BASELINE config 2 at (a multiple of) its full size through size-independent properties: the run is
    streamed in batches of 32 waves with double buffering like a real job.  Default 20 M reads; set
    TDG_FULL_SIZE=1 for the 100 M reads the config names.  Properties: the tallies of independently
    generated batches follow the generating distribution (uniform barcodes within 2 %, the expected share of the
    5 % contaminants picked up by chance), forward == backward likelihood, a checksum of per-read results is independent of
    how the reads were batched, and every batch reproduces itself bit for bit when re-run at the end."""
    import os
    import zlib
    from tagdust_b200.api import MODE_GET_LABEL
    tags = TAGS6_ED3[:48]
    desc = compile_architecture(["B:" + ",".join(tags), "R:N"], BG, 150.0, 150)
    total = 100_000_000 if os.environ.get("TDG_FULL_SIZE") else 20_000_000
    per = 32 * WAVE
    nb = (total + per - 1) // per
    model = gpu_ctx.model(desc, 150)
    batches = [gpu_ctx.batch(per, 150) for _ in range(2)]
    kw = dict(threshold=1.5, minlen=16, dust=100)
    counts = np.zeros(49, np.int64)
    n_model = n_model_ok = n_random = n_random_assigned = 0
    first_crc = None
    first_in = None
    pending = None

    def consume(res, truth):
        nonlocal n_model, n_model_ok, n_random, n_random_assigned
        ok = res["read_type"] == 0
        bc = res["barcode"] & 0xFFFF
        m = truth >= 0
        n_model += int(m.sum()); n_model_ok += int((ok & m & (bc == truth)).sum())
        n_random += int((~m).sum()); n_random_assigned += int((ok & ~m & (res["barcode"] != -1) & (bc < 48)).sum())
        counts[:] += np.bincount(bc[ok & (res["barcode"] != -1)], minlength=49)[:49]
        assert np.all(np.abs(res["f_score"] - res["b_score"]) < 5e-2)
        return zlib.crc32(res["mapq"].tobytes() + res["barcode"].tobytes() + res["read_type"].tobytes())

    for k in range(nb):
        n = min(per, total - k * per)
        codes, lens, truth = synth.make_reads_fast(n, 150, tags, error_rate=0.01, random_frac=0.05, seed=1000 + k)
        b = batches[k % 2]
        b.clear(); b.append(codes, lens)
        gpu_ctx.submit(model, b, MODE_GET_LABEL, **kw)
        if pending is not None:
            crc = consume(gpu_ctx.wait(pending[0]), pending[1])
            if first_crc is None:
                first_crc = crc
        if k == 0:
            first_in = (codes, lens, truth)
        pending = (b, truth)
    crc = consume(gpu_ctx.wait(pending[0]), pending[1])
    if first_crc is None:
        first_crc = crc
    assert n_model + n_random == total
    assert n_model_ok / n_model > 0.995
    # a uniform-random 6-mer is within one substitution of one of the 48 tags with probability 48*19/4096 = 22 %
    assert 0.10 < n_random_assigned / n_random < 0.35
    exp = counts[:48].sum() / 48.0
    assert np.all(np.abs(counts[:48] - exp) < 0.02 * exp), counts
    # the first batch again, split differently (one third / two thirds): same per-read bits -> same checksum
    codes, lens, truth = first_in
    cut = (len(lens) // 3) // 32 * 32 + 7
    parts = []
    for lo, hi in ((0, cut), (cut, len(lens))):
        b = batches[0]
        b.clear(); b.append(codes[lo:hi], lens[lo:hi])
        gpu_ctx.submit(model, b, MODE_GET_LABEL, **kw)
        parts.append(gpu_ctx.wait(b))
    again = zlib.crc32(np.concatenate([p["mapq"] for p in parts]).tobytes() + np.concatenate([p["barcode"] for p in parts]).tobytes()
                       + np.concatenate([p["read_type"] for p in parts]).tobytes())
    assert again == first_crc
    for b in batches:
        b.close()
    model.close()


@pytest.mark.parametrize("bl,umi", [(9, 0), (10, 10), (11, 0), (12, 12), (14, 13), (16, 15)])
def test_long_barcodes_and_umis(gpu_ctx, oracle, ref, bl, umi):
    """Barcodes / UMIs of 9-16 nt run the unrolled standard-pattern kernels (kMaxStdCols): parity with the
    oracle and with the reference itself on ragged, N-containing reads."""
    from refharness import background_logp
    rng = np.random.default_rng(100 + bl)
    tags = []
    while len(tags) < 12:
        t = "".join(rng.choice(list("ACGT"), size=bl))
        if t not in tags:
            tags.append(t)
    segs = (["F:" + "N" * umi] if umi else []) + ["B:" + ",".join(tags), "R:N"]
    n = 1200
    codes, lens, truth = synth.make_reads(n, 70, tags, umi_len=umi, error_rate=0.02, random_frac=0.1, seed=bl, len_jitter=4, n_frac=0.01)
    p = ref.param_new(segs, threshold=1.0, minlen=16, dust=100, threads=4)
    mb = ref.model_new(p, background=background_logp((2501.0, 2480.0, 2510.0, 2492.0, 21.0)), average_length=70.0, max_seq_len=80)
    desc = ref.flatten(mb, p)
    kw = dict(threshold=1.0, minlen=16, dust=100)
    gpu = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_LABEL, **kw)
    ora = oracle.run(desc, MODE_GET_LABEL, codes, lens, threads=8, **kw)
    rep = compare(gpu, ora, lens, MODE_GET_LABEL, "long-barcode")
    assert all(v == 0 for v in rep.values()), rep
    want = ref.run_phmm(mb, p, 1, codes[:300], lens[:300])          # the reference's own run_pHMM(MODE_GET_LABEL)
    for k in ("mapq", "read_type", "barcode", "fingerprint"):
        assert np.array_equal(bits(gpu[k][:300]), bits(np.asarray(want[k]).astype(gpu[k].dtype))), k
    sel = (gpu["read_type"] == 0) & (truth >= 0)
    assert ((gpu["barcode"][sel] & 0xFFFF) == truth[sel]).mean() > 0.99
    ref.model_free(mb); ref.param_free(p)


@pytest.mark.parametrize("smem_state", [True, False])
@pytest.mark.parametrize("arch", ["long_partial", "long_linker"])
def test_column_loop_paths(gpu_ctx, oracle, ref, arch, smem_state, monkeypatch):
    """Segments longer than the unrolled kernels cover (P/O/G > 8 columns, standard pattern > 16) run the
    column-loop paths, with the profile state in shared memory when it fits and in thread-local arrays
    otherwise (TDG_NO_SMEM_STATE forces the latter): both against the oracle and the reference."""
    from refharness import background_logp
    if not smem_state:
        monkeypatch.setenv("TDG_NO_SMEM_STATE", "1")
    five, three = "AGGGAGGACGATGCGGTC", "GATCGGAAGAGCAC"
    tags = TAGS6_ED3[:8]
    if arch == "long_partial":
        segs = ["P:" + five, "B:" + ",".join(tags), "R:N", "P:" + three]
        kw_model = dict(five=(18.0, 15.2, 2.1), three=(14.0, 11.9, 1.7))
        gen = dict(linker5=five, linker3=three)
    else:
        segs = ["S:ACGTACGGTTCAGCATGCAAGGCTAACG", "B:" + ",".join(tags), "R:N"]
        kw_model = {}
        gen = dict(linker5="ACGTACGGTTCAGCATGCAAGGCTAACG")
    n = 1000
    codes, lens, truth = synth.make_reads(n, 90, tags, error_rate=0.02, random_frac=0.1, seed=77, len_jitter=5, n_frac=0.01, **gen)
    if arch == "long_partial":          # partial adapters: chop a random number of 5' bases off some reads
        rng = np.random.default_rng(3)
        for r in range(0, n, 3):
            k = int(rng.integers(1, 12))
            codes[r, : lens[r] - k] = codes[r, k: lens[r]].copy()
            lens[r] -= k
            codes[r, lens[r]:] = 0
    p = ref.param_new(segs, threshold=1.0, minlen=16, dust=100, threads=4)
    mb = ref.model_new(p, background=background_logp((2501.0, 2480.0, 2510.0, 2492.0, 21.0)), average_length=90.0, max_seq_len=100, **kw_model)
    desc = ref.flatten(mb, p)
    kw = dict(threshold=1.0, minlen=16, dust=100)
    gpu = run_gpu(gpu_ctx, desc, codes, lens, MODE_GET_LABEL, **kw)
    ora = oracle.run(desc, MODE_GET_LABEL, codes, lens, threads=8, **kw)
    rep = compare(gpu, ora, lens, MODE_GET_LABEL, arch)
    assert all(v == 0 for v in rep.values()), rep
    want = ref.run_phmm(mb, p, 1, codes[:250], lens[:250])
    for k in ("mapq", "read_type", "barcode", "fingerprint"):
        assert np.array_equal(bits(gpu[k][:250]), bits(np.asarray(want[k]).astype(gpu[k].dtype))), k
    ref.model_free(mb); ref.param_free(p)

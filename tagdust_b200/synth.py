"""Synthetic read generator for parity tests and bench.py.

Follows the reference's own generator (simulate_reads.c:142-322,480-547): each read is
[5' linker][UMI][barcode][read body][3' linker]; substitution errors at `error_rate` hit
only the linker/barcode/UMI part; the body is uniform random; the last `random_frac` of
the set are uniform-random reads of the same length ("contaminants").  Codes are the
reference's nucleotide codes A,C,G,T,N = 0..4 (nuc_code.c:46-74).
"""
import numpy as np

CODE = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3, "N": 4}

#: the 96 six-nt tags of the reference's dev/EDITTAG_6nt_ed_3.txt are shipped as a data
#: fixture (tests/golden/edittag_6nt_ed3.txt); see tests/golden/README.md.


def encode(seq: str) -> np.ndarray:
    return np.array([CODE.get(ch.upper(), 4) for ch in seq], dtype=np.uint8)


def load_tags(path, n=None):
    tags = []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if ":" in line and not line.startswith("["):
                tags.append(line.split(":")[1].strip())
    return tags if n is None else tags[:n]


def make_reads(n, read_len, barcodes, *, umi_len=0, linker5="", linker3="", error_rate=0.01,
               random_frac=0.05, seed=7, n_frac=0.0, len_jitter=0, second_barcodes=None):
    """Returns (codes[n, stride] uint8, lens[n] int32, truth[n] int32 (-1 = random read)).

    Layout: linker5 + UMI + barcode (+ second barcode) + body + linker3, total read_len
    (body absorbs the difference).  codes[r, len] is 0, like the reference's NUL terminator.
    """
    rng = np.random.default_rng(seed)
    stride = ((read_len + len_jitter + 1 + 15) // 16) * 16
    codes = np.zeros((n, stride), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.int32)
    truth = np.full(n, -1, dtype=np.int32)
    bcs = [encode(b) for b in barcodes]
    bcs2 = [encode(b) for b in second_barcodes] if second_barcodes else None
    l5, l3 = encode(linker5), encode(linker3)
    n_model = n - int(round(n * random_frac))
    for r in range(n):
        L = read_len + (int(rng.integers(-len_jitter, len_jitter + 1)) if len_jitter else 0)
        lens[r] = L
        if r >= n_model:
            codes[r, :L] = rng.integers(0, 4, size=L)
            continue
        k = int(rng.integers(0, len(bcs)))
        truth[r] = k
        parts = [l5]
        if umi_len:
            parts.append(rng.integers(0, 4, size=umi_len).astype(np.uint8))
        parts.append(bcs[k])
        if bcs2:
            k2 = int(rng.integers(0, len(bcs2)))
            truth[r] = k * len(bcs2) + k2
            parts.append(bcs2[k2])
        head = np.concatenate(parts).astype(np.uint8)
        if error_rate > 0:
            hit = rng.random(head.shape[0]) < error_rate
            head = np.where(hit, (head + rng.integers(1, 4, size=head.shape[0])) % 4, head).astype(np.uint8)
        body_len = L - head.shape[0] - l3.shape[0]
        if body_len < 0:
            raise ValueError("read_len too short for the architecture")
        body = rng.integers(0, 4, size=body_len).astype(np.uint8)
        seq = np.concatenate([head, body, l3]).astype(np.uint8)
        if n_frac > 0:
            seq = np.where(rng.random(L) < n_frac, 4, seq).astype(np.uint8)
        codes[r, :L] = seq
    return codes, lens, truth


def make_reads_fast(n, read_len, barcodes, *, error_rate=0.01, random_frac=0.05, seed=7):
    """Vectorised generator for the big bench sets (barcode + body only, fixed length)."""
    rng = np.random.default_rng(seed)
    stride = ((read_len + 1 + 15) // 16) * 16
    codes = np.zeros((n, stride), dtype=np.uint8)
    codes[:, :read_len] = rng.integers(0, 4, size=(n, read_len), dtype=np.uint8)
    bcs = np.stack([encode(b) for b in barcodes])
    bl = bcs.shape[1]
    n_model = n - int(round(n * random_frac))
    k = rng.integers(0, len(barcodes), size=n_model)
    head = bcs[k]
    hit = rng.random(head.shape) < error_rate
    head = np.where(hit, (head + rng.integers(1, 4, size=head.shape)) % 4, head).astype(np.uint8)
    codes[:n_model, :bl] = head
    truth = np.full(n, -1, dtype=np.int32)
    truth[:n_model] = k
    lens = np.full(n, read_len, dtype=np.int32)
    return codes, lens, truth


# ---- the other BASELINE.json configurations, shared by bench.py and the tests -------------------------
LINKER12 = "ACGTTGCAGTCA"


def cfg3_workload(tags, n, seed=2):
    """cfg3, read 1: UMI (8 uniform nt) + 12-nt linker + barcode (95 tags: 1 + 2 + 96 + 1 = 100 HMMs, the reference's
    `float total_prob[100]`, barcode_hmm.c:4186) + 124 random nt; read 2 is a plain 150-nt read (R:N, never reaches the HMM)."""
    tags = list(tags)[:95]
    segs = ["F:NNNNNNNN", "S:" + LINKER12, "B:" + ",".join(tags), "R:N"]
    c, lens, truth = make_reads_fast(n, 150, [LINKER12 + t for t in tags], seed=seed)
    rng = np.random.default_rng(seed + 1)
    codes = np.zeros_like(c)
    codes[:, :8] = rng.integers(0, 4, size=(n, 8))
    codes[:, 8:150] = c[:, :142]
    return segs, tags, codes, lens, truth


def cfg4_workload(tags, n, seed=4):
    """cfg4, read 1: I7 (24 tags) + I5 (16 tags) + 138 random nt = 384 combinations, -1 B: -2 B: -3 R:N; read 2 R:N."""
    i7, i5 = list(tags)[:24], list(tags)[24:40]
    segs = ["B:" + ",".join(i7), "B:" + ",".join(i5), "R:N"]
    codes, lens, truth = make_reads_fast(n, 150, [a + b for a in i7 for b in i5], seed=seed)
    return segs, (i7, i5), codes, lens, truth


def candidate_architectures(tags, n_arch):
    """cfg5: distinct candidate architectures for library-prep detection (test_architectures.c): barcode sets of different
    sizes, with/without UMI, linker, optional G, partial adapters.  Index 5 is the cfg2 architecture (the true one)."""
    out = []
    sizes = [4, 8, 12, 16, 24, 32, 48, 64]
    k = 0
    while len(out) < n_arch:
        nb = sizes[k % len(sizes)]
        variant = (k // len(sizes)) % 8
        t = tags[(k * 3) % 16:(k * 3) % 16 + nb]
        b = "B:" + ",".join(t)
        segs = {
            0: [b, "R:N"],
            1: ["F:NNNN", b, "R:N"],
            2: [b, "S:GGG", "R:N"],
            3: ["O:N", b, "R:N"],
            4: ["F:NNNNNNNN", "S:" + LINKER12, b, "R:N"],
            5: ["S:" + LINKER12[:6], b, "R:N"],
            6: [b, "R:N", "S:TTTTTT"],
            7: ["G:G", b, "S:T", "R:N"],
        }[variant]
        out.append(segs)
        k += 1
    if n_arch > 5:
        out[5] = ["B:" + ",".join(tags[:48]), "R:N"]
    assert len({tuple(x) for x in out}) == n_arch
    return out

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

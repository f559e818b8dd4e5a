#!/usr/bin/env python
"""Generate tests/golden/vectors_<case>.npz from the UNMODIFIED reference (oracle/_ref).

For every architecture of tests/cases.py: seeded synthetic reads, the flattened model the
reference built (init_model_bag), and the reference's per-read outputs of
backward()/forward_max_posterior_decoding() and run_pHMM(MODE_GET_LABEL).
Run on a box where /root/reference exists (oracle/_ref built by `make -C oracle ref`)."""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import CASES, build_ref_model, make_case_reads  # noqa: E402
from refharness import RefHarness  # noqa: E402

N_READS = 300
THRESHOLD = 1.5


def main():
    R = RefHarness()
    only = set(sys.argv[1:])          # optional: regenerate just the named cases
    for name in CASES:
        if only and name not in only:
            continue
        codes, lens, truth = make_case_reads(name, N_READS, seed=23)
        p, mb, desc = build_ref_model(R, name, threshold=THRESHOLD)
        sc = R.decode_scores(mb, codes, lens)
        run = R.run_phmm(mb, p, 1, codes, lens)
        out = dict(codes=codes, lens=lens, truth=truth, threshold=np.float32(THRESHOLD),
                   f_score=sc["f_score"], b_score=sc["b_score"], r_score=sc["r_score"],
                   bar_prob=sc["bar_prob"].astype(np.float32), labels=sc["labels"], mapq=run["mapq"],
                   read_type=run["read_type"], barcode=run["barcode"], fingerprint=run["fingerprint"],
                   seq_out=run["seq"], len_out=run["len"],
                   seg_type=np.frombuffer(desc.seg_type, dtype=np.uint8), average_raw_length=np.int32(desc.average_raw_length))
        for f in desc.FIELDS:
            out["model_" + f] = getattr(desc, f)
        np.savez_compressed(os.path.join(HERE, f"vectors_{name}.npz"), **out)
        print(name, "read_type counts", np.bincount(run["read_type"], minlength=7).tolist())
        R.model_free(mb); R.param_free(p)
    ref_dev = "/root/reference/dev"
    if os.path.isdir(ref_dev):
        for src, dst in (("EDITTAG_6nt_ed_3.txt", "edittag_6nt_ed3.txt"), ("EDITTAG_6nt_ed_4.txt", "edittag_6nt_ed4.txt")):
            shutil.copy(os.path.join(ref_dev, src), os.path.join(HERE, dst))


if __name__ == "__main__":
    main()

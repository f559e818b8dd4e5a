"""N>1 host logic on CPU: world_size-2 gloo run of the rank plumbing bench.py uses."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, %r)
    sys.path.insert(0, os.path.join(%r, "tests"))
    import numpy as np
    from tagdust_b200 import dist_util, synth
    from tagdust_b200._capi import MODE_GET_LABEL
    from tagdust_b200.api import compile_architecture
    from refharness import Oracle, background_logp
    from cases import TAGS6_ED4
    rank, world, local = dist_util.init("gloo")
    tags = TAGS6_ED4[:4]
    desc = compile_architecture(["B:" + ",".join(tags), "R:N"], background_logp(), 30.0, 30)
    codes, lens, truth = synth.make_reads(200, 30, tags, seed=1)          # same global set on every rank
    s, e = dist_util.rank_slice(len(lens), rank, world)
    out = Oracle().run(desc, MODE_GET_LABEL, codes[s:e], lens[s:e], threshold=1.5, threads=1)  # stand-in for the GPU shard
    tall = dist_util.merge_tallies(np.bincount(out["read_type"], minlength=7))
    per_bar = dist_util.merge_tallies(np.bincount((out["barcode"][out["read_type"] == 0] & 0xFFFF), minlength=5))
    dist_util.barrier()
    dist_util.cpu_barrier()          # the host-side barrier bench.py uses around its rank-0-only strong-scaling leg
    mx = dist_util.max_over_ranks([float(rank + 1), 10.0 - rank])
    if rank == 0:
        full = Oracle().run(desc, MODE_GET_LABEL, codes, lens, threshold=1.5, threads=1)
        print(json.dumps({"tall": tall.tolist(), "per_bar": per_bar.tolist(), "mx": mx,
                          "want_tall": np.bincount(full["read_type"], minlength=7).tolist(),
                          "want_bar": np.bincount((full["barcode"][full["read_type"] == 0] & 0xFFFF), minlength=5).tolist(),
                          "slice": [s, e]}))
    dist_util.finalize()
""") % (ROOT, ROOT)


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = [x for x in r.stdout.splitlines() if x.startswith("{")][-1]
    d = json.loads(line)
    assert d["tall"] == d["want_tall"] and d["per_bar"] == d["want_bar"]
    assert d["mx"] == [2.0, 10.0]
    assert d["slice"] == [0, 100]


def test_rank_slice_matches_reference_rule():
    from tagdust_b200.dist_util import rank_slice
    for n, w in ((10, 3), (1000001, 8), (7, 8), (0, 2)):
        got = [rank_slice(n, r, w) for r in range(w)]
        assert got[0][0] == 0 and got[-1][1] == n
        for a, b in zip(got, got[1:]):
            assert a[1] == b[0]

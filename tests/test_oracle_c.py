"""CPU: the C oracle (oracle.c) against the Python oracle, which is pinned to the real reference."""
import numpy as np
import pytest

from helpers import po, synth_small, rows_to_tuples


@pytest.mark.parametrize("scen", ["mixed_small", "ont_like", "deep_underflow", "amplicon_like", "maxdepth"])
def test_c_oracle_equals_python_oracle(lib, golden_synth, scen):
    from lvc_b200 import packing
    from oracle.c_oracle import COracle
    g = golden_synth[scen]
    reads = synth_small.rows_to_reads(g["reads"])
    for tname, res in g["results"].items():
        th = res["thresholds"]
        oc = po.OracleCaller(g["ref"], th["minBQ"], th["minMQ"], th["minDP"], th["minAD"], th["ratio"])
        oc.process_reads(reads)
        co = COracle(g["ref"], th["minBQ"], th["minMQ"], with_hist=True)
        co.process(packing.pack_reads(rows_to_tuples(g["reads"]), th["minMQ"]))
        L, S, emit, n = co.genotype(th["minDP"], th["minAD"], th["ratio"])
        assert sorted(np.nonzero(co.cov)[0].tolist()) == sorted(oc.memory.keys())
        lik = oc.likelihoods()
        for p, site in oc.memory.items():
            assert co.depth[p] == site["totalDepth"]
            for a, quals in site["snvs"].items():
                c = po.CHAR_TO_NIBBLE[a]
                assert co.ad[p, c] == len(quals) and co.qsum[p, c] == sum(quals)
                assert L[p, c] == lik[p][a]                      # bit-exact: same multiplication order
                for q in set(quals):
                    assert co.hist[p, c, q] == quals.count(q)
            assert int(co.ad[p].sum()) == sum(len(q) for q in site["snvs"].values())
        want = oc.prepare_variants()
        assert n == len(want)
        assert sorted((v["start"], po.CHAR_TO_NIBBLE[v["alleles"][1]]) for v in want) == \
            sorted(zip(*[x.tolist() for x in np.nonzero(emit)]))

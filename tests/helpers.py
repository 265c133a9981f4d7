"""Shared helpers for the parity tests (test infrastructure)."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "covid-spings-variant-caller_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import pileup_oracle as po  # noqa: E402
from oracle import synth_small  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DENORMAL_MIN_NORMAL = 2.2250738585072014e-308


def rows_to_tuples(rows):
    """golden read rows -> packer tuples (flag, pos, mapq, ops, seq, qual)."""
    out = []
    for name, flag, pos, mapq, cig, seq, qual in rows:
        out.append((flag, pos, mapq, po.parse_cigar(cig), seq, [ord(c) - 33 for c in qual]))
    return out


def golden_variant_key(v):
    return (v["start"], v["alleles"][1])


def variants_from_golden(gv):
    out = []
    for v in gv:
        gl = v["GL"] if isinstance(v["GL"], int) else float.fromhex(v["GL"]["hex"])
        out.append({"start": v["start"], "stop": v["stop"], "alleles": tuple(v["alleles"]),
                    "qual": float.fromhex(v["qual_hex"]),
                    "info": {"DP": v["DP"], "AD": v["AD"], "GL": gl, "PL": v["PL"], "SCORE": v["SCORE"]}})
    return out


def close_lik(a: float, b: float, n_factors: int = 1000, rel: float = 1e-9) -> bool:
    """likelihood comparison: 1e-9 relative in the normal range; inside the denormal band the reference's
    sequential product is itself order dependent -> absolute tolerance of a few denormal ulps
    per rounding step."""
    if a == b:
        return True
    if max(abs(a), abs(b)) < DENORMAL_MIN_NORMAL * 4:
        return abs(a - b) <= n_factors * 5e-324 + rel * max(abs(a), abs(b))
    return abs(a - b) <= rel * max(abs(a), abs(b))


def assert_variants_equal(got, want, what=""):
    """Emitted records must be identical: same order, same integer fields, same text for the floats."""
    assert len(got) == len(want), f"{what}: {len(got)} records, expected {len(want)}"
    for g, w in zip(got, want):
        assert g["start"] == w["start"] and g["stop"] == w["stop"], (what, g, w)
        assert tuple(g["alleles"]) == tuple(w["alleles"]), (what, g, w)
        for k in ("DP", "AD", "PL", "SCORE"):
            assert g["info"][k] == w["info"][k], (what, k, g, w)
        if isinstance(w["info"]["GL"], int):
            assert g["info"]["GL"] == w["info"]["GL"] and isinstance(g["info"]["GL"], int), (what, g, w)
        else:
            assert math.isclose(g["info"]["GL"], w["info"]["GL"], rel_tol=1e-9, abs_tol=1e-9), (what, g, w)
        assert math.isclose(float(g["qual"]), float(w["qual"]), rel_tol=1e-12), (what, g, w)


def memory_tables(mem):
    """memory dict (golden JSON or live) -> ({pos: depth}, {(pos, base, q): n}, {pos: [alleles in order]})"""
    depth, hist, order = {}, {}, {}
    for p, s in mem.items():
        p = int(p)
        depth[p] = s["totalDepth"]
        order[p] = list(s["snvs"].keys())
        for b, quals in s["snvs"].items():
            for q in quals:
                hist[(p, b, int(q))] = hist.get((p, b, int(q)), 0) + 1
    return depth, hist, order

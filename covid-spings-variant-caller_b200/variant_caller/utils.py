"""Scalar helpers with the reference's names and meaning (variant_caller/utils.py:9-24).

The device evaluates the same quantities from integer histograms (csrc/genotype.cuh); these host
versions serve callers that import them and the finalisation of emitted records."""
import math
from typing import Dict, List

from lvc_b200.records import to_phred_scale  # noqa: F401  (utils.py:12-13)


def from_phred_scale(score: float) -> float:
    """Phred -> error probability, host libm pow (utils.py:9-10)."""
    return math.pow(10, score / -10)


def genotype_likelihood(hypothesis: str, alleles: Dict[str, List[float]]):
    """Haploid Li-2011 likelihood of `hypothesis` (utils.py:16-24): product of (1 - e) over the reads that
    show it times the product of e over all other reads, left-to-right in fp64."""
    value = 1.0
    for e in alleles[hypothesis]:
        value *= (1.0 - e)
    rest = 1.0
    for allele, errors in alleles.items():
        if allele == hypothesis:
            continue
        part = 1.0
        for e in errors:
            part *= e
        rest *= part
    return value * rest

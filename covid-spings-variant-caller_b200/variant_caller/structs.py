"""Type hints of the persistent state and of the emitted records (mirrors variant_caller/structs.py:2-14
of the reference; runtime objects are plain dicts there too)."""
from typing import Dict, List, Tuple, TypedDict


class Site(TypedDict):
    reference: str                  # FASTA character at the position, case preserved
    totalDepth: int                 # pileup entries that passed the base-quality rule (incl. deletions)
    snvs: Dict[str, List[int]]      # allele letter -> phred qualities
    indels: Dict[str, List[int]]    # always {} (process_indel is disabled in the reference, :94)


class Variant(TypedDict):
    start: int
    stop: int
    alleles: Tuple[str, str]
    qual: float
    info: Dict

"""B200-native hot path of the COVID-SpiNGS live variant caller (pileup -> count tables -> Li-2011 GL).

Layout:
  capi.py      ctypes binding of liblvc_b200.so (include/lvc.h); no CPU fallback
  packing.py   structure-of-arrays batch packer (host)
  samio.py     SAM / BAM / FASTA readers, BAM writer (host ingest without pysam)
  records.py   host finalisation of device candidates into the reference's Variant dicts + VCF text
  synth.py     deterministic synthetic workloads of BASELINE.json's configs
  dist.py      multi-GPU sharding (samples per GPU; read chunks + NCCL table reduce)
"""
__all__ = ["capi", "packing", "samio", "records"]

"""ctypes wrapper of oracle/liboracle.so (oracle.c) -- test infrastructure, NOT product code."""
import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


class _State(C.Structure):
    _fields_ = [("G", C.c_int64), ("depth", C.c_void_p), ("cov", C.c_void_p), ("ad", C.c_void_p),
                ("qsum", C.c_void_p), ("q2sum", C.c_void_p), ("first", C.c_void_p), ("pe", C.c_void_p),
                ("p1", C.c_void_p), ("esum", C.c_void_p), ("hist", C.c_void_p), ("ordinal", C.c_uint64),
                ("e_lut", C.c_double * 256)]


def _load():
    src = os.path.join(HERE, "oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, src, "-lm"], check=True)
    lib = C.CDLL(LIB)
    lib.orc_process.restype = C.c_int
    lib.orc_process.argtypes = [C.POINTER(_State), C.c_uint32] + [C.c_void_p] * 9 + [C.c_int, C.c_int, C.c_int]
    lib.orc_genotype.restype = C.c_int64
    lib.orc_genotype.argtypes = [C.POINTER(_State), C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_void_p,
                                 C.c_void_p, C.c_void_p]
    lib.orc_admit.restype = C.c_int
    lib.orc_admit.argtypes = [C.c_uint32] + [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_int64, C.c_void_p]
    assert lib.orc_sizeof_state() == C.sizeof(_State)
    return lib


class COracle:
    """Reference-order streaming restatement over packed batches (any object with the ReadBatch fields)."""

    def __init__(self, ref: str, min_bq: int, min_mq: int, max_depth: int = 8000, with_hist: bool = False):
        self.lib = _load()
        self.G = len(ref)
        self.ref = np.frombuffer(ref.encode("latin-1"), dtype=np.uint8).copy()
        self.min_bq, self.min_mq, self.max_depth = min_bq, min_mq, max_depth
        G = self.G
        self.depth = np.zeros(G, np.uint32)
        self.cov = np.zeros(G, np.uint32)
        self.ad = np.zeros((G, 16), np.uint32)
        self.qsum = np.zeros((G, 16), np.uint64)
        self.q2sum = np.zeros((G, 16), np.uint64)
        self.first = np.full((G, 16), 0xFFFFFFFF, np.uint32)
        self.pe = np.ones((G, 16), np.float64)
        self.p1 = np.ones((G, 16), np.float64)
        self.esum = np.zeros((G, 16), np.float64)
        self.hist = np.zeros((G, 16, 256), np.uint32) if with_hist else None
        st = _State()
        st.G = G
        for name in ("depth", "cov", "ad", "qsum", "q2sum", "first", "pe", "p1", "esum"):
            setattr(st, name, getattr(self, name).ctypes.data)
        st.hist = self.hist.ctypes.data if with_hist else None
        st.ordinal = 0
        for q in range(256):
            st.e_lut[q] = math.pow(10, q / -10)
        self.st = st

    def process(self, b):
        rc = self.lib.orc_process(C.byref(self.st), b.n_reads, b.pos.ctypes.data, b.flag.ctypes.data,
                                  b.mapq.ctypes.data, None, b.cigar_off.ctypes.data, b.cigar.ctypes.data,
                                  b.seq_off.ctypes.data, b.seq4.ctypes.data, b.qual.ctypes.data, self.min_bq,
                                  self.min_mq, self.max_depth)
        if rc:
            raise RuntimeError(f"orc_process failed: {rc}")

    def genotype(self, min_dp: int, min_ad: int, ratio: float):
        L = np.zeros((self.G, 16), np.float64)
        S = np.zeros(self.G, np.float64)
        emit = np.zeros((self.G, 16), np.uint8)
        n = self.lib.orc_genotype(C.byref(self.st), self.ref.ctypes.data, int(min_dp), int(min_ad), float(ratio),
                                  L.ctypes.data, S.ctypes.data, emit.ctypes.data)
        return L, S, emit, int(n)

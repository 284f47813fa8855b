"""ctypes front-end to the CPU oracle (oracle/nw_oracle.c) and, when built,
the compiled reference harness (oracle/_ref/libnwref.so).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, from
__graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference
legs -- as the checker or the CPU baseline, never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_SRC = "/root/reference"

DNA = "ACGT"
PROTEIN = "ARNDCQEGHILKMFPSTWYV"

DIAG, LEFT, UP, MATCH = 1, 2, 4, 8


class _Result(C.Structure):
    _fields_ = [
        ("final_score", C.c_int32),
        ("branch_count", C.c_uint32),
        ("greatest_abs", C.c_int32),
        ("pad0", C.c_uint32),
        ("table_hash", C.c_uint64),
        ("arrow_hash", C.c_uint64),
        ("count", C.c_uint64),
        ("count_hash", C.c_uint64),
        ("lastrow_count_hash", C.c_uint64),
        ("lastcol_count_hash", C.c_uint64),
        ("arrow_digest", C.c_uint64),
        ("lastrow_count_digest", C.c_uint64),
        ("lastcol_count_digest", C.c_uint64),
    ]


@dataclass
class FillResult:
    final_score: int
    branch_count: int
    greatest_abs: int
    table_hash: int
    arrow_hash: int
    count: int
    count_hash: int
    lastrow_count_hash: int
    lastcol_count_hash: int
    arrow_digest: int = 0          # order-independent digests (nw_oracle.h): what the GPU computes on device
    lastrow_count_digest: int = 0
    lastcol_count_digest: int = 0
    scores: np.ndarray | None = None   # (B+1, A+1) int32
    codes: np.ndarray | None = None    # (B+1, A+1) uint8
    counts: np.ndarray | None = None   # (B+1, A+1) uint64
    packed: np.ndarray | None = None   # (B, pitch) uint8, include/nwb.h layout
    fill_seconds: float | None = None
    total_seconds: float | None = None


def build(with_reference: bool = True) -> None:
    """Compile the oracle and (if /root/reference exists) oracle/_ref."""
    target = "all" if with_reference else "oracle"
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libnw_oracle.so")
        if not os.path.exists(path):
            build(with_reference=False)
        _lib = C.CDLL(path)
        _lib.nwo_fill.restype = C.c_int
        _lib.nwo_fill.argtypes = [
            C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(_Result)]
        _lib.nwo_generate.restype = None
        _lib.nwo_generate.argtypes = [C.c_uint64, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        _lib.nwo_enumerate.restype = C.c_uint64
        _lib.nwo_enumerate.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int,
                                       C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    return _lib


def have_reference() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libnwref.so"))


def reference_cli() -> str | None:
    p = os.path.join(REF_DIR, "needleman-wunsch")
    return p if os.path.exists(p) else None


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        path = os.path.join(REF_DIR, "libnwref.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref/libnwref.so is not built (needs /root/reference; run oracle.build())")
        _ref = C.CDLL(path)
        _ref.nwref_fill.restype = C.c_int
        _ref.nwref_fill.argtypes = [
            C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_void_p, C.c_void_p, C.POINTER(_Result), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _ref.nwref_sizeof_score_cell.restype = C.c_int
        _ref.nwref_sizeof_walk_cell.restype = C.c_int
    return _ref


def _b(s) -> bytes:
    return s if isinstance(s, (bytes, bytearray)) else s.encode("latin-1")


def generate(seed: int, n: int, alphabet: str = DNA) -> bytes:
    """SURVEY.md 8d SplitMix64 generator."""
    buf = C.create_string_buffer(n + 1)
    lib().nwo_generate(C.c_uint64(seed & (2**64 - 1)), alphabet.encode(), len(alphabet), buf, n)
    return buf.raw[:n]


def generate_pair(seed: int, a: int, b: int, alphabet: str = DNA) -> tuple[bytes, bytes]:
    """top string from `seed`, side string from `seed + 1`."""
    return generate(seed, a, alphabet), generate(seed + 1, b, alphabet)


def packed_pitch(a: int) -> int:
    """Row pitch in bytes of the include/nwb.h nibble table: 32 cells = 16 B granules."""
    return max(16, ((a + 31) // 32) * 16)


def _mk(res: _Result, **kw) -> FillResult:
    return FillResult(res.final_score, res.branch_count, res.greatest_abs, res.table_hash,
                      res.arrow_hash, res.count, res.count_hash, res.lastrow_count_hash,
                      res.lastcol_count_hash, res.arrow_digest, res.lastrow_count_digest,
                      res.lastcol_count_digest, **kw)


def fill(top, side, m: int, k: int, d: int, *, want_scores=False, want_codes=False,
         want_counts=False, want_packed=False, pitch: int | None = None) -> FillResult:
    top, side = _b(top), _b(side)
    a, b = len(top), len(side)
    scores = np.empty((b + 1, a + 1), np.int32) if want_scores else None
    codes = np.empty((b + 1, a + 1), np.uint8) if want_codes else None
    counts = np.empty((b + 1, a + 1), np.uint64) if want_counts else None
    if want_packed:
        pitch = pitch or packed_pitch(a)
        packed = np.zeros((b, pitch), np.uint8)
    else:
        packed, pitch = None, 0
    res = _Result()
    ptr = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)
    rc = lib().nwo_fill(top, a, side, b, m, k, d, ptr(scores), ptr(codes), ptr(counts),
                        ptr(packed), pitch, C.byref(res))
    if rc != 0:
        raise RuntimeError("nwo_fill failed")
    return _mk(res, scores=scores, codes=codes, counts=counts, packed=packed)


def reference_fill(top, side, m: int, k: int, d: int, *, threads: int = 1, tflag: bool = True,
                   enumerate_count: bool = False, want_scores=False, want_codes=False) -> FillResult:
    """Run the compiled, unmodified reference (oracle/_ref) and read its tables."""
    top, side = _b(top), _b(side)
    if b"\0" in top or b"\0" in side:
        raise ValueError("reference strings are NUL-terminated")
    a, b = len(top), len(side)
    scores = np.empty((b + 1, a + 1), np.int32) if want_scores else None
    codes = np.empty((b + 1, a + 1), np.uint8) if want_codes else None
    res = _Result()
    fs, ts = C.c_double(), C.c_double()
    ptr = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)
    rc = ref().nwref_fill(top, side, m, k, d, threads, int(tflag), int(enumerate_count),
                          ptr(scores), ptr(codes), C.byref(res), C.byref(fs), C.byref(ts))
    if rc != 0:
        raise RuntimeError("nwref_fill failed")
    return _mk(res, scores=scores, codes=codes, fill_seconds=fs.value, total_seconds=ts.value)


def enumerate_alignments(top, side, codes: np.ndarray, limit: int = 0) -> list[tuple[bytes, bytes]]:
    """All optimal alignments in the reference's order (diag, left, up)."""
    top, side = _b(top), _b(side)
    out: list[tuple[bytes, bytes]] = []
    CB = C.CFUNCTYPE(None, C.POINTER(C.c_char), C.POINTER(C.c_char), C.c_int, C.c_void_p)

    def cb(x, y, n, _u):
        out.append((C.string_at(x, n), C.string_at(y, n)))

    cbf = CB(cb)
    codes = np.ascontiguousarray(codes, np.uint8)
    lib().nwo_enumerate(top, len(top), side, len(side), codes.ctypes.data_as(C.c_void_p),
                        limit, C.cast(cbf, C.c_void_p), None)
    return out

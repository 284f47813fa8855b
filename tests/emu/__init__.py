"""ctypes driver for the test-only SIMT emulator build of the kernels
(tests/emu/libnwb_emu.so).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class _Out(C.Structure):
    _fields_ = [("opt_score", C.c_int), ("branch_count", C.c_uint), ("greatest_abs", C.c_int),
                ("pad", C.c_int), ("count", C.c_ulonglong), ("pitch", C.c_ulonglong),
                ("spitch", C.c_ulonglong), ("dig_row", C.c_ulonglong), ("dig_col", C.c_ulonglong)]


_lib = None


def build() -> None:
    subprocess.run(["make", "-s", "-C", HERE], check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(os.path.join(HERE, "libnwb_emu.so"))
        _lib.emu_pitch_pk.restype = C.c_size_t
        _lib.emu_pitch_pk.argtypes = [C.c_int, C.c_int, C.c_int]
        for name in ("emu_pitch_i32", "emu_spitch_i32"):
            f = getattr(_lib, name)
            f.restype = C.c_size_t
            f.argtypes = [C.c_int, C.c_int]
    return _lib


def _b(s):
    return s if isinstance(s, (bytes, bytearray)) else s.encode("latin-1")


def fill_i32(top, side, m, k, d, *, flags=0, grid=2, split=0):
    """Run nwb_fill_i32_kernel under the emulator.  Returns dict of outputs."""
    top, side = _b(top), _b(side)
    a, b = len(top), len(side)
    L = lib()
    pitch = L.emu_pitch_i32(a, b)
    spitch = L.emu_spitch_i32(a, b)
    arrows = np.full((b, pitch), 0xEE, np.uint8)
    scores = np.zeros((b, spitch), np.int32) if flags & 1 else None
    cntmat = np.zeros((b, spitch), np.uint64) if flags & 0x20 else None
    out = _Out()
    ptr = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)
    L.emu_fill_i32.restype = C.c_int
    L.emu_fill_i32.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_uint, C.c_uint, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.POINTER(_Out)]
    rc = L.emu_fill_i32(top, a, side, b, m, k, d, flags, grid, split, ptr(arrows), ptr(scores),
                        ptr(cntmat), C.byref(out))
    assert rc == 0
    return dict(opt_score=out.opt_score, branch_count=out.branch_count, greatest_abs=out.greatest_abs,
                count=out.count, arrows=arrows, scores=scores, cntmat=cntmat, pitch=pitch)


def pk_supported(m, k, d) -> bool:
    return bool(lib().emu_pk_supported(m, k, d))


def hx_supported(m, k, d) -> bool:
    return bool(lib().emu_hx_supported(m, k, d))


def fill_pk(top, side, m, k, d, *, K=4, R=1, grid=2, split=0, warps=4, count=False, hx=False):
    """Run nwb_fill_pk_kernel<K> (+ the branch-count pass) under the emulator."""
    top, side = _b(top), _b(side)
    a, b = len(top), len(side)
    L = lib()
    pitch = L.emu_pitch_pk(a, b, K)
    arrows = np.full((b, pitch), 0xEE, np.uint8)
    out = _Out()
    L.emu_fill_pk.restype = C.c_int
    L.emu_fill_pk.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_int, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(_Out)]
    # count: False / True (fused into the fill) / 2, 4, 8 (second sweep over the arrow codes with that many
    # cells per lane, csrc/nwb_count.cuh)
    rc = L.emu_fill_pk(top, a, side, b, m, k, d, K, R, int(count), grid, warps, split, int(hx),
                       arrows.ctypes.data_as(C.c_void_p), C.byref(out))
    assert rc == 0, rc
    return dict(opt_score=out.opt_score, branch_count=out.branch_count, arrows=arrows, pitch=pitch, count=out.count,
                dig_row=out.dig_row, dig_col=out.dig_col)


def sparse_count(packed: np.ndarray, a: int, mode: int = 0, min_col: int = 0) -> dict:
    """nwb_sparse_count_kernel (csrc/nwb_count_sparse.cuh) over a (B, pitch) nibble table in the include/nwb.h
    layout.  state: 1 = done (count is final), 2 = gave up (the dense sweep would run).  min_col > 1: the table holds
    columns min_col .. a only (the last rank of a strip group); flow that leaves to the left makes the sweep give up."""
    packed = np.ascontiguousarray(packed, np.uint8)
    b, pitch = packed.shape
    assert pitch % 4 == 0 and pitch * 2 >= a
    res = (C.c_ulonglong * 3)()
    L = lib()
    L.emu_sparse_count.restype = C.c_int
    L.emu_sparse_count.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ulonglong)]
    # mode: 0 = as the product (64-column window first, then 256), 1 = 256-column window only, 2 = 64-column only
    assert L.emu_sparse_count(packed.ctypes.data_as(C.c_void_p), pitch, a, b, mode, min_col, res) == 0
    return dict(count=int(res[0]), state=int(res[1]), rows=int(res[2]))


def arrow_digest(packed: np.ndarray, a: int, w_begin: int = 0, w_end: int | None = None, grid: int = 2) -> int:
    """nwb_arrow_digest_kernel (csrc/nwb_digest.cuh) over words [w_begin, w_end) of a (B, pitch) nibble table."""
    packed = np.ascontiguousarray(packed, np.uint8)
    b, pitch = packed.shape
    w_end = (a + 7) // 8 if w_end is None else w_end
    L = lib()
    L.emu_arrow_digest.restype = C.c_ulonglong
    L.emu_arrow_digest.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint]
    return int(L.emu_arrow_digest(packed.ctypes.data_as(C.c_void_p), pitch, a, b, w_begin, w_end, grid))


def bpitch_pk(a: int, b: int) -> int:
    L = lib()
    L.emu_bpitch_pk.restype = C.c_size_t
    L.emu_bpitch_pk.argtypes = [C.c_int, C.c_int]
    return int(L.emu_bpitch_pk(a, b))


def fill_pk_rank(top, side, m, k, d, *, rank, world, inbox=None, hx=True, grid=2):
    """One rank of a column-strip group (K = 4, R = 2) under the emulator.  inbox: uint32[bpitch] written by
    rank - 1 (None for rank 0).  Returns the rank's arrow table (only its own columns are written), its
    outbox for rank + 1, partial_r, branch count and strip range."""
    top, side = _b(top), _b(side)
    a, b = len(top), len(side)
    L = lib()
    pitch = L.emu_pitch_pk(a, b, 4)
    bp = bpitch_pk(a, b)
    arrows = np.zeros((b, pitch), np.uint8)
    outbox = np.zeros(bp, np.uint32)
    if inbox is not None:
        inbox = np.ascontiguousarray(inbox, np.uint32)
        assert inbox.shape == (bp,)
    pr = C.c_longlong()
    br = C.c_uint()
    info = (C.c_int * 4)()
    L.emu_fill_pk_rank.restype = C.c_int
    L.emu_fill_pk_rank.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int,
                                   C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_longlong),
                                   C.POINTER(C.c_uint), C.POINTER(C.c_int)]
    rc = L.emu_fill_pk_rank(top, a, side, b, m, k, d, grid, int(hx), rank, world,
                            None if inbox is None else inbox.ctypes.data_as(C.c_void_p),
                            outbox.ctypes.data_as(C.c_void_p), arrows.ctypes.data_as(C.c_void_p),
                            C.byref(pr), C.byref(br), info)
    assert rc == 0, rc
    return dict(arrows=arrows, outbox=outbox, partial_r=pr.value, branch_count=br.value,
                strip_begin=info[0], strip_end=info[1], n_strips=info[2], pitch=pitch)


def fill_batch(tops, sides, m, k, d, *, grid=1, bx=-1, count=False):
    """Run the batch kernel under the emulator.  bx: 0 = nwb_batch_pk_kernel (one pair per warp), 1 =
    nwb_batch_bx_kernel (two pairs per warp), 2 = nwb_batch_cx_kernel (uniform shapes, pairs back to back),
    -1 = the library's own choice; the result's "kernel" says which ran ("pk", "bx", "cx")."""
    n = len(tops)
    toff = np.zeros(n + 1, np.int64)
    soff = np.zeros(n + 1, np.int64)
    np.cumsum([len(t) for t in tops], out=toff[1:])
    np.cumsum([len(s) for s in sides], out=soff[1:])
    total = sum(max(1, (len(t) + 255) // 256) * 128 * len(s) for t, s in zip(tops, sides))
    arrows = np.full(total + 16, 0xEE, np.uint8)
    aoff = np.zeros(n + 1, np.int64)
    scores = np.zeros(n, np.int32)
    branches = np.zeros(n, np.uint32)
    L = lib()
    L.emu_fill_batch.restype = C.c_int
    L.emu_fill_batch.argtypes = [C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int,
                                 C.c_int, C.c_uint, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.POINTER(C.c_int)]
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    used = C.c_int(0)
    rc = L.emu_fill_batch(b"".join(tops), p(toff), b"".join(sides), p(soff), n, m, k, d, grid, bx, p(arrows), p(aoff),
                          p(scores), p(branches), C.byref(used))
    assert rc == 0, rc
    counts = None
    if count:
        counts = np.zeros(n, np.uint64)
        L.emu_batch_count.restype = C.c_int
        L.emu_batch_count.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
        assert L.emu_batch_count(p(toff), p(soff), n, grid, p(arrows), p(aoff), p(counts)) == 0
    tabs = []
    for i in range(n):
        pitch = max(1, (len(tops[i]) + 255) // 256) * 128
        tabs.append(arrows[aoff[i]:aoff[i] + pitch * len(sides[i])].reshape(len(sides[i]), pitch))
    return dict(scores=scores, branches=branches, tables=tabs, bx=used.value >= 1, kernel=("pk", "bx", "cx")[used.value],
                counts=counts)


def fill_batch_bp(tops, sides, m, k, d, *, grid=1, warps=2, words=0):
    """nwb_batch_bp_kernel (csrc/nwb_batch_bp.cuh: bit-parallel rows, one thread per pair) under the emulator,
    followed by nwb_batch_pk_kernel over the pairs it left over.  words: 0 = as the library picks them (2, 4 or 8 per row
    vector by the longest top string), or 2 / 4 / 8.  Returns None when the batch does not qualify."""
    build()
    L = lib()
    n = len(tops)
    tops = [_b(t) for t in tops]
    sides = [_b(s) for s in sides]
    toff = np.zeros(n + 1, dtype=np.int64)
    soff = np.zeros(n + 1, dtype=np.int64)
    toff[1:] = np.cumsum([len(t) for t in tops])
    soff[1:] = np.cumsum([len(s) for s in sides])
    aoff = np.zeros(n + 1, dtype=np.int64)
    total = int(sum(128 * len(s) for s in sides))
    arrows = np.full(total + 16, 0xEE, dtype=np.uint8)
    scores = np.zeros(n + 1, dtype=np.int32)
    branches = np.zeros(n + 1, dtype=np.uint32)
    nfb = C.c_longlong(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L.emu_fill_batch_bp.restype = C.c_int
    L.emu_fill_batch_bp.argtypes = [C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                    C.c_uint, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_longlong)]
    # the kernels read up to 3 bytes past a string (aligned word loads): pad like the product's device buffers
    rc = L.emu_fill_batch_bp(b"".join(tops) + b"\0" * 16, p(toff), b"".join(sides) + b"\0" * 16, p(soff), n, m, k, d, grid, warps, words,
                             p(arrows), p(aoff), p(scores), p(branches), C.byref(nfb))
    if rc == -6:
        return None
    assert rc == 0, rc
    tables = [arrows[int(aoff[i]):int(aoff[i + 1])].reshape(len(sides[i]), 128) for i in range(n)]
    return {"tables": tables, "scores": scores[:n].copy(), "branches": branches[:n].copy(), "n_fallback": int(nfb.value)}


def batch_lcount(tops, sides, tables, *, grid=1, warps=2):
    """nwb_batch_lcount_kernel (csrc/nwb_batch_lcount.cuh: one thread per pair, sparse backward sweep over arrow
    tables a batch fill has written) under the emulator, followed by the dense kernel over the pairs it gave up on.
    tables[i]: (B_i, pitch_i) uint8.  Returns (counts, n_fallback)."""
    build()
    L = lib()
    n = len(tops)
    toff = np.zeros(n + 1, dtype=np.int64)
    soff = np.zeros(n + 1, dtype=np.int64)
    toff[1:] = np.cumsum([len(t) for t in tops])
    soff[1:] = np.cumsum([len(s) for s in sides])
    aoff = np.zeros(n + 1, dtype=np.int64)
    aoff[1:] = np.cumsum([t.size for t in tables])
    arrows = np.concatenate([np.ascontiguousarray(t, np.uint8).ravel() for t in tables] + [np.full(64, 0xEE, np.uint8)])
    counts = np.zeros(n + 1, dtype=np.uint64)
    nfb = C.c_longlong(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L.emu_batch_lcount.restype = C.c_int
    L.emu_batch_lcount.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_uint, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.POINTER(C.c_longlong)]
    rc = L.emu_batch_lcount(p(toff), p(soff), n, grid, warps, p(arrows), p(aoff), p(counts), C.byref(nfb))
    assert rc == 0, rc
    return counts[:n].copy(), int(nfb.value)


def fill_batch_i32(tops, sides, m, k, d, *, grid=1, want_scores=False, want_abs=True):
    """nwb_batch_i32_kernel (csrc/nwb_batch_i32.cuh: any m / k / d, one warp per pair) under the emulator."""
    n = len(tops)
    toff = np.zeros(n + 1, np.int64)
    soff = np.zeros(n + 1, np.int64)
    np.cumsum([len(t) for t in tops], out=toff[1:])
    np.cumsum([len(s) for s in sides], out=soff[1:])
    total = sum(max(1, (len(t) + 255) // 256) * 128 * len(s) for t, s in zip(tops, sides))
    arrows = np.full(total + 16, 0xEE, np.uint8)
    aoff = np.zeros(n + 1, np.int64)
    scores = np.zeros(n, np.int32)
    branches = np.zeros(n, np.uint32)
    gabs = np.zeros(n, np.int32) if want_abs else None
    smat = np.zeros(2 * total + 16, np.int32) if want_scores else None
    scoff = np.zeros(n + 1, np.int64) if want_scores else None
    L = lib()
    L.emu_fill_batch_i32.restype = C.c_int
    L.emu_fill_batch_i32.argtypes = [C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                     C.c_uint] + [C.c_void_p] * 7
    p = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)
    rc = L.emu_fill_batch_i32(b"".join(tops), p(toff), b"".join(sides), p(soff), n, m, k, d, grid, p(arrows), p(aoff),
                              p(scores), p(branches), p(gabs), p(smat), p(scoff))
    assert rc == 0, rc
    tabs, mats = [], []
    for i in range(n):
        ns = max(1, (len(tops[i]) + 255) // 256)
        tabs.append(arrows[aoff[i]:aoff[i] + ns * 128 * len(sides[i])].reshape(len(sides[i]), ns * 128))
        if want_scores:
            mats.append(smat[scoff[i]:scoff[i] + ns * 256 * len(sides[i])].reshape(len(sides[i]), ns * 256)[:, :len(tops[i])])
    return dict(scores=scores, branches=branches, abs=gabs, tables=tabs, score_rows=mats)


def unpack_arrows(packed: np.ndarray, a: int) -> np.ndarray:
    """(B, pitch) nibble table -> (B, A) uint8 codes (DIAG|LEFT|UP)."""
    lo = packed & 0xF
    hi = packed >> 4
    inter = np.empty((packed.shape[0], packed.shape[1] * 2), np.uint8)
    inter[:, 0::2] = lo
    inter[:, 1::2] = hi
    return inter[:, :a]

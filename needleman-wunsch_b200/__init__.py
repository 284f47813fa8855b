"""needleman-wunsch_b200 -- B200-native Needleman-Wunsch score-table fill.

Python host mirror of the C ABI in include/nwb.h (ctypes over libnwb.so).  The
library is the product; this module only marshals arguments.  There is no CPU
fallback: if libnwb.so is missing or no CUDA device is present every call
raises.

The names follow the reference's operator interface for this path
(skotchandsoda/needleman-wunsch): `Computation` mirrors `computation_t`
(computation.h:45-75) and `Computation.compute_table_scores()` mirrors
`compute_table_scores(C)` (needleman-wunsch.c:583).

The directory name contains a hyphen; import it as
    import importlib; nwb = importlib.import_module("needleman-wunsch_b200")
or through the repo-root shim `import nw_b200`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnwb.so")

# arrow bits / flags / errors: include/nwb.h
DIAG, LEFT, UP, MATCH = 1, 2, 4, 8
WANT_SCORES = 0x01
WANT_COUNT = 0x02
WANT_ARROWS_HOST = 0x04
TRACK_ABS = 0x08
FORCE_GENERAL = 0x10
WANT_COUNT_MATRIX = 0x20
NO_BRANCH_COUNT = 0x40
WANT_COUNT_DIGEST = 0x80
WANT_DIGEST = 0x100
QUEUE = 0x200          # Plan(): one of several plans on a GPU working through a queue of fills (consecutive fills overlap)
KIND_I32, KIND_PK = 0, 1
# nwb_summary.count_path
COUNT_NONE, COUNT_FUSED, COUNT_DENSE, COUNT_SPARSE, COUNT_SPARSE_BAILED = 0, 1, 2, 3, 4


class NwbError(RuntimeError):
    def __init__(self, code: int, what: str):
        self.code = code
        detail = ""
        try:
            detail = _lib.nwb_last_cuda_error().decode() if _lib is not None else ""
            msg = _lib.nwb_strerror(code).decode() if _lib is not None else str(code)
        except Exception:  # pragma: no cover
            msg = str(code)
        super().__init__(f"{what}: {msg} ({code}){' -- ' + detail if detail else ''}")


class Summary(C.Structure):
    _fields_ = [("opt_score", C.c_int32), ("branch_count", C.c_uint32), ("greatest_abs", C.c_int32),
                ("kernel_kind", C.c_int32), ("count", C.c_uint64), ("partial_r", C.c_int64),
                ("count_path", C.c_int32), ("count_rows", C.c_uint32),
                ("lastrow_count_digest", C.c_uint64), ("lastcol_count_digest", C.c_uint64)]


_lib = None

_SIGS = {
    "nwb_strerror": (C.c_char_p, [C.c_int]),
    "nwb_last_cuda_error": (C.c_char_p, []),
    "nwb_device_count": (C.c_int, []),
    "nwb_abi_version": (C.c_int, []),
    "nwb_host_alloc": (C.c_void_p, [C.c_size_t]),
    "nwb_host_free": (None, [C.c_void_p]),
    "nwb_tune": (C.c_int, [C.c_char_p, C.c_int]),
    "nwb_tune_reset": (None, []),
    "nwb_table_summary": (C.c_int, [C.c_void_p, C.POINTER(Summary)]),
    "nwb_arrow_digest": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "nwb_plan_arrow_digest": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "nwb_plan_count_path_name": (C.c_char_p, [C.c_void_p]),
    "nwb_batch_digest": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_uint64)]),
    "nwb_fill": (C.c_int, [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint,
                           C.POINTER(C.c_void_p)]),
    "nwb_fill_on": (C.c_int, [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint,
                              C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "nwb_free": (None, [C.c_void_p]),
    "nwb_cache_clear": (None, []),
    "nwb_top_len": (C.c_int, [C.c_void_p]),
    "nwb_side_len": (C.c_int, [C.c_void_p]),
    "nwb_opt_score": (C.c_int32, [C.c_void_p]),
    "nwb_count_u64": (C.c_uint64, [C.c_void_p]),
    "nwb_branch_count": (C.c_uint32, [C.c_void_p]),
    "nwb_greatest_abs_interior": (C.c_int32, [C.c_void_p]),
    "nwb_score": (C.c_int32, [C.c_void_p, C.c_int, C.c_int]),
    "nwb_arrows": (C.c_uint, [C.c_void_p, C.c_int, C.c_int]),
    "nwb_arrow_rows": (C.c_void_p, [C.c_void_p]),
    "nwb_arrow_pitch": (C.c_size_t, [C.c_void_p]),
    "nwb_count_at": (C.c_uint64, [C.c_void_p, C.c_int, C.c_int]),
    "nwb_score_rows": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "nwb_count_rows": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "nwb_kernel_ms": (C.c_float, [C.c_void_p]),
    "nwb_kernel_kind": (C.c_int, [C.c_void_p]),
    "nwb_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "nwb_plan_destroy": (None, [C.c_void_p]),
    "nwb_plan_upload": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]),
    "nwb_plan_run": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nwb_plan_summary": (C.c_int, [C.c_void_p, C.POINTER(Summary)]),
    "nwb_plan_arrows_device": (C.c_void_p, [C.c_void_p]),
    "nwb_plan_arrow_pitch": (C.c_size_t, [C.c_void_p]),
    "nwb_plan_download_arrows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]),
    "nwb_plan_launches": (C.c_int64, [C.c_void_p]),
    "nwb_plan_kernel_name": (C.c_char_p, [C.c_void_p]),
    "nwb_plan_kernel_ms": (C.c_float, [C.c_void_p]),
    "nwb_plan_strip_range": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nwb_plan_reset_inbox": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nwb_plan_run_pipelined": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nwb_plan_attach_right": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nwb_plan_ipc_size": (C.c_size_t, []),
    "nwb_plan_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nwb_plan_ipc_attach_right": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nwb_measure_int_issue": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "nwb_fill_batch": (C.c_int, [C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                 C.c_int, C.c_uint, C.c_int, C.POINTER(C.c_void_p)]),
    "nwb_batch_create": (C.c_int, [C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                   C.c_int, C.c_uint, C.c_int, C.POINTER(C.c_void_p)]),
    "nwb_batch_run": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nwb_batch_refill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nwb_batch_fetch": (C.c_int, [C.c_void_p]),
    "nwb_batch_launches": (C.c_int64, [C.c_void_p]),
    "nwb_batch_kernel_name": (C.c_char_p, [C.c_void_p]),
    "nwb_batch_arrows_device": (C.c_void_p, [C.c_void_p]),
    "nwb_batch_free": (None, [C.c_void_p]),
    "nwb_batch_size": (C.c_int64, [C.c_void_p]),
    "nwb_batch_opt_score": (C.c_int32, [C.c_void_p, C.c_int64]),
    "nwb_batch_branch_count": (C.c_uint32, [C.c_void_p, C.c_int64]),
    "nwb_batch_count_u64": (C.c_uint64, [C.c_void_p, C.c_int64]),
    "nwb_batch_arrow_rows": (C.c_void_p, [C.c_void_p, C.c_int64, C.POINTER(C.c_size_t)]),
    "nwb_batch_kernel_ms": (C.c_float, [C.c_void_p]),
    "nwb_batch_greatest_abs": (C.c_int32, [C.c_void_p, C.c_int64]),
    "nwb_batch_score_rows": (C.c_void_p, [C.c_void_p, C.c_int64, C.POINTER(C.c_size_t)]),
    "nwb_strip_partition": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nwb_batch_partition": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "nwb_strip_group_score": (C.c_int32, [C.c_int64, C.c_int, C.c_int, C.c_int]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


def build(verbose: bool = False) -> str:
    """Compile csrc/ into libnwb.so for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building libnwb.so failed")
    return LIB_PATH


def use_experiments_build() -> None:
    """Measurement tools only: load libnwb_exp.so (`make -C needleman-wunsch_b200/csrc exp`, -DNWB_EXPERIMENTS: the
    measured-slower nwb_fill_hy.cuh geometry and the wait-skipping debug bits).  Call before anything else."""
    global LIB_PATH
    if _lib is not None:
        raise RuntimeError("libnwb.so is already loaded")
    LIB_PATH = os.path.join(HERE, "libnwb_exp.so")


def load_library() -> C.CDLL:
    """Load libnwb.so.  Raises if it was not built -- there is nothing to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C needleman-wunsch_b200/csrc` "
                               "(the fill has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(lib, name)
            f.restype = res
            f.argtypes = args
        _lib = lib
    return _lib


def device_count() -> int:
    return load_library().nwb_device_count()


def measure_int_issue(mode: int = 1, device: int = 0) -> tuple[float, float]:
    """(thread-instructions per clk per SM, G thread-instructions/s) of the INT/DPX microbenchmark."""
    a, b = C.c_double(), C.c_double()
    rc = load_library().nwb_measure_int_issue(device, mode, C.byref(a), C.byref(b))
    if rc != 0:
        raise NwbError(rc, "nwb_measure_int_issue")
    return a.value, b.value


def tune(key: str, value: int) -> None:
    """nwb_tune(): explicit kernel-selection / watchdog overrides for tests and measurements (include/nwb.h)."""
    rc = load_library().nwb_tune(key.encode(), int(value))
    if rc != 0:
        raise NwbError(rc, f"nwb_tune({key!r})")


class PinnedBuffer:
    """Page-locked host memory (nwb_host_alloc) holding a copy of `data`; pass it where host strings are expected."""

    def __init__(self, data: bytes):
        self.nbytes = len(data)
        self.ptr = load_library().nwb_host_alloc(max(1, self.nbytes))
        if not self.ptr:
            raise NwbError(-2, "nwb_host_alloc")
        C.memmove(self.ptr, data, self.nbytes)

    def close(self) -> None:
        if self.ptr:
            load_library().nwb_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _host_ptr(x):
    if isinstance(x, PinnedBuffer):
        return C.c_void_p(x.ptr)
    return C.cast(C.c_char_p(x), C.c_void_p)


def cache_clear() -> None:
    """Release the device workspace nwb_fill()/nwb_fill_on() keep between calls."""
    load_library().nwb_cache_clear()


def tune_reset() -> None:
    load_library().nwb_tune_reset()


class tuned:
    """Context manager: `with nwb.tuned(pk_hx=0, count_mode=2): ...` sets the overrides and restores the defaults."""

    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        for k, v in self.kw.items():
            tune(k, v)
        return self

    def __exit__(self, *exc):
        tune_reset()
        return False


def strip_partition(top_len: int, rank: int, world: int, strip_width: int = 256) -> tuple[int, int]:
    """Interior columns [begin, end) of rank `rank` in a column-strip group (host-only, include/nwb.h)."""
    b, e = C.c_int(), C.c_int()
    _ck(load_library().nwb_strip_partition(top_len, strip_width, rank, world, C.byref(b), C.byref(e)),
        "nwb_strip_partition")
    return b.value, e.value


def batch_partition(n_pairs: int, rank: int, world: int) -> tuple[int, int]:
    """(first pair, pair count) of rank `rank` when a batch is sharded across `world` GPUs (host-only)."""
    f, c = C.c_int64(), C.c_int64()
    _ck(load_library().nwb_batch_partition(n_pairs, rank, world, C.byref(f), C.byref(c)), "nwb_batch_partition")
    return f.value, c.value


def strip_group_score(partial_r_sum: int, top_len: int, side_len: int, d: int) -> int:
    """Optimal score of a strip group from the sum over ranks of Summary.partial_r (host-only)."""
    return int(load_library().nwb_strip_group_score(partial_r_sum, top_len, side_len, d))


DNA = "ACGT"
PROTEIN = "ARNDCQEGHILKMFPSTWYV"


def generate(seed: int, n: int, alphabet: str = DNA, count: int = 1, seed_stride: int = 0) -> bytes:
    """The benchmark's input generator: SplitMix64 seeded with `seed`, one character per draw,
    `alphabet[z % len(alphabet)]`.  count > 1 concatenates `count` strings of n characters with seeds
    seed, seed + seed_stride, ... (config 4: pair p has top seed 0x5EED4000 + 2p and side seed + 1).
    Vectorised with numpy; tests/test_oracle.py checks it against the oracle's scalar generator."""
    a = np.frombuffer(alphabet.encode("latin-1"), np.uint8)
    with np.errstate(over="ignore"):
        seeds = np.uint64(seed & (2**64 - 1)) + np.arange(count, dtype=np.uint64) * np.uint64(seed_stride & (2**64 - 1))
        i = np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        z = seeds[:, None] + i[None, :]
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
        idx = (z % np.uint64(len(a))).astype(np.intp)
    return a[idx].tobytes()


def generate_pair(seed: int, a: int, b: int, alphabet: str = DNA) -> tuple[bytes, bytes]:
    """top string from `seed`, side string from `seed + 1`."""
    return generate(seed, a, alphabet), generate(seed + 1, b, alphabet)


def _b(s) -> bytes:
    return bytes(s) if isinstance(s, (bytes, bytearray, memoryview)) else s.encode("latin-1")


def _ck(rc: int, what: str) -> None:
    if rc != 0:
        raise NwbError(rc, what)


class Table:
    """Result of one fill (nwb_table): the GPU-produced score/arrow table."""

    def __init__(self, handle: int, top: bytes, side: bytes):
        self._h = C.c_void_p(handle)
        self.top, self.side = top, side
        L = load_library()
        self.top_len = L.nwb_top_len(self._h)
        self.side_len = L.nwb_side_len(self._h)
        self.opt_score = L.nwb_opt_score(self._h)
        self.count = L.nwb_count_u64(self._h)
        self.branch_count = L.nwb_branch_count(self._h)
        self.greatest_abs = L.nwb_greatest_abs_interior(self._h)
        self.kernel_ms = L.nwb_kernel_ms(self._h)
        self.kernel_kind = L.nwb_kernel_kind(self._h)
        self.pitch = L.nwb_arrow_pitch(self._h)

    def summary(self) -> Summary:
        s = Summary()
        _ck(load_library().nwb_table_summary(self._h, C.byref(s)), "nwb_table_summary")
        return s

    def arrow_digest(self) -> int:
        """Digest of the whole arrow table, computed on the device (NWB_WANT_DIGEST)."""
        d = C.c_uint64()
        _ck(load_library().nwb_arrow_digest(self._h, C.byref(d)), "nwb_arrow_digest")
        return d.value

    def score(self, i: int, j: int) -> int:
        return load_library().nwb_score(self._h, i, j)

    def arrows(self, i: int, j: int) -> int:
        return load_library().nwb_arrows(self._h, i, j)

    def count_at(self, i: int, j: int) -> int:
        return load_library().nwb_count_at(self._h, i, j)

    def arrow_rows(self) -> np.ndarray | None:
        """(B, pitch) uint8 view of the host nibble table (copy)."""
        p = load_library().nwb_arrow_rows(self._h)
        if not p or self.side_len == 0:
            return None
        buf = (C.c_uint8 * (self.pitch * self.side_len)).from_address(p)
        return np.frombuffer(buf, np.uint8).reshape(self.side_len, self.pitch).copy()

    def arrow_codes(self) -> np.ndarray:
        """(B, A) uint8 DIAG|LEFT|UP codes of the interior cells."""
        rows = self.arrow_rows()
        if rows is None:
            raise RuntimeError("arrows are not on the host (NWB_WANT_ARROWS_HOST)")
        return unpack_arrows(rows, self.top_len)

    def _rows(self, fn, ctype, dtype):
        pitch = C.c_size_t()
        p = fn(self._h, C.byref(pitch))
        if not p or self.side_len == 0:
            return None
        buf = (ctype * (pitch.value * self.side_len)).from_address(p)
        return np.frombuffer(buf, dtype).reshape(self.side_len, pitch.value)[:, :self.top_len].copy()

    def score_rows(self) -> np.ndarray | None:
        """(B, A) int32 interior scores (NWB_WANT_SCORES)."""
        return self._rows(load_library().nwb_score_rows, C.c_int32, np.int32)

    def count_rows(self) -> np.ndarray | None:
        """(B, A) uint64 interior counts (NWB_WANT_COUNT_MATRIX)."""
        return self._rows(load_library().nwb_count_rows, C.c_uint64, np.uint64)

    def close(self) -> None:
        if self._h:
            load_library().nwb_free(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def unpack_arrows(packed: np.ndarray, a: int) -> np.ndarray:
    """(B, pitch) nibble table -> (B, A) uint8 codes (include/nwb.h layout)."""
    out = np.empty((packed.shape[0], packed.shape[1] * 2), np.uint8)
    out[:, 0::2] = packed & 0xF
    out[:, 1::2] = packed >> 4
    return out[:, :a] & 7


def fill(top, side, m: int, k: int, d: int, flags: int = 0, device: int = 0, num_gpus: int = 1) -> Table:
    """nwb_fill_on(): one pair, host buffers in, table handle out."""
    top, side = _b(top), _b(side)
    h = C.c_void_p()
    rc = load_library().nwb_fill_on(top, len(top), side, len(side), m, k, d, flags, device, num_gpus, C.byref(h))
    _ck(rc, "nwb_fill_on")
    return Table(h.value, top, side)


class Plan:
    """Device-resident workspace (nwb_plan): strings in HBM, repeated fills."""

    def __init__(self, max_top: int, max_side: int, flags: int = 0, device: int = 0,
                 strip_rank: int = 0, strip_world: int = 1):
        self._h = C.c_void_p()
        _ck(load_library().nwb_plan_create(max_top, max_side, flags, device, strip_rank, strip_world,
                                           C.byref(self._h)), "nwb_plan_create")
        self.top_len = self.side_len = 0

    def upload(self, top, side) -> None:
        top, side = _b(top), _b(side)
        _ck(load_library().nwb_plan_upload(self._h, top, len(top), side, len(side)), "nwb_plan_upload")
        self.top_len, self.side_len = len(top), len(side)

    def run(self, m: int, k: int, d: int, stream: int | None = None) -> None:
        _ck(load_library().nwb_plan_run(self._h, m, k, d, C.c_void_p(stream or 0)), "nwb_plan_run")

    def summary(self) -> Summary:
        s = Summary()
        _ck(load_library().nwb_plan_summary(self._h, C.byref(s)), "nwb_plan_summary")
        return s

    def kernel_ms(self) -> float:
        return load_library().nwb_plan_kernel_ms(self._h)

    def launches(self) -> int:
        return load_library().nwb_plan_launches(self._h)

    def kernel_name(self) -> str:
        return load_library().nwb_plan_kernel_name(self._h).decode()

    def arrow_pitch(self) -> int:
        return load_library().nwb_plan_arrow_pitch(self._h)

    def arrow_digest(self) -> int:
        """Digest of this plan's share of the arrow table (device kernel; the shares of a strip group add up)."""
        d = C.c_uint64()
        _ck(load_library().nwb_plan_arrow_digest(self._h, C.byref(d)), "nwb_plan_arrow_digest")
        return d.value

    def count_path(self) -> str:
        return load_library().nwb_plan_count_path_name(self._h).decode()

    def arrows_device(self) -> int:
        return load_library().nwb_plan_arrows_device(self._h) or 0

    def strip_range(self) -> tuple[int, int]:
        b, e = C.c_int(), C.c_int()
        _ck(load_library().nwb_plan_strip_range(self._h, C.byref(b), C.byref(e)), "nwb_plan_strip_range")
        return b.value, e.value

    def download_arrows(self, row_begin: int = 0, row_end: int | None = None) -> np.ndarray:
        row_end = self.side_len if row_end is None else row_end
        pitch = self.arrow_pitch()
        out = np.zeros((row_end - row_begin, pitch), np.uint8)
        _ck(load_library().nwb_plan_download_arrows(self._h, out.ctypes.data_as(C.c_void_p), pitch,
                                                    row_begin, row_end), "nwb_plan_download_arrows")
        return out

    def run_pipelined(self, m: int, k: int, d: int, stream: int | None = None) -> None:
        """One fill of a queue of fills of a strip group: no reset_inbox / barrier between consecutive runs."""
        _ck(load_library().nwb_plan_run_pipelined(self._h, m, k, d, C.c_void_p(stream or 0)), "nwb_plan_run_pipelined")

    def attach_right(self, right: "Plan") -> None:
        """Same process, another GPU: my last strip streams into `right`'s inbox (peer access)."""
        _ck(load_library().nwb_plan_attach_right(self._h, right._h), "nwb_plan_attach_right")

    def reset_inbox(self, stream: int | None = None) -> None:
        _ck(load_library().nwb_plan_reset_inbox(self._h, C.c_void_p(stream or 0)), "nwb_plan_reset_inbox")

    def ipc_export(self) -> bytes:
        n = load_library().nwb_plan_ipc_size()
        buf = C.create_string_buffer(n)
        _ck(load_library().nwb_plan_ipc_export(self._h, buf), "nwb_plan_ipc_export")
        return buf.raw

    def ipc_attach_right(self, blob: bytes) -> None:
        buf = C.create_string_buffer(blob, len(blob))
        _ck(load_library().nwb_plan_ipc_attach_right(self._h, buf), "nwb_plan_ipc_attach_right")

    def close(self) -> None:
        if self._h:
            load_library().nwb_plan_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """Batch of independent pairs (nwb_batch): one warp per pair on the GPU."""

    def __init__(self, tops: list[bytes], sides: list[bytes], m: int, k: int, d: int, flags: int = 0,
                 device: int = 0):
        assert len(tops) == len(sides)
        self.n = len(tops)
        self.tops, self.sides = tops, sides
        tcat, scat = b"".join(tops), b"".join(sides)
        self._toff = np.zeros(self.n + 1, np.int64)
        self._soff = np.zeros(self.n + 1, np.int64)
        np.cumsum([len(t) for t in tops], out=self._toff[1:])
        np.cumsum([len(s) for s in sides], out=self._soff[1:])
        self._h = C.c_void_p()
        _ck(load_library().nwb_batch_create(tcat, self._toff.ctypes.data_as(C.c_void_p), scat,
                                            self._soff.ctypes.data_as(C.c_void_p), self.n, m, k, d, flags, device,
                                            C.byref(self._h)), "nwb_batch_create")

    @classmethod
    def from_arrays(cls, tcat: bytes, toff: np.ndarray, scat: bytes, soff: np.ndarray, m, k, d, flags=0, device=0):
        self = cls.__new__(cls)
        self.n = len(toff) - 1
        self.tops = self.sides = None
        self._toff = np.ascontiguousarray(toff, np.int64)
        self._soff = np.ascontiguousarray(soff, np.int64)
        self._h = C.c_void_p()
        _ck(load_library().nwb_batch_create(tcat, self._toff.ctypes.data_as(C.c_void_p), scat,
                                            self._soff.ctypes.data_as(C.c_void_p), self.n, m, k, d, flags, device,
                                            C.byref(self._h)), "nwb_batch_create")
        return self

    def run(self, stream: int | None = None) -> None:
        _ck(load_library().nwb_batch_run(self._h, C.c_void_p(stream or 0)), "nwb_batch_run")

    def refill(self, tcat: bytes, scat: bytes) -> None:
        """New strings for the same shapes from host buffers, H2D chunks overlapped with the kernels (nwb_batch_refill)."""
        keep = (tcat, scat)  # the copies are asynchronous: keep the buffers alive until fetch()
        self._refill_src = keep
        _ck(load_library().nwb_batch_refill(self._h, _host_ptr(tcat), _host_ptr(scat)), "nwb_batch_refill")

    def fetch(self) -> None:
        _ck(load_library().nwb_batch_fetch(self._h), "nwb_batch_fetch")

    def kernel_ms(self) -> float:
        return load_library().nwb_batch_kernel_ms(self._h)

    def digest(self, first_pair: int = 0) -> tuple[int, int, int, int]:
        """(arrow, score, branch, count) digests of the whole batch, computed on the device (nwb_batch_digest)."""
        out = (C.c_uint64 * 4)()
        _ck(load_library().nwb_batch_digest(self._h, first_pair, out), "nwb_batch_digest")
        return tuple(int(x) for x in out)

    def kernel_name(self) -> str:
        return (load_library().nwb_batch_kernel_name(self._h) or b"").decode()

    def count(self, p: int) -> int:
        """Optimal alignments of pair p mod 2^64 (NWB_WANT_COUNT; after fetch())."""
        return int(load_library().nwb_batch_count_u64(self._h, p))

    def launches(self) -> int:
        return load_library().nwb_batch_launches(self._h)

    def opt_score(self, p: int) -> int:
        return load_library().nwb_batch_opt_score(self._h, p)

    def branch_count(self, p: int) -> int:
        return load_library().nwb_batch_branch_count(self._h, p)

    def greatest_abs(self, p: int) -> int:
        return load_library().nwb_batch_greatest_abs(self._h, p)

    def score_rows(self, p: int) -> np.ndarray | None:
        """(B, A) int32 interior scores of pair p (NWB_WANT_SCORES; after fetch())."""
        pitch = C.c_size_t()
        ptr = load_library().nwb_batch_score_rows(self._h, p, C.byref(pitch))
        a = int(self._toff[p + 1] - self._toff[p])
        b = int(self._soff[p + 1] - self._soff[p])
        if not ptr or b == 0:
            return None
        buf = (C.c_int32 * (pitch.value * b)).from_address(ptr)
        return np.frombuffer(buf, np.int32).reshape(b, pitch.value)[:, :a].copy()

    def arrow_rows(self, p: int) -> np.ndarray | None:
        pitch = C.c_size_t()
        ptr = load_library().nwb_batch_arrow_rows(self._h, p, C.byref(pitch))
        b = int(self._soff[p + 1] - self._soff[p])
        if not ptr or b == 0:
            return None
        buf = (C.c_uint8 * (pitch.value * b)).from_address(ptr)
        return np.frombuffer(buf, np.uint8).reshape(b, pitch.value).copy()

    def close(self) -> None:
        if self._h:
            load_library().nwb_batch_free(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Computation:
    """Mirror of the reference's computation_t (computation.h:45-75) around the
    C ABI: construct with the reference's init_computation() arguments, then
    call compute_table_scores() where the reference calls it
    (needleman-wunsch.c:662)."""

    def __init__(self, s1, s2, m: int, k: int, d: int, *, sflag=False, tflag=False, qflag=False,
                 lflag=False, num_gpus: int = 1, device: int = 0):
        self.top_string, self.side_string = _b(s1), _b(s2)
        self.match_score, self.mismatch_penalty, self.indel_penalty = m, k, d
        self.M, self.N = len(self.top_string) + 1, len(self.side_string) + 1
        self.sflag, self.tflag, self.qflag, self.lflag = sflag, tflag, qflag, lflag
        self.num_gpus, self.device = num_gpus, device
        self.table: Table | None = None

    def flags(self) -> int:
        f = 0
        if self.sflag:
            f |= WANT_COUNT
        if self.tflag:
            f |= WANT_SCORES | TRACK_ABS | WANT_ARROWS_HOST
        # the walk runs iff !q || l || s || t (needleman-wunsch.c:667); with the
        # fused count, -q -s alone no longer needs the arrows on the host
        if (not self.qflag) or self.lflag or self.tflag:
            f |= WANT_ARROWS_HOST
        return f

    def compute_table_scores(self) -> "Computation":
        self.table = fill(self.top_string, self.side_string, self.match_score, self.mismatch_penalty,
                          self.indel_penalty, self.flags(), self.device, self.num_gpus)
        return self

    # what print_summary() reads (computation.c:271-281)
    @property
    def solution_count(self) -> int:
        return self.table.count & 0xFFFFFFFF

    @property
    def optimal_score(self) -> int:
        return self.table.opt_score

    @property
    def branch_count(self) -> int:
        return self.table.branch_count

    @property
    def greatest_abs_val(self) -> int:
        return self.table.greatest_abs

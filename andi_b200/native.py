"""ctypes binding of libandi_b200.so (include/andi_b200.h).

This module is plumbing: it loads the in-tree CUDA library and exposes its C ABI with numpy
arrays. There is no CPU fallback -- if the library is missing or no CUDA device is present
every computing call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("ANDI_B200_LIB", HERE / "libandi_b200.so"))  # override: kernel experiments only

MODELS = {"RAW": 0, "JC": 1, "KIMURA": 2, "LOGDET": 3, "ANI": 4}
ESA_SEARCH, ESA_FULL = 0, 1
ERRORS = {1: "bad argument", 2: "CUDA failure", 3: "out of memory", 4: "sequence too long"}

# every symbol include/andi_b200.h declares (tests check that the library exports them all)
ABI_SYMBOLS = [
    "andi_ctx_create", "andi_ctx_destroy", "andi_last_error", "andi_pool_set_host", "andi_pool_set_device",
    "andi_pool_size", "andi_pool_info", "andi_threshold", "andi_esa_build", "andi_esa_build_rs", "andi_esa_free",
    "andi_esa_len", "andi_esa_download", "andi_esa_get_match", "andi_dist_row", "andi_dist_anchor",
    "andi_dist_rows", "andi_dist_rows_device", "andi_get_stats", "andi_reset_stats",
    "andi_pool_export", "andi_pool_import", "andi_dist_matrix_multi", "andi_device_count",
]


class AndiError(RuntimeError):
    pass


class Model(C.Structure):
    _fields_ = [("counts", C.c_uint32 * 16), ("seq_len", C.c_uint32)]


class LcpInter(C.Structure):
    _fields_ = [("l", C.c_int32), ("i", C.c_int32), ("j", C.c_int32), ("m", C.c_int32)]


class PoolView(C.Structure):
    """andi_pool_view: the packed pool of a context (device pointers + per-sequence facts)."""

    _fields_ = [("d_code", C.c_void_p), ("d_spec", C.c_void_p), ("words", C.c_size_t), ("n", C.c_size_t),
                ("lens", C.POINTER(C.c_size_t)), ("gc", C.POINTER(C.c_double)), ("has_separator", C.POINTER(C.c_int)),
                ("any_separator", C.c_int)]


class Stats(C.Structure):
    _fields_ = [
        ("esa_ms", C.c_double), ("walk_ms", C.c_double), ("total_ms", C.c_double),
        ("esa_launches", C.c_uint64), ("cub_calls", C.c_uint64), ("walk_launches", C.c_uint64), ("pairs", C.c_uint64),
        ("subjects", C.c_uint64), ("sa_rounds", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("p2p_bytes", C.c_uint64), ("rows_ms", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def load() -> C.CDLL:
    """Load libandi_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise AndiError(
            f"{LIB_PATH} is missing: build it with `make -C andi_b200/csrc` "
            "(or __graft_entry__.build()). There is no CPU fallback."
        )
    L = C.CDLL(str(LIB_PATH))
    vp, sz = C.c_void_p, C.c_size_t
    L.andi_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.andi_ctx_destroy.argtypes = [vp]
    L.andi_ctx_destroy.restype = None
    L.andi_last_error.argtypes = [vp]
    L.andi_last_error.restype = C.c_char_p
    L.andi_pool_set_host.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(sz), sz]
    L.andi_pool_set_device.argtypes = [vp, vp, C.POINTER(sz), C.POINTER(sz), sz]
    L.andi_pool_size.argtypes = [vp]
    L.andi_pool_size.restype = sz
    L.andi_pool_info.argtypes = [vp, sz, C.POINTER(sz), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.andi_threshold.argtypes = [C.c_double, C.c_double, sz]
    L.andi_threshold.restype = sz
    L.andi_esa_build.argtypes = [vp, sz, C.c_uint, C.POINTER(vp)]
    L.andi_esa_build_rs.argtypes = [vp, C.c_char_p, sz, C.c_uint, C.POINTER(vp)]
    L.andi_esa_free.argtypes = [vp]
    L.andi_esa_free.restype = None
    L.andi_esa_len.argtypes = [vp]
    L.andi_esa_len.restype = C.c_int32
    L.andi_esa_download.argtypes = [vp, vp, vp, vp, vp, vp]
    L.andi_esa_get_match.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(sz), sz, C.POINTER(LcpInter)]
    L.andi_dist_row.argtypes = [vp, vp, C.POINTER(sz), sz, sz, C.c_int, C.POINTER(Model)]
    L.andi_dist_anchor.argtypes = [vp, vp, C.c_char_p, sz, sz, C.c_int, C.POINTER(Model)]
    L.andi_dist_rows.argtypes = [vp, sz, sz, C.c_double, C.c_int, C.c_int, vp]
    L.andi_dist_rows_device.argtypes = [vp, sz, sz, C.c_double, C.c_int, C.c_int, vp]
    L.andi_pool_export.argtypes = [vp, C.POINTER(PoolView)]
    L.andi_pool_import.argtypes = [vp, C.POINTER(PoolView), C.c_int]
    L.andi_device_count.argtypes = []
    L.andi_dist_matrix_multi.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_char_p), C.POINTER(sz), sz, C.c_double, C.c_int,
                                         C.c_int, vp, vp, vp, C.c_char_p, sz]
    L.andi_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.andi_reset_stats.argtypes = [vp]
    L.andi_reset_stats.restype = None
    _lib = L
    return L


def threshold(p_value: float, gc: float, rs_len: int) -> int:
    """min_anchor_length (src/sequence.c:296-304) -- host double precision, no GPU needed."""
    return int(load().andi_threshold(p_value, gc, rs_len))


class Esa:
    """Device-resident index of one subject (the reference's esa_s, src/esa.h:42-59)."""

    def __init__(self, ctx: "Context", handle):
        self.ctx, self.h = ctx, handle
        self.N = int(load().andi_esa_len(handle))
        ctx._esas.add(self)  # the context frees its live indexes before it goes (andi_esa_free needs the context)

    def download(self, full: bool = False) -> dict:
        N = self.N
        out = {"SA": np.empty(N, np.int32), "LCP": np.empty(N + 1, np.int32)}
        if full:
            out["CLD"] = np.empty(N + 1, np.int32)
            out["FVC"] = np.empty(N, np.uint8)
            out["cache"] = np.empty((1 << 20, 4), np.int32)
        ptr = lambda k: out[k].ctypes.data if k in out else None
        self.ctx._ck(load().andi_esa_download(self.h, ptr("SA"), ptr("LCP"), ptr("CLD"), ptr("FVC"), ptr("cache")))
        return out

    def get_match(self, queries) -> np.ndarray:
        """get_match (src/esa.c:614-624) for a batch: int32 array (nq, 4) = l, i, j, m."""
        qs = [bytes(q) for q in queries]
        n = len(qs)
        arr = (C.c_char_p * n)(*qs)
        lens = (C.c_size_t * n)(*[len(q) for q in qs])
        out = (LcpInter * n)()
        self.ctx._ck(load().andi_esa_get_match(self.h, arr, lens, n, out))
        return np.frombuffer(out, dtype=np.int32).reshape(n, 4).copy()

    def free(self):
        if self.h and self.ctx.h:
            load().andi_esa_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One GPU. Mirrors what src/process.c:230-270 + src/dist_hack.h hold for a run."""

    def __init__(self, device: int = 0, stream: int | None = None):
        L = load()
        h = C.c_void_p()
        rc = L.andi_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc:
            raise AndiError(f"andi_ctx_create: {ERRORS.get(rc, rc)}: {L.andi_last_error(None).decode()}")
        self.h = h
        self._keep = None
        self._esas = weakref.WeakSet()

    def _ck(self, rc: int):
        if rc:
            raise AndiError(f"{ERRORS.get(rc, rc)}: {load().andi_last_error(self.h).decode()}")

    def set_pool(self, seqs):
        seqs = [bytes(s) for s in seqs]
        n = len(seqs)
        arr = (C.c_char_p * n)(*seqs)
        lens = (C.c_size_t * n)(*[len(s) for s in seqs])
        self._ck(load().andi_pool_set_host(self.h, arr, lens, n))
        self.n = n

    def set_pool_device(self, dev_ptr: int, offsets, lens):
        n = len(lens)
        offs = (C.c_size_t * n)(*[int(o) for o in offsets])
        ls = (C.c_size_t * n)(*[int(x) for x in lens])
        self._ck(load().andi_pool_set_device(self.h, C.c_void_p(dev_ptr), offs, ls, n))
        self.n = n

    def pool_export(self) -> dict:
        """The packed pool as plain values: device pointers of the two planes (on this context's
        device), their length in u64 words, and the per-sequence facts as numpy arrays."""
        v = PoolView()
        self._ck(load().andi_pool_export(self.h, C.byref(v)))
        n = v.n
        return {"d_code": v.d_code, "d_spec": v.d_spec, "words": v.words, "n": n,
                "lens": np.ctypeslib.as_array(v.lens, (n,)).astype(np.uint64).copy(),
                "gc": np.ctypeslib.as_array(v.gc, (n,)).copy(),
                "has_separator": np.ctypeslib.as_array(v.has_separator, (n,)).astype(np.int32).copy(),
                "any_separator": bool(v.any_separator)}

    def pool_import(self, view: dict, src_device: int = -1):
        """Adopt a packed pool (a copy of its planes): from another context of this process
        (src_device = its device: peer copy) or from planes already on this device (-1)."""
        n = int(view["n"])
        lens = (C.c_size_t * n)(*[int(x) for x in view["lens"]])
        gc = (C.c_double * n)(*[float(x) for x in view["gc"]])
        sep = (C.c_int * n)(*[int(x) for x in view["has_separator"]])
        v = PoolView(C.c_void_p(view["d_code"]), C.c_void_p(view["d_spec"]) if view.get("d_spec") else None, int(view["words"]), n,
                     lens, gc, sep, int(bool(view["any_separator"])))
        self._ck(load().andi_pool_import(self.h, C.byref(v), src_device))
        self.n = n

    def pool_info(self, k: int):
        ln, gc, sep = C.c_size_t(), C.c_double(), C.c_int()
        self._ck(load().andi_pool_info(self.h, k, C.byref(ln), C.byref(gc), C.byref(sep)))
        return ln.value, gc.value, bool(sep.value)

    def esa_build(self, subject: int, full: bool = False) -> Esa:
        h = C.c_void_p()
        self._ck(load().andi_esa_build(self.h, subject, ESA_FULL if full else ESA_SEARCH, C.byref(h)))
        return Esa(self, h)

    def esa_build_rs(self, rs: bytes, full: bool = False) -> Esa:
        h = C.c_void_p()
        self._ck(load().andi_esa_build_rs(self.h, rs, len(rs), ESA_FULL if full else ESA_SEARCH, C.byref(h)))
        return Esa(self, h)

    def dist_row(self, esa: Esa, query_ids, threshold: int, model: str = "JC") -> np.ndarray:
        n = len(query_ids)
        ids = (C.c_size_t * n)(*[int(q) for q in query_ids])
        out = (Model * n)()
        self._ck(load().andi_dist_row(self.h, esa.h, ids, n, threshold, MODELS[model], out))
        return np.frombuffer(out, dtype=np.uint32).reshape(n, 17).copy()

    def dist_anchor(self, esa: Esa, query: bytes, threshold: int, model: str = "JC") -> np.ndarray:
        out = Model()
        self._ck(load().andi_dist_anchor(self.h, esa.h, query, len(query), threshold, MODELS[model], C.byref(out)))
        return np.array(list(out.counts) + [out.seq_len], dtype=np.uint32)

    def dist_rows(self, s_begin: int = 0, s_end: int | None = None, p_value: float = 0.025, model: str = "JC",
                  low_memory: bool = False, out: np.ndarray | None = None) -> np.ndarray:
        s_end = self.n if s_end is None else s_end
        rows = s_end - s_begin
        if out is None:
            out = np.empty((rows, self.n, 17), np.uint32)
        assert out.dtype == np.uint32 and out.size == rows * self.n * 17 and out.flags.c_contiguous
        self._ck(load().andi_dist_rows(self.h, s_begin, s_end, p_value, MODELS[model], int(low_memory),
                                       C.c_void_p(out.ctypes.data)))
        return out

    def dist_rows_device(self, dev_ptr: int, s_begin: int, s_end: int, p_value: float = 0.025, model: str = "JC",
                         low_memory: bool = False):
        """Rows [s_begin, s_end) written to device memory at dev_ptr ((s_end-s_begin) * n * 68 bytes)."""
        self._ck(load().andi_dist_rows_device(self.h, s_begin, s_end, p_value, MODELS[model], int(low_memory),
                                              C.c_void_p(dev_ptr)))

    def stats(self) -> dict:
        s = Stats()
        self._ck(load().andi_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        load().andi_reset_stats(self.h)

    def close(self):
        if self.h:
            for e in list(self._esas):  # andi_esa_free uses the context's device and stream
                e.free()
            load().andi_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def dist_matrix_multi(devices, seqs, p_value: float = 0.025, model: str = "JC", low_memory: bool = False) -> np.ndarray:
    """andi_dist_matrix_multi: the whole matrix on several GPUs of this process (one host thread
    per device, pool packed once and copied to the peers, subjects from a shared queue)."""
    seqs = [bytes(s) for s in seqs]
    n = len(seqs)
    arr = (C.c_char_p * n)(*seqs)
    lens = (C.c_size_t * n)(*[len(s) for s in seqs])
    dev = (C.c_int * len(devices))(*[int(d) for d in devices])
    out = np.empty((n, n, 17), np.uint32)
    msg = C.create_string_buffer(512)
    rc = load().andi_dist_matrix_multi(dev, len(devices), arr, lens, n, p_value, MODELS[model], int(low_memory),
                                       C.c_void_p(out.ctypes.data), None, None, msg, 512)
    if rc:
        raise AndiError(f"{ERRORS.get(rc, rc)}: {msg.value.decode()}")
    return out

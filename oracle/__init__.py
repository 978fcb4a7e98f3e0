"""oracle -- TEST INFRASTRUCTURE ONLY.

ctypes bindings for (a) ``andi_oracle.c``, our CPU restatement of the reference's hot path, and
(b) ``oracle/_ref/libandi_ref.so``, the unmodified reference sources compiled against the
shims under ``oracle/shim`` (see ``oracle/Makefile``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this package. The product (``andi_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "_build" / "libandi_oracle.so"
REF_SO = HERE / "_ref" / "libandi_ref.so"
REF_ANDI = HERE / "_ref" / "andi"
REF_TEST_FASTA = HERE / "_ref" / "test_fasta"

MODELS = {"RAW": 0, "JC": 1, "KIMURA": 2, "LOGDET": 3, "ANI": 4}


class Model(C.Structure):
    _fields_ = [("counts", C.c_uint32 * 16), ("seq_len", C.c_uint32)]


class Interval(C.Structure):
    _fields_ = [("l", C.c_int32), ("i", C.c_int32), ("j", C.c_int32), ("m", C.c_int32)]


class OrcEsa(C.Structure):
    _fields_ = [
        ("S", C.c_char_p),
        ("len", C.c_int32),
        ("SA", C.POINTER(C.c_int32)),
        ("LCP", C.POINTER(C.c_int32)),
        ("CLD", C.POINTER(C.c_int32)),
        ("FVC", C.POINTER(C.c_char)),
        ("cache", C.POINTER(Interval)),
    ]


class RefEsa(C.Structure):
    """Layout of the reference's esa_s (src/esa.h:42-59)."""

    _fields_ = [
        ("S", C.c_char_p),
        ("SA", C.POINTER(C.c_int32)),
        ("LCP", C.POINTER(C.c_int32)),
        ("len", C.c_int32),
        ("cache", C.POINTER(Interval)),
        ("FVC", C.POINTER(C.c_char)),
        ("CLD", C.POINTER(C.c_int32)),
    ]


class RefSeq(C.Structure):
    """seq_t (src/sequence.h:19-26)."""

    _fields_ = [("S", C.c_char_p), ("len", C.c_size_t), ("name", C.c_char_p)]


class RefSubject(C.Structure):
    """seq_subject (src/sequence.h:32-46)."""

    _fields_ = [("RS", C.c_void_p), ("RSlen", C.c_size_t), ("gc", C.c_double), ("threshold", C.c_size_t)]


def build_oracle(force: bool = False) -> Path:
    src = HERE / "andi_oracle.c"
    if force or not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < max(
        src.stat().st_mtime, (HERE / "andi_oracle.h").stat().st_mtime
    ):
        subprocess.run(["make", "-C", str(HERE), "oracle"], check=True, capture_output=True)
    return ORACLE_SO


def build_ref(reference: str = "/root/reference") -> bool:
    """Compile the reference into oracle/_ref (only possible where the reference tree exists)."""
    if not Path(reference).is_dir():
        return REF_SO.exists()
    subprocess.run(["make", "-C", str(HERE), "ref", f"REF={reference}"], check=True, capture_output=True)
    return True


_oracle = None
_ref = None


def lib() -> C.CDLL:
    global _oracle
    if _oracle is None:
        build_oracle()
        L = C.CDLL(str(ORACLE_SO))
        L.orc_normalize.restype = C.c_size_t
        L.orc_normalize.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        L.orc_make_rs.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.orc_gc.restype = C.c_double
        L.orc_gc.argtypes = [C.c_char_p, C.c_size_t]
        L.orc_shustring_cum_prob.restype = C.c_double
        L.orc_shustring_cum_prob.argtypes = [C.c_size_t, C.c_double, C.c_size_t]
        L.orc_min_anchor_length.restype = C.c_size_t
        L.orc_min_anchor_length.argtypes = [C.c_double, C.c_double, C.c_size_t]
        L.orc_suffix_array.argtypes = [C.c_char_p, C.POINTER(C.c_int32), C.c_int32]
        L.orc_esa_build.argtypes = [C.POINTER(OrcEsa), C.c_char_p, C.c_int32]
        L.orc_esa_free.argtypes = [C.POINTER(OrcEsa)]
        for f in (L.orc_get_match, L.orc_get_match_cld, L.orc_get_match_cached):
            f.restype = Interval
            f.argtypes = [C.POINTER(OrcEsa), C.c_char_p, C.c_size_t]
        for f in (L.orc_dist_anchor, L.orc_dist_anchor_spec):
            f.restype = Model
            f.argtypes = [C.POINTER(OrcEsa), C.c_char_p, C.c_size_t, C.c_size_t, C.c_int]
        L.orc_sweep_check.restype = C.c_size_t
        L.orc_sweep_check.argtypes = [C.POINTER(OrcEsa), C.c_int]
        L.orc_model_average.restype = Model
        L.orc_model_average.argtypes = [C.POINTER(Model), C.POINTER(Model)]
        L.orc_model_coverage.restype = C.c_double
        L.orc_model_coverage.argtypes = [C.POINTER(Model)]
        L.orc_estimate.restype = C.c_double
        L.orc_estimate.argtypes = [C.POINTER(Model), C.c_int]
        L.orc_rows.argtypes = [
            C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_size_t, C.c_size_t, C.c_size_t,
            C.c_int, C.c_double, C.POINTER(Model),
        ]
        _oracle = L
    return _oracle


def ref_available() -> bool:
    return REF_SO.exists()


def ref() -> C.CDLL:
    """The unmodified reference as a shared library (None-safe: raises if it was never built)."""
    global _ref
    if _ref is None:
        if not REF_SO.exists():
            raise FileNotFoundError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(str(REF_SO))
        L.seq_subject_init.argtypes = [C.POINTER(RefSubject), C.POINTER(RefSeq)]
        L.seq_subject_free.argtypes = [C.POINTER(RefSubject)]
        L.esa_init.argtypes = [C.POINTER(RefEsa), C.POINTER(RefSubject)]
        L.esa_free.argtypes = [C.POINTER(RefEsa)]
        for f in (L.get_match, L.get_match_cached):
            f.restype = Interval
            f.argtypes = [C.POINTER(RefEsa), C.c_char_p, C.c_size_t]
        L.dist_anchor.restype = Model
        L.dist_anchor.argtypes = [C.POINTER(RefEsa), C.c_char_p, C.c_size_t, C.c_size_t]
        L.min_anchor_length.restype = C.c_size_t
        L.min_anchor_length.argtypes = [C.c_double, C.c_double, C.c_size_t]
        L.shustring_cum_prob.restype = C.c_double
        L.shustring_cum_prob.argtypes = [C.c_size_t, C.c_double, C.c_size_t]
        for name in ("estimate_RAW", "estimate_JC", "estimate_KIMURA", "estimate_LOGDET", "estimate_ANI", "model_coverage"):
            f = getattr(L, name)
            f.restype = C.c_double
            f.argtypes = [C.POINTER(Model)]
        L.model_average.restype = Model
        L.model_average.argtypes = [C.POINTER(Model), C.POINTER(Model)]
        for f in (L.ref_rows, L.ref_rows_lm):
            f.argtypes = [
                C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_size_t, C.c_size_t, C.c_size_t,
                C.c_int, C.c_int, C.c_double, C.POINTER(Model), C.POINTER(C.c_double),
            ]
        _ref = L
    return _ref


def _seq_arrays(seqs):
    seqs = [s if isinstance(s, bytes) else bytes(s) for s in seqs]
    n = len(seqs)
    ptrs = (C.c_char_p * n)(*seqs)
    lens = (C.c_size_t * n)(*[len(s) for s in seqs])
    return seqs, ptrs, lens


def models_to_numpy(buf, rows: int, n: int) -> np.ndarray:
    return np.frombuffer(buf, dtype=np.uint32).reshape(rows, n, 17).copy()


def rows(seqs, model: str = "JC", p_value: float = 0.025, s_begin: int = 0, s_end: int | None = None) -> np.ndarray:
    """Oracle rows [s_begin, s_end) of the all-pairs matrix -> uint32 array (rows, n, 17)."""
    seqs, ptrs, lens = _seq_arrays(seqs)
    n = len(seqs)
    s_end = n if s_end is None else s_end
    out = (Model * ((s_end - s_begin) * n))()
    rc = lib().orc_rows(ptrs, lens, n, s_begin, s_end, MODELS[model], p_value, out)
    if rc:
        raise RuntimeError(f"orc_rows failed: {rc}")
    return models_to_numpy(out, s_end - s_begin, n)


def ref_rows(seqs, model: str = "JC", p_value: float = 0.025, s_begin: int = 0, s_end: int | None = None,
             threads: int = 1, low_memory: bool = False):
    """Reference rows (unmodified reference code) -> (uint32 array (rows, n, 17), timings)."""
    seqs, ptrs, lens = _seq_arrays(seqs)
    n = len(seqs)
    s_end = n if s_end is None else s_end
    out = (Model * ((s_end - s_begin) * n))()
    t = (C.c_double * 3)()
    fn = ref().ref_rows_lm if low_memory else ref().ref_rows
    rc = fn(ptrs, lens, n, s_begin, s_end, threads, MODELS[model], p_value, out, t)
    if rc:
        raise RuntimeError(f"ref_rows failed: {rc}")
    return models_to_numpy(out, s_end - s_begin, n), {"wall_s": t[0], "esa_s": t[1], "walk_s": t[2]}


class OracleEsa:
    """ESA of one subject built by the oracle (keeps RS alive)."""

    def __init__(self, seq: bytes):
        L = lib()
        n = len(seq)
        self.n = n
        self.rs_buf = C.create_string_buffer(2 * n + 2)
        L.orc_make_rs(seq, n, self.rs_buf)
        self.N = 2 * n + 1
        self.E = OrcEsa()
        rc = L.orc_esa_build(C.byref(self.E), self.rs_buf, self.N)
        if rc:
            raise RuntimeError(f"orc_esa_build failed: {rc}")

    @property
    def rs(self) -> bytes:
        return self.rs_buf.raw[: self.N]

    def array(self, name: str) -> np.ndarray:
        N = self.N
        if name == "SA":
            return np.ctypeslib.as_array(self.E.SA, (N,)).copy()
        if name == "LCP":
            return np.ctypeslib.as_array(self.E.LCP, (N + 1,)).copy()
        if name == "CLD":
            return np.ctypeslib.as_array(self.E.CLD, (N + 1,)).copy()
        if name == "FVC":
            return np.frombuffer(C.string_at(self.E.FVC, N), dtype=np.uint8).copy()
        if name == "cache":
            return np.frombuffer(C.string_at(self.E.cache, 16 << 20), dtype=np.int32).reshape(-1, 4).copy()
        raise KeyError(name)

    def get_match(self, q: bytes, kind: str = "spec"):
        f = {"spec": lib().orc_get_match, "cld": lib().orc_get_match_cld, "cached": lib().orc_get_match_cached}[kind]
        r = f(C.byref(self.E), q, len(q))
        return r.l, r.i, r.j, r.m

    def dist_anchor(self, q: bytes, threshold: int, model: str = "JC", spec: bool = False) -> np.ndarray:
        f = lib().orc_dist_anchor_spec if spec else lib().orc_dist_anchor
        m = f(C.byref(self.E), q, len(q), threshold, MODELS[model])
        return np.array(list(m.counts) + [m.seq_len], dtype=np.uint32)

    def close(self):
        if self.E.SA:
            lib().orc_esa_free(C.byref(self.E))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefEsaHandle:
    """ESA of one subject built by the unmodified reference (esa_init)."""

    def __init__(self, seq: bytes, p_value: float = 0.025):
        L = ref()
        C.c_double.in_dll(L, "ANCHOR_P_VALUE").value = p_value
        self.seq = seq
        self.base = RefSeq(seq, len(seq), b"x")
        self.subject = RefSubject()
        if L.seq_subject_init(C.byref(self.subject), C.byref(self.base)):
            raise RuntimeError("seq_subject_init failed")
        self.E = RefEsa()
        rc = L.esa_init(C.byref(self.E), C.byref(self.subject))
        if rc:
            raise RuntimeError(f"esa_init failed: {rc}")
        self.N = int(self.subject.RSlen)
        self.threshold = int(self.subject.threshold)
        self.gc = float(self.subject.gc)

    @property
    def rs(self) -> bytes:
        return C.string_at(self.subject.RS, self.N)

    def array(self, name: str) -> np.ndarray:
        N = self.N
        if name == "SA":
            return np.ctypeslib.as_array(self.E.SA, (N,)).copy()
        if name == "LCP":
            return np.ctypeslib.as_array(self.E.LCP, (N + 1,)).copy()
        if name == "CLD":
            return np.ctypeslib.as_array(self.E.CLD, (N + 1,)).copy()
        if name == "FVC":
            return np.frombuffer(C.string_at(self.E.FVC, N), dtype=np.uint8).copy()
        if name == "cache":
            return np.frombuffer(C.string_at(self.E.cache, 16 << 20), dtype=np.int32).reshape(-1, 4).copy()
        raise KeyError(name)

    def get_match(self, q: bytes, cached: bool = True):
        f = ref().get_match_cached if cached else ref().get_match
        r = f(C.byref(self.E), q, len(q))
        return r.l, r.i, r.j, r.m

    def dist_anchor(self, q: bytes, threshold: int | None = None, model: str = "JC") -> np.ndarray:
        L = ref()
        C.c_int.in_dll(L, "MODEL").value = MODELS[model]
        m = L.dist_anchor(C.byref(self.E), q, len(q), self.threshold if threshold is None else threshold)
        return np.array(list(m.counts) + [m.seq_len], dtype=np.uint32)

    def close(self):
        L = ref()
        if self.E.SA:
            L.esa_free(C.byref(self.E))
            L.seq_subject_free(C.byref(self.subject))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_test_fasta(seed: int, length: int, dists, raw: bool = False, line_length: int | None = None) -> bytes:
    """FASTA text from the reference's simulator (test/test_fasta.cxx), only where it was built."""
    cmd = [str(REF_TEST_FASTA), "-s", str(seed), "-l", str(length)]
    if line_length:
        cmd += ["-L", str(line_length)]
    if raw:
        cmd.append("-r")
    for d in dists:
        cmd += ["-d", repr(float(d))]
    return subprocess.run(cmd, check=True, capture_output=True).stdout


def parse_fasta(text: bytes):
    """Minimal FASTA reader for fixtures: returns [(name, normalized sequence bytes)]."""
    out = []
    name, chunks = None, []
    for line in text.splitlines():
        if line.startswith(b">"):
            if name is not None:
                out.append((name, b"".join(chunks)))
            name = line[1:].split()[0] if len(line) > 1 else b""
            chunks = []
        else:
            chunks.append(line.strip())
    if name is not None:
        out.append((name, b"".join(chunks)))
    res = []
    for nm, s in out:
        buf = C.create_string_buffer(s, len(s) + 1)
        flag = C.c_int(0)
        n = lib().orc_normalize(buf, C.byref(flag))
        res.append((nm.decode(), buf.raw[:n]))
    return res

"""The round-2 walk kernels (andi_b200/csrc/walk_v3.cuh) on the CPU: their per-lane logic
(walk_v3_lane.h) compiles both into the CUDA kernels and into a serial warp emulation
(csrc/emu/emu_v3.cpp). This test builds the index with numpy from the oracle's suffix array, runs
the emulation (32 lock-step lanes per warp, the kernel's service policy, PHASE 1 then PHASE 2),
reduces its records the way k_walk_reduce does and compares with the oracle. Test infrastructure
only; nothing here touches a GPU."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle
from andi_b200 import synth

ROOT = Path(__file__).resolve().parent.parent
EMU = ROOT / "andi_b200" / "csrc" / "emu"
UNIT_WORDS = 38
CODE = np.full(256, 255, dtype=np.uint8)
for ch, v in ((b"A", 0), (b"C", 1), (b"G", 2), (b"T", 3), (b"#", 1)):
    CODE[ch[0]] = v
STATS = ["trips", "ext_trips", "cand_trips", "steps", "lucky_hits", "lookups", "tag0", "tag1", "wide_gaps", "slow_steps", "slow_tail", "slow_tag3", "wide_pairs", "slow_long", "cols_trips", "pushes", "drained", "drains", "drain_rounds", "coop_scans", "ext_rounds", "bursts",
         "warp_trips", "running_lanes", "services", "served_lanes"]


def load_emu():
    so = EMU / "libemu_v3.so"
    src = [EMU / "emu_v3.cpp", EMU.parent / "walk_v3_lane.h"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in src):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so), str(src[0])], check=True,
                       cwd=EMU, capture_output=True)
    L = C.CDLL(str(so))
    L.emu_walk_v3.restype = C.c_long
    L.emu_walk_v3.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
    assert L.emu_v3_stats() == len(STATS)
    return L


@pytest.fixture(scope="module")
def emu():
    return load_emu()


def pack(text: bytes) -> np.ndarray:
    """text.cuh: 2 bits per character, 32 characters per u64 word, four zero guard words."""
    codes = CODE[np.frombuffer(text, dtype=np.uint8)].astype(np.uint64)
    assert (codes < 4).all()
    words = len(text) // 32 + 4
    padded = np.zeros(words * 32, dtype=np.uint64)
    padded[: len(text)] = codes
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]
    return (padded.reshape(words, 32) << shifts).sum(axis=1, dtype=np.uint64)


def directory_view(rs: bytes, SA: np.ndarray, K: int) -> np.ndarray:
    """fdir of sa_bucket.cuh / esa_kernels.cuh, restated with numpy: per k-mer tag 0 + longest
    present prefix / tag 1 + text position of its only suffix and the 15 bases behind the k-mer there / tag 2 + the text positions of its two suffixes
    (31 bits each) / tag 3 + first SA index, count."""
    N, mid = len(rs), len(rs) // 2
    codes = CODE[np.frombuffer(rs, dtype=np.uint8)].astype(np.int64)
    pos = np.arange(N)

    def mers(m):  # key of the m-mer at every position where it lies inside the text and off '#'
        ok = (pos + m <= N) & ~((pos <= mid) & (mid < pos + m))
        key = np.zeros(N, dtype=np.int64)
        for c in range(m):
            key = key * 4 + np.where(pos + c < N, codes[np.minimum(pos + c, N - 1)], 0)
        return ok, key

    ok, key = mers(K)
    rank = np.empty(N, dtype=np.int64)
    rank[SA] = np.arange(N)
    count = np.bincount(key[ok], minlength=4**K)
    first = np.full(4**K, N, dtype=np.int64)
    np.minimum.at(first, key[ok], rank[ok])
    # the valid suffixes of a k-mer are contiguous in SA
    last = np.zeros(4**K, dtype=np.int64)
    np.maximum.at(last, key[ok], rank[ok])
    present = count > 0
    assert (last[present] - first[present] + 1 == count[present]).all()
    plen = np.zeros(4**K, dtype=np.int64)
    for m in range(1, K):
        okm, keym = mers(m)
        have = np.bincount(keym[okm], minlength=4**m) > 0
        prefix = np.arange(4**K) >> (2 * (K - m))
        plen = np.where(have[prefix], m, plen)
    fdir = np.where(count == 0, plen, 0).astype(np.uint64)
    one = count == 1
    p_one = SA[first[one]].astype(np.int64)
    nxt = np.zeros(len(p_one), dtype=np.uint64)
    for j in range(15):  # the 15 bases behind the k-mer (whatever the planes hold there: zero past the end)
        at = p_one + K + j
        nxt |= np.where(at < N, codes[np.minimum(at, N - 1)], 0).astype(np.uint64) << np.uint64(2 * j)
    fdir[one] = (np.uint64(1) << np.uint64(62)) | (nxt << np.uint64(31)) | p_one.astype(np.uint64)
    two = count == 2
    fdir[two] = (np.uint64(2) << np.uint64(62)) | (SA[first[two] + 1].astype(np.uint64) << np.uint64(31)) | SA[first[two]].astype(np.uint64)
    many = count > 2
    fdir[many] = (np.uint64(3) << np.uint64(62)) | (count[many].astype(np.uint64) << np.uint64(32)) | first[many].astype(np.uint64)
    return fdir


def reduce_records(rec: np.ndarray, qlens, chunk: int, cpq: int, threshold: int, skip: int, quarter: bool = True, queries=None) -> np.ndarray:
    """k_walk_reduce for records whose boundaries all synchronised, plus walk_tail
    (src/process.c:199-211): the last anchor's interior by the len/4 split of src/model.c:247-254
    (RAW / JC / KIMURA) or by the composition of its query slice (src/model.c:259-278)."""
    out = np.zeros((len(qlens), 17), dtype=np.uint32)
    for k, qlen in enumerate(qlens):
        if k == skip:
            out[k, 0] = out[k, 16] = 9  # src/dist_hack.h:61-64
            continue
        nch = -(-qlen // chunk)
        r = rec[k * cpq : k * cpq + nch].astype(np.int64)
        total = r[:, :16].sum(axis=0) + r[:-1, 16:32].astype(np.int32).sum(axis=0)
        # chains that reached the end of the query apart (flag 0): the first such chunk holds the true final state
        apart = np.nonzero(r[:-1, 37] == 0)[0]
        fin = r[apart[0]] if len(apart) else r[-1]
        last_q, last_len, paired = int(fin[34]), int(fin[35]), int(fin[36])
        start, tail = (0, qlen) if last_len >= qlen else ((last_q, last_len) if (paired or last_len >= 2 * threshold) else (0, 0))
        if quarter:
            for cell in (0, 5, 10):
                total[cell] += tail // 4
            total[15] += tail // 4 + tail % 4
        else:
            comp = np.bincount(CODE[np.frombuffer(queries[k][start : start + tail], dtype=np.uint8)], minlength=4)
            for b in range(4):
                total[5 * b] += int(comp[b])
        out[k, :16] = total.astype(np.uint32)
        out[k, 16] = qlen
    return out


def emulate_rows(emu, seqs, chunk, warps=3, stats=None, model="JC"):
    """All rows of the matrix through the emulation; returns (rows, per-phase statistics)."""
    qplanes = [pack(s) for s in seqs]
    q_off = np.cumsum([0] + [len(p) for p in qplanes[:-1]]).astype(np.uint64)
    pool = np.concatenate(qplanes)
    q_len = np.array([len(s) for s in seqs], dtype=np.uint32)
    cpq = max(1, -(-int(q_len.max()) // chunk))
    rows = np.zeros((len(seqs), len(seqs), 17), dtype=np.uint32)
    total = np.zeros(2 * len(STATS), dtype=np.uint64)
    for i, subject in enumerate(seqs):
        o = oracle.OracleEsa(subject)
        rs, SA = o.rs, o.array("SA").astype(np.uint32)
        N = len(rs)
        t = int(oracle.lib().orc_min_anchor_length(0.025, oracle.lib().orc_gc(subject, len(subject)), N))
        K = min(t, max(4, round(np.log(N) / np.log(4))))
        s_code, fdir = pack(rs), directory_view(rs, SA, K)
        rec = np.zeros((len(seqs) * cpq, UNIT_WORDS), dtype=np.uint32)
        st = np.zeros(2 * len(STATS), dtype=np.uint64)
        rc = emu.emu_walk_v3(s_code.ctypes.data, N, N // 2, SA.ctypes.data, fdir.ctypes.data, K, i, t, pool.ctypes.data,
                             q_off.ctypes.data, q_len.ctypes.data, len(seqs), chunk, cpq, rec.ctypes.data, st.ctypes.data, warps,
                             int(model in ("RAW", "JC", "KIMURA")))
        assert rc == 0
        total += st
        rows[i] = reduce_records(rec, [int(x) for x in q_len], chunk, cpq, t, skip=i, quarter=model in ("RAW", "JC", "KIMURA"),
                                 queries=seqs)
        o.close()
    n = len(STATS)
    return rows, {"phase1": dict(zip(STATS, total[:n].tolist())), "phase2": dict(zip(STATS, total[n:].tolist()))}


# (short chunks on purpose: boundaries inside repeats, inside anchors longer than a chunk and in anchor-free texts make the
# boundary replay run on beyond the next chunk, some of them to the end of the query)
@pytest.mark.parametrize("name,chunk", [("subst", 1000), ("subst", 4096), ("indel", 700), ("repeat", 1 << 20), ("repeat", 600),
                                        ("identical", 1 << 20), ("identical", 512), ("lowent", 1 << 20), ("lowent", 100), ("short", 1 << 20),
                                        ("short", 64), ("unrelated", 1 << 20), ("unrelated", 900), ("revcomp", 2048), ("copies", 1 << 20),
                                        ("copies", 1024), ("near", 2048), ("near", 1 << 20)])
@pytest.mark.parametrize("warps,model", [(1, "JC"), (5, "JC"), (3, "LOGDET")])
def test_v3_lane_logic_matches_the_oracle(emu, name, chunk, warps, model):
    from conftest import near_identical_sequences, stress_sequences

    seqs = near_identical_sequences() if name == "near" else [s for s in stress_sequences()[name] if b"!" not in s]
    want = oracle.rows(seqs, model)
    got, stats = emulate_rows(emu, seqs, chunk, warps, model=model)
    assert np.array_equal(got, want), (name, chunk, warps, model, stats)


def test_v3_lane_logic_on_random_pools(emu):
    """Seeded random pools: planted repeats (2 to 20 copies, 50 to 3000 bases), divergences from 0 to 0.3, indels, reverse
    complements, chunk lengths from 64 bases to one chunk per query, JC and LOGDET -- everything the lane logic has a special
    path for (cooperative bucket scan, replay beyond the next chunk, EXT bursts), against the oracle."""
    rng = np.random.default_rng(20261017)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for case in range(24):
        n = int(rng.integers(3000, 30000))
        base = synth.base_genome(n, int(rng.integers(1 << 30)))
        for _ in range(int(rng.integers(0, 4))):
            ln = int(rng.integers(50, min(3000, n // 4)))
            src = int(rng.integers(0, n - ln))
            element = base[src : src + ln].copy()
            for _ in range(int(rng.integers(1, 20))):
                at = int(rng.integers(0, n - ln))
                base[at : at + ln] = element
        seqs = []
        for _ in range(int(rng.integers(2, 4))):
            d = float(rng.choice([0.0, 1e-4, 1e-3, 0.01, 0.03, 0.1, 0.3]))
            s = synth.ACGT[synth.mutate(base, d, int(rng.integers(1 << 30)))].tobytes()
            if rng.random() < 0.4:
                s = synth.with_indels(s, int(rng.integers(1, 8)), int(rng.integers(1, 300)), int(rng.integers(1 << 30)))
            if rng.random() < 0.15:
                s = s.translate(comp)[::-1]
            seqs.append(s)
        chunk = int(rng.choice([64, 200, 513, 1024, 4096, 1 << 20]))
        model = str(rng.choice(["JC", "LOGDET"]))
        got, stats = emulate_rows(emu, seqs, chunk, int(rng.integers(1, 5)), model=model)
        assert np.array_equal(got, oracle.rows(seqs, model)), (case, n, chunk, model, [len(s) for s in seqs], stats)

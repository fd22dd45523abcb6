"""Logic of the GPU deflate encoder (vcf_b200/csrc/deflate_core.cuh) checked
without a GPU: tests/emul/deflate_emul.cpp compiles the same __host__ __device__ code with g++ and
runs the CTA's phases thread by thread; zlib's decoder must return the input."""
import ctypes
import os
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emul", "deflate_emul.cpp")
HDR = os.path.join(ROOT, "vcf_b200", "csrc", "deflate_core.cuh")
SO = os.path.join(ROOT, "build", "deflate_emul.so")


@pytest.fixture(scope="module")
def emul():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-shared", "-fPIC", "-o", SO, SRC])
    L = ctypes.CDLL(SO)
    L.dfl_emul.restype = ctypes.c_longlong
    L.dfl_emul.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_longlong,
                           ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong,
                           ctypes.POINTER(ctypes.c_longlong)]

    def run(data: np.ndarray, piece=258, nt=512, row=0, pixel=1, dists=None, model=1):
        data = np.ascontiguousarray(data, dtype=np.uint8).ravel()
        cap = data.size + data.size // 1000 + 5 * (data.size // (piece * nt) + 2) * (piece * nt // 65535 + 2) + 64
        out = np.empty(cap, np.uint8)
        ns = ctypes.c_longlong(0)
        d = np.asarray(dists if dists is not None else [], np.int32)
        n = L.dfl_emul(data.ctypes.data, data.size, piece, nt, row, pixel, d.ctypes.data, d.size, model, out.ctypes.data,
                       cap, ctypes.byref(ns))
        assert n > 0, f"emulation failed: {n}"
        return out[:n].tobytes(), ns.value
    return run


def _roundtrip(run, data, **kw):
    data = np.ascontiguousarray(data, dtype=np.uint8).ravel()
    raw, n_stored = run(data, **kw)
    back = zlib.decompress(raw, -15)
    assert back == data.tobytes()
    return len(raw), n_stored


def test_index_planes_roundtrip_and_size(emul):
    from oracle import vcf_oracle as O
    img = O.synthetic_frame(540, 960, 2, "natural")
    for q in (4, 16, 64):
        idx = O.encode_array(img, 8, q)
        size, _ = _roundtrip(emul, idx)
        nseg = -(-idx.size // (258 * 512))
        c = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_RLE)
        rle = len(c.compress(idx.tobytes()) + c.flush())
        ref = len(zlib.compress(idx.tobytes(), 6))
        assert size <= 1.05 * rle + 100 * nseg, (q, size, rle)        # same parse family as zlib's Z_RLE
        assert size <= 1.30 * ref + 100 * nseg, (q, size, ref)        # and close to zlib's default


def test_row_candidates_size(emul):
    """With the row geometry of the H x W x 3 index image (the previous sample of the channel and the
    samples above as match candidates, cost model on) the stream is never larger than the run-length
    parse and close to zlib level 6 (1080p: 1.03 / 1.11 / 1.09 / 1.05 x at q = 8 / 16 / 32 / 64,
    profiles/r2_deflate_candidates.json)."""
    from oracle import vcf_oracle as O
    img = O.synthetic_frame(540, 960, 2, "natural")
    for q in (8, 16, 32, 64):
        idx = O.encode_array(img, 8, q)
        row, px = idx.shape[1] * idx.shape[2], idx.shape[2]
        plain, _ = _roundtrip(emul, idx, model=0)
        runs, _ = _roundtrip(emul, idx)
        rows, _ = _roundtrip(emul, idx, row=row, pixel=px)
        ref = len(zlib.compress(idx.tobytes(), 6))
        assert runs <= plain + 64, (q, runs, plain)
        assert rows <= runs + 64, (q, rows, runs)
        assert rows <= 1.16 * ref, (q, rows, ref)
    noise = O.encode_array(O.synthetic_frame(272, 480, 3, "noise"), 8, 32)
    rows, _ = _roundtrip(emul, noise, row=noise.shape[1] * 3, pixel=3)
    assert rows <= len(zlib.compress(noise.tobytes(), 6))


def test_row_candidates_fuzz(emul):
    """Periodic and repeated-row inputs at every kind of row length (shorter than a match, longer than
    a piece, beyond the 32 KB window), matches that overlap their source, sources in the previous
    segment."""
    rng = np.random.default_rng(21)
    for it in range(80):
        row = int(rng.choice([1, 2, 3, 5, 7, 48, 255, 258, 300, 777, 5760, 32765, 32767, 32768, 40000]))
        px = int(rng.choice([1, 1, 2, 3, 4]))
        nrows = int(rng.integers(1, max(2, 120000 // row)))
        k = int(rng.integers(1, 40))
        base = rng.choice(k, size=row, p=rng.dirichlet(np.full(k, 0.4))).astype(np.uint8)
        rows = np.tile(base, (nrows, 1))
        flips = rng.random(rows.shape) < float(rng.choice([0.0, 0.002, 0.05, 0.5]))
        rows[flips] = rng.integers(0, 256, int(flips.sum()), dtype=np.uint8)
        if it % 4 == 0:
            rows = np.cumsum(rows, axis=0, dtype=np.uint8)      # rows that differ from the one above
        data = rows.ravel()[: int(rng.integers(1, rows.size + 1))]
        piece = int(rng.choice([8, 64, 258, 258, 516]))
        nt = int(rng.choice([1, 3, 32, 512]))
        _roundtrip(emul, data, piece=piece, nt=nt, row=row, pixel=px)
    for dists in ([1, 2], [1, 4, 5, 6, 7, 8, 9, 10], [1, 32768], [1, 32769, 40000]):
        data = np.tile(rng.integers(0, 5, 33000, dtype=np.uint8), 3)
        _roundtrip(emul, data, dists=dists)


def test_edge_cases(emul):
    rng = np.random.default_rng(7)
    cases = [
        np.zeros(0, np.uint8),
        np.array([5], np.uint8),
        np.array([5, 5], np.uint8),
        np.array([5, 5, 5], np.uint8),
        np.array([5, 5, 5, 5], np.uint8),
        np.full(257, 128, np.uint8), np.full(258, 128, np.uint8), np.full(259, 128, np.uint8),
        np.full(260, 128, np.uint8), np.full(261, 128, np.uint8), np.full(262, 128, np.uint8),
        np.full(100000, 128, np.uint8),
        np.arange(256, dtype=np.uint8),
        np.tile(np.arange(256, dtype=np.uint8), 300),
        rng.integers(0, 256, 70001, dtype=np.uint8),                       # incompressible: stored blocks
        rng.integers(0, 256, 65536 * 3 + 17, dtype=np.uint8),
        rng.integers(0, 2, 50000, dtype=np.uint8),
        np.repeat(rng.integers(0, 256, 3000, dtype=np.uint8), rng.integers(1, 600, 3000)),   # runs of every length
    ]
    for piece, nt in ((258, 512), (16, 4), (8, 3), (1032, 512), (516, 512), (256, 512)):
        for data in cases:
            _roundtrip(emul, data, piece=piece, nt=nt)
            _roundtrip(emul, data, piece=piece, nt=nt, model=0)
            _roundtrip(emul, data, piece=piece, nt=nt, row=48, pixel=3)


def test_every_run_length_at_every_alignment(emul):
    for L in list(range(1, 40)) + [255, 256, 257, 258, 259, 260, 300, 515, 516, 517, 518]:
        for lead in range(0, 9):
            data = np.concatenate([np.arange(1, lead + 1, dtype=np.uint8), np.full(L, 200, np.uint8), np.array([3, 3, 9], np.uint8)])
            _roundtrip(emul, data, piece=16, nt=8)
            _roundtrip(emul, data, piece=258, nt=512)


def test_stored_fallback_is_taken_for_noise(emul):
    rng = np.random.default_rng(3)
    data = rng.integers(0, 256, 200000, dtype=np.uint8)
    size, n_stored = _roundtrip(emul, data)
    assert n_stored > 0 and size <= data.size + 5 * 8 + 2


def test_skewed_alphabets_hit_the_length_limit(emul):
    # Fibonacci-like frequencies force code lengths beyond 15 bits before limiting
    fib = [1, 1]
    while len(fib) < 30:
        fib.append(fib[-1] + fib[-2])
    rng = np.random.default_rng(5)
    sym = np.concatenate([np.full(f, i, np.uint8) for i, f in enumerate(fib[:24])])
    rng.shuffle(sym)
    sym = sym[sym != np.roll(sym, 1)]          # no runs: literals only
    _roundtrip(emul, sym, piece=1032, nt=512)
    _roundtrip(emul, sym, piece=4096, nt=512)


def test_random_fuzz(emul):
    rng = np.random.default_rng(11)
    for it in range(60):
        n = int(rng.integers(1, 200000))
        k = int(rng.integers(1, 256))
        p = rng.dirichlet(np.full(k, 0.3))
        data = rng.choice(k, size=n, p=p).astype(np.uint8)
        if it % 3 == 0:
            data = np.repeat(data[: n // 8 + 1], rng.integers(1, 20, n // 8 + 1))
        piece = int(rng.choice([8, 16, 64, 258, 516, 1032]))
        nt = int(rng.choice([1, 2, 7, 32, 512]))
        _roundtrip(emul, data, piece=piece, nt=nt)


def test_crc32_in_chunks_equals_zlib(emul):
    L = ctypes.CDLL(SO)
    L.crc_emul.restype = ctypes.c_uint32
    L.crc_emul.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int]
    rng = np.random.default_rng(9)
    for n in (0, 1, 2, 7, 8, 9, 255, 256, 511, 512, 513, 1000, 65536, 100003, 3_000_001):
        data = rng.integers(0, 256, n + 16, dtype=np.uint8)
        for lead in (0, 3):                                   # aligned and unaligned starts
            view = np.ascontiguousarray(data[lead:lead + n])
            for chunk in (512, 1, 13, 4096):
                if chunk == 1 and n > 70000:
                    continue
                assert L.crc_emul(view.ctypes.data, n, chunk) == (zlib.crc32(view.tobytes()) & 0xFFFFFFFF), (n, lead, chunk)


def test_full_alphabet_and_degenerate_alphabets(emul):
    """All 286 literal/length symbols in one segment (every byte value, every match length), and
    the smallest alphabets (one literal + end of block; one literal + one length)."""
    rng = np.random.default_rng(13)
    parts = []
    for L in range(3, 259):                       # a run of L + 1 equal bytes = literal + match of length L
        parts.append(np.full(L + 1, int(rng.integers(0, 256)), np.uint8))
        parts.append(np.array([(int(parts[-1][0]) + 1) % 256], np.uint8))
    parts.append(np.arange(256, dtype=np.uint8))
    data = np.concatenate(parts)
    _roundtrip(emul, data, piece=1 << 16, nt=1)               # one thread: the sequential parse, all lengths intact
    _roundtrip(emul, data, piece=258, nt=512)
    for data in (np.array([7], np.uint8), np.full(2, 7, np.uint8), np.full(259, 7, np.uint8), np.full(258 * 4 + 1, 7, np.uint8)):
        _roundtrip(emul, data, piece=258, nt=512)
        _roundtrip(emul, data, piece=8, nt=2)

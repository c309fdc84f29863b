"""Synthetic inputs for the BASELINE.json configs + a ctypes binding of the system libzstd (1.5.5).

libzstd plays two roles here, neither on the measured path:
  * GENERATOR of the compressed frames of every synthetic config (the reference's own interop contract is
    "decode what C zstd level 3 produced", ruzstd/fuzz/fuzz_targets/interop.rs:31-33,55-65);
  * secondary oracle in tests (ZSTD_decompress on the same frames).
All randomness is numpy PCG64 with fixed seeds; datasets are cached under $B200Z_CACHE (default /tmp/b200z_cache).
"""
import ctypes as C
import hashlib
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_ROOT = os.path.dirname(os.path.abspath(__file__))
CACHE = os.environ.get("B200Z_CACHE", "/tmp/b200z_cache")

ZSTD_c_compressionLevel, ZSTD_c_windowLog, ZSTD_c_contentSizeFlag, ZSTD_c_checksumFlag = 100, 101, 200, 201

_z = None


def zstd():
    global _z
    if _z is None:
        for name in ("libzstd.so.1", "/usr/lib/x86_64-linux-gnu/libzstd.so.1"):
            try:
                _z = C.CDLL(name)
                break
            except OSError:
                continue
        if _z is None:
            raise RuntimeError("system libzstd.so.1 not found (needed only to GENERATE test/bench inputs)")
        z = _z
        z.ZSTD_compressBound.restype = C.c_size_t; z.ZSTD_compressBound.argtypes = [C.c_size_t]
        z.ZSTD_createCCtx.restype = C.c_void_p
        z.ZSTD_freeCCtx.argtypes = [C.c_void_p]
        z.ZSTD_CCtx_setParameter.restype = C.c_size_t; z.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
        z.ZSTD_compress2.restype = C.c_size_t; z.ZSTD_compress2.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        z.ZSTD_compress_usingDict.restype = C.c_size_t
        z.ZSTD_compress_usingDict.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        z.ZSTD_isError.restype = C.c_uint; z.ZSTD_isError.argtypes = [C.c_size_t]
        z.ZSTD_decompress.restype = C.c_size_t; z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        z.ZSTD_createDCtx.restype = C.c_void_p
        z.ZSTD_freeDCtx.argtypes = [C.c_void_p]
        z.ZSTD_decompress_usingDict.restype = C.c_size_t
        z.ZSTD_decompress_usingDict.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        z.ZSTD_DCtx_setParameter.restype = C.c_size_t; z.ZSTD_DCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
        z.ZSTD_decompressDCtx.restype = C.c_size_t; z.ZSTD_decompressDCtx.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    return _z


def compress(data, level=3, window_log=None, checksum=True, raw_dict=None):
    """One zstd frame for `data` (bytes / uint8 array)."""
    z = zstd()
    a = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data)
    cap = z.ZSTD_compressBound(a.nbytes)
    dst = np.empty(cap, dtype=np.uint8)
    cctx = z.ZSTD_createCCtx()
    try:
        if raw_dict is not None:
            d = np.ascontiguousarray(np.frombuffer(raw_dict, dtype=np.uint8) if isinstance(raw_dict, (bytes, bytearray)) else raw_dict)
            n = z.ZSTD_compress_usingDict(cctx, dst.ctypes.data, cap, a.ctypes.data, a.nbytes, d.ctypes.data, d.nbytes, level)
        else:
            z.ZSTD_CCtx_setParameter(cctx, ZSTD_c_compressionLevel, level)
            if window_log is not None:
                z.ZSTD_CCtx_setParameter(cctx, ZSTD_c_windowLog, window_log)
            z.ZSTD_CCtx_setParameter(cctx, ZSTD_c_checksumFlag, 1 if checksum else 0)
            n = z.ZSTD_compress2(cctx, dst.ctypes.data, cap, a.ctypes.data, a.nbytes)
    finally:
        z.ZSTD_freeCCtx(cctx)
    if z.ZSTD_isError(n):
        raise RuntimeError("libzstd compress failed")
    return dst[:n].tobytes()


def decompress(frame, out_size, raw_dict=None, window_log_max=None):
    """libzstd decode (secondary oracle)."""
    z = zstd()
    src = bytes(frame)
    dst = np.empty(max(out_size, 1), dtype=np.uint8)
    dctx = z.ZSTD_createDCtx()
    try:
        if window_log_max:
            z.ZSTD_DCtx_setParameter(dctx, 100, window_log_max)
        if raw_dict is not None:
            d = bytes(raw_dict)
            n = z.ZSTD_decompress_usingDict(dctx, dst.ctypes.data, out_size, src, len(src), d, len(d))
        else:
            n = z.ZSTD_decompressDCtx(dctx, dst.ctypes.data, out_size, src, len(src))
    finally:
        z.ZSTD_freeDCtx(dctx)
    if z.ZSTD_isError(n):
        raise RuntimeError("libzstd decompress failed")
    return dst[:n].tobytes()


# ------------------------------------------------------------------------------------------------------------
# plaintext generators
# ------------------------------------------------------------------------------------------------------------
_VOCAB = {}


def _vocab(nwords=50000, seed=0xE90001):
    """Pseudo-word vocabulary: word lengths ~ English, letters from a skewed alphabet; a few markup tokens."""
    key = (nwords, seed)
    if key in _VOCAB:
        return _VOCAB[key]
    rng = np.random.Generator(np.random.PCG64(seed))
    letters = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
    lp = 1.0 / np.arange(1, 27) ** 0.9
    lp /= lp.sum()
    lens = np.clip(rng.poisson(5.2, nwords) + 1, 1, 16)
    lens[:200] = np.clip(rng.poisson(2.0, 200) + 1, 1, 5)      # frequent words are short
    allc = letters[np.searchsorted(np.cumsum(lp), rng.random(int(lens.sum())))]
    cuts = np.concatenate([[0], np.cumsum(lens)])
    words = []
    for i in range(nwords):
        w = allc[cuts[i]:cuts[i + 1]].tobytes()
        if i % 97 == 0:
            w = w.capitalize()
        words.append(w + b" ")
    markup = [b"[[", b"]] ", b"'''", b"''", b"<ref>", b"</ref> ", b"{{cite ", b"}} ", b"== ", b" ==\n", b"\n\n", b"\n* ", b"&quot;", b"|", b". ",
              b", ", b"<page>\n", b"</page>\n", b"<title>", b"</title>\n", b"<text xml:space=\"preserve\">", b"</text>\n", b"1", b"2", b"19", b"20", b"0"]
    for i, m in enumerate(markup):
        words[3 + 7 * i] = m
    blob = np.frombuffer(b"".join(words), dtype=np.uint8)
    wl = np.array([len(w) for w in words], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(wl)[:-1]])
    _VOCAB[key] = (blob, off, wl)
    return _VOCAB[key]


def gen_text(nbytes, seed, zipf_s=1.05, reuse=0.50, run_lo=3, run_hi=24):
    """enwik-shaped text: Zipf word model + markup tokens + re-use of recent phrases."""
    blob, off, wl = _vocab()
    rng = np.random.Generator(np.random.PCG64(seed))
    nw = len(wl)
    p = 1.0 / np.arange(1, nw + 1) ** zipf_s
    p /= p.sum()
    cdf = np.cumsum(p)
    avg = float((p * wl).sum())
    n = int(nbytes / avg * 1.05) + 64
    ids = np.searchsorted(cdf, rng.random(n)).astype(np.int64)
    np.clip(ids, 0, nw - 1, out=ids)
    # phrase re-use: overwrite runs of word ids with a copy of an earlier run (up to ~10k words back ~ 64 KiB)
    nruns = int(n * reuse / ((run_lo + run_hi) / 2))
    if n - run_hi - 1 <= 64:   # a segment of a few dozen words: nothing to re-use from
        nruns = 0
    starts = rng.integers(64, max(65, n - run_hi - 1), nruns)
    lens = rng.integers(run_lo, run_hi, nruns)
    back = rng.integers(8, 10000, nruns)
    for s, ln, bk in zip(starts.tolist(), lens.tolist(), back.tolist()):
        src = s - bk
        if src >= 0:
            ids[s:s + ln] = ids[src:src + ln]
    l = wl[ids]
    ends = np.cumsum(l)
    total = int(ends[-1])
    # gather: byte k of the output belongs to word j = searchsorted(ends, k, 'right'); vectorised via repeat
    src_start = off[ids] - (ends - l)
    idx = np.repeat(src_start, l) + np.arange(total, dtype=np.int64)
    out = blob[idx]
    assert total >= nbytes
    return out[:nbytes]


def gen_skewed_bytes(nbytes, seed, entropy_bits=5.5):
    """i.i.d. bytes from a fixed Zipf-like 256-symbol distribution (C3: Huffman-heavy, no matches)."""
    rng0 = np.random.Generator(np.random.PCG64(0xC30000))
    perm = rng0.permutation(256)
    lo, hi = 0.01, 3.0
    for _ in range(40):  # tune the exponent so the entropy lands on the target
        s = (lo + hi) / 2
        p = 1.0 / np.arange(1, 257) ** s
        p /= p.sum()
        h = -(p * np.log2(p)).sum()
        if h > entropy_bits:
            lo = s
        else:
            hi = s
    rng = np.random.Generator(np.random.PCG64(seed))
    sym = np.searchsorted(np.cumsum(p), rng.random(nbytes))
    return perm[np.clip(sym, 0, 255)].astype(np.uint8)


def gen_binary_records(nbytes, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    rec = int(rng.integers(32, 129))
    proto = rng.integers(0, 256, rec, dtype=np.uint8)
    n = nbytes // rec + 1
    a = np.tile(proto, n)[:nbytes].copy()
    cols = rng.integers(0, rec, 4)
    for c in cols:  # a few varying fields (counters / ids)
        k = len(a[c::rec])
        a[c::rec] = (np.arange(k) * int(rng.integers(1, 7)) + int(rng.integers(0, 255))).astype(np.uint8)
    noise = rng.random(nbytes) < 0.05
    a[noise] = rng.integers(0, 256, int(noise.sum()), dtype=np.uint8)
    return a


def gen_silesia_mix(nbytes, seed):
    """C4: concatenation of segments -- constant runs (RLE blocks), text, binary records, low-entropy bytes, random (Raw blocks)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    parts, have = [], 0
    classes = ["const", "text", "records", "lowent", "random"]
    probs = [0.10, 0.35, 0.25, 0.15, 0.15]
    while have < nbytes:
        ln = int(np.exp(rng.uniform(np.log(16 << 10), np.log(512 << 10))))
        ln = min(ln, nbytes - have)
        cls = classes[int(rng.choice(5, p=probs))]
        s = int(rng.integers(1, 1 << 31))
        if cls == "const":
            seg = np.full(ln, int(rng.integers(0, 256)), dtype=np.uint8)
        elif cls == "text":
            seg = gen_text(ln, s)
        elif cls == "records":
            seg = gen_binary_records(ln, s)
        elif cls == "lowent":
            seg = gen_skewed_bytes(ln, s, 3.0)
        else:
            seg = rng.integers(0, 256, ln, dtype=np.uint8)
        parts.append(seg)
        have += ln
    return np.concatenate(parts)[:nbytes]


# ------------------------------------------------------------------------------------------------------------
# configs -> (compressed frames concatenated, frame table, plaintext)
# ------------------------------------------------------------------------------------------------------------
class FrameSet:
    """Independent frames packed back to back.  comp: uint8 array; src_off/src_size/out_off/out_size: uint64 arrays;
    plain: uint8 array holding every frame's expected plaintext at out_off."""

    def __init__(self, comp, src_off, src_size, out_off, out_size, plain, raw_dict=None, name=""):
        self.comp, self.src_off, self.src_size, self.out_off, self.out_size, self.plain = comp, src_off, src_size, out_off, out_size, plain
        self.raw_dict, self.name = raw_dict, name

    @property
    def nframes(self): return len(self.src_off)
    @property
    def C(self): return int(self.src_size.sum())
    @property
    def D(self): return int(self.out_size.sum())

    def frames_io(self, slack=0):
        from numpy import zeros
        a = zeros(self.nframes, dtype=[("src_off", "<u8"), ("src_size", "<u8"), ("out_off", "<u8"), ("out_cap", "<u8")])
        a["src_off"], a["src_size"], a["out_off"], a["out_cap"] = self.src_off, self.src_size, self.out_off, self.out_size + slack
        return a

    def sha256(self):
        return hashlib.sha256(self.plain.tobytes()).hexdigest()

    def subset(self, lo, hi):
        """Frames [lo, hi) as their own FrameSet: compressed bytes and plaintext re-based to offset 0 (views, no copies)."""
        if lo >= hi:
            z = np.zeros(0, dtype=np.uint64)
            return FrameSet(self.comp[:0], z, z, z, z, self.plain[:0], self.raw_dict, self.name)
        c0, c1 = int(self.src_off[lo]), int(self.src_off[hi - 1] + self.src_size[hi - 1])
        p0, p1 = int(self.out_off[lo]), int(self.out_off[hi - 1] + self.out_size[hi - 1])
        return FrameSet(self.comp[c0:c1], self.src_off[lo:hi] - np.uint64(c0), self.src_size[lo:hi], self.out_off[lo:hi] - np.uint64(p0), self.out_size[lo:hi],
                        self.plain[p0:p1], self.raw_dict, self.name)


def _compress_many(pieces, threads, **kw):
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(lambda p: compress(p, **kw), pieces))


def _pack(frames, pieces, plain, raw_dict=None, name=""):
    sizes = np.array([len(f) for f in frames], dtype=np.uint64)
    src_off = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
    comp = np.frombuffer(b"".join(frames), dtype=np.uint8)
    out_size = np.array([len(p) for p in pieces], dtype=np.uint64)
    out_off = np.concatenate([[0], np.cumsum(out_size)[:-1]]).astype(np.uint64)
    return FrameSet(comp, src_off, sizes, out_off, out_size, plain, raw_dict, name)


def _cached(name, builder):
    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, name + ".npz")
    if os.path.exists(path):
        try:
            z = np.load(path, allow_pickle=False)
            rd = z["raw_dict"] if "raw_dict" in z.files and z["raw_dict"].size else None
            return FrameSet(z["comp"], z["src_off"], z["src_size"], z["out_off"], z["out_size"], z["plain"], rd, name)
        except Exception:
            pass
    fs = builder()
    fs.name = name
    try:
        np.savez(path, comp=fs.comp, src_off=fs.src_off, src_size=fs.src_size, out_off=fs.out_off, out_size=fs.out_size, plain=fs.plain,
                 raw_dict=fs.raw_dict if fs.raw_dict is not None else np.zeros(0, np.uint8))
    except Exception:
        pass
    return fs


def nthreads():
    return max(1, min(32, os.cpu_count() or 1))


def c2_text_plain(total_bytes, seed=0xE90001, chunk=8 << 20):
    n = (total_bytes + chunk - 1) // chunk
    _vocab()
    with ThreadPoolExecutor(max_workers=nthreads()) as ex:
        chunks = list(ex.map(lambda i: gen_text(min(chunk, total_bytes - i * chunk), seed + 1000 * (i + 1)), range(n)))
    return np.concatenate(chunks)


def config_c2b(total_bytes=1 << 30, frame_bytes=131072, seed=0xE90001, cache=True):
    """C2b (headline): independent single-block frames of `frame_bytes` of enwik-shaped text, level 3, checksum on."""
    def build():
        plain = c2_text_plain(total_bytes, seed)
        pieces = [plain[i:i + frame_bytes] for i in range(0, total_bytes, frame_bytes)]
        frames = _compress_many(pieces, nthreads(), level=3, checksum=True)
        return _pack(frames, pieces, plain)
    name = f"c2b_{total_bytes}_{frame_bytes}_{seed:x}"
    return _cached(name, build) if cache else build()


def config_c2a(total_bytes=64 << 20, nframes=1, seed=0xE90001, cache=True):
    """C2a: the faithful enwik9.zst form -- frames of chained blocks, windowLog=17."""
    def build():
        plain = c2_text_plain(total_bytes, seed)
        per = total_bytes // nframes
        pieces = [plain[i * per:(i + 1) * per] for i in range(nframes)]
        frames = _compress_many(pieces, nthreads(), level=3, window_log=17, checksum=True)
        return _pack(frames, pieces, plain[:per * nframes])
    name = f"c2a_{total_bytes}_{nframes}_{seed:x}"
    return _cached(name, build) if cache else build()


def config_c3(nframes=10000, frame_bytes=65536, cache=True):
    """C3: frames of i.i.d. skewed bytes -> one block, 4-stream Huffman literals, almost no sequences."""
    def build():
        pieces = [gen_skewed_bytes(frame_bytes, 0xC30000 + i) for i in range(nframes)]
        frames = _compress_many(pieces, nthreads(), level=3, checksum=True)
        return _pack(frames, pieces, np.concatenate(pieces))
    name = f"c3_{nframes}_{frame_bytes}"
    return _cached(name, build) if cache else build()


def config_c4(nframes=4096, frame_bytes=1 << 20, cache=True):
    """C4: Silesia-mix-shaped 1 MiB frames, 8 chained blocks each, Raw/RLE/Compressed blocks, repeat modes."""
    def build():
        with ThreadPoolExecutor(max_workers=nthreads()) as ex:
            pieces = list(ex.map(lambda i: gen_silesia_mix(frame_bytes, 0xC40000 + i), range(nframes)))
        frames = _compress_many(pieces, nthreads(), level=3, checksum=True)
        return _pack(frames, pieces, np.concatenate(pieces))
    name = f"c4_{nframes}_{frame_bytes}"
    return _cached(name, build) if cache else build()


def config_c5(nframes=100000, dict_bytes=112640, cache=True):
    """C5: small frames built from one raw-content dictionary (no dict id in the frame header)."""
    def build():
        d = gen_text(dict_bytes, 0xD1C7)
        pieces = []
        for i in range(nframes):
            rng = np.random.Generator(np.random.PCG64(0xC50000 + i))
            size = int(np.exp(rng.uniform(np.log(1024), np.log(8192))))
            parts, have = [], 0
            fresh = gen_text(size, 0xC5F000 + i)
            fpos = 0
            while have < size:
                if rng.random() < 0.8:
                    ln = int(rng.integers(32, 513)); o = int(rng.integers(0, dict_bytes - ln)); parts.append(d[o:o + ln])
                else:
                    ln = int(rng.integers(16, 128)); parts.append(fresh[fpos:fpos + ln]); fpos += ln
                have += len(parts[-1])
            pieces.append(np.concatenate(parts)[:size])
        frames = _compress_many(pieces, nthreads(), level=3, raw_dict=d)
        return _pack(frames, pieces, np.concatenate(pieces), raw_dict=d)
    name = f"c5_{nframes}_{dict_bytes}"
    return _cached(name, build) if cache else build()

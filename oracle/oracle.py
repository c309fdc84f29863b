"""ctypes binding of oracle/ruzstd_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The class mirrors ruzstd's FrameDecoder / StreamingDecoder surface (decoding/frame_decoder.rs:154-627,
decoding/streaming_decoder.rs:45-156) so the parity tests read like the reference's own tests.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libruzstd_oracle.so")


def build(force=False):
    src = [os.path.join(_HERE, "ruzstd_oracle.c"), os.path.join(_HERE, "ruzstd_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None
READ_FN = C.CFUNCTYPE(C.c_long, C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t)
WRITE_FN = C.CFUNCTYPE(C.c_long, C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t)


class BlockTrace(C.Structure):
    _fields_ = [("block_type", C.c_uint32), ("literals_type", C.c_uint32), ("num_streams", C.c_uint32),
                ("regenerated_size", C.c_uint32), ("num_sequences", C.c_uint32), ("huf_max_bits", C.c_uint32),
                ("lit_offset", C.c_uint64), ("seq_offset", C.c_uint64), ("out_offset", C.c_uint64),
                ("out_size", C.c_uint64), ("offset_hist_after", C.c_uint32 * 3), ("pad", C.c_uint32)]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u8p, sz = C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t
        L.zo_new.restype = vp
        L.zo_free.argtypes = [vp]
        L.zo_set_max_window_size.argtypes = [vp, C.c_uint64]
        L.zo_max_window_size.argtypes = [vp]; L.zo_max_window_size.restype = C.c_uint64
        L.zo_init.argtypes = [vp, READ_FN, vp]
        L.zo_add_dict.argtypes = [vp, C.c_char_p, sz]
        L.zo_add_raw_content_dict.argtypes = [vp, C.c_uint32, C.c_char_p, sz]
        L.zo_force_dict.argtypes = [vp, C.c_uint32]
        L.zo_decode_blocks.argtypes = [vp, READ_FN, vp, C.c_int, sz, C.POINTER(C.c_int)]
        L.zo_read.argtypes = [vp, vp, sz]; L.zo_read.restype = C.c_long
        L.zo_collect_to_writer.argtypes = [vp, WRITE_FN, vp]; L.zo_collect_to_writer.restype = C.c_long
        L.zo_can_collect.argtypes = [vp]; L.zo_can_collect.restype = sz
        L.zo_is_finished.argtypes = [vp]
        L.zo_blocks_decoded.argtypes = [vp]; L.zo_blocks_decoded.restype = sz
        L.zo_bytes_read_from_source.argtypes = [vp]; L.zo_bytes_read_from_source.restype = C.c_uint64
        L.zo_content_size.argtypes = [vp]; L.zo_content_size.restype = C.c_uint64
        L.zo_window_size.argtypes = [vp]; L.zo_window_size.restype = C.c_uint64
        L.zo_frame_dict_id.argtypes = [vp, C.POINTER(C.c_uint32)]
        L.zo_get_checksum_from_data.argtypes = [vp, C.POINTER(C.c_uint32)]
        L.zo_get_calculated_checksum.argtypes = [vp, C.POINTER(C.c_uint32)]
        L.zo_decode_from_to.argtypes = [vp, C.c_char_p, sz, vp, sz, C.POINTER(sz), C.POINTER(sz)]
        L.zo_decode_all.argtypes = [vp, C.c_char_p, sz, vp, sz, C.POINTER(sz)]
        L.zo_last_error_stage.argtypes = [vp]
        L.zo_trace_enable.argtypes = [vp, C.c_int]
        L.zo_trace_num_blocks.argtypes = [vp]; L.zo_trace_num_blocks.restype = sz
        L.zo_trace_blocks.argtypes = [vp]; L.zo_trace_blocks.restype = C.POINTER(BlockTrace)
        L.zo_trace_literals.argtypes = [vp, C.POINTER(sz)]; L.zo_trace_literals.restype = u8p
        L.zo_trace_sequences.argtypes = [vp, C.POINTER(sz)]; L.zo_trace_sequences.restype = C.POINTER(C.c_uint32)
        L.zo_kat_bitreader_reversed.argtypes = [C.c_char_p, sz, C.c_char_p, sz, C.POINTER(C.c_uint64)]
        L.zo_kat_bitreader_reversed.restype = C.c_long
        L.zo_kat_bitreader_forward.argtypes = [C.c_char_p, sz, C.c_char_p, sz, C.POINTER(C.c_uint64)]
        L.zo_kat_fse_build.argtypes = [C.POINTER(C.c_int32), sz, C.c_uint8, C.c_uint8, C.POINTER(C.c_uint32)]
        L.zo_kat_fse_read.argtypes = [C.c_char_p, sz, C.c_uint8, C.c_uint8, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), sz]
        L.zo_kat_fse_read.restype = C.c_long
        L.zo_kat_huf_build.argtypes = [C.c_char_p, sz, C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), sz]
        L.zo_kat_huf_build.restype = C.c_long
        L.zo_kat_do_offset_history.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.zo_kat_do_offset_history.restype = C.c_uint32
        L.zo_kat_decode_dict.argtypes = [C.c_char_p, sz, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(sz)]
        L.zo_xxh64.argtypes = [C.c_char_p, sz]; L.zo_xxh64.restype = C.c_uint64
        L.zo_bulk_decode.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp, C.c_char_p, sz, C.c_int]
        _lib = L
    return _lib


def error_names():
    """{code: name} parsed from ruzstd_oracle.h (single source of truth for the numbering)."""
    import re
    txt = open(os.path.join(_HERE, "ruzstd_oracle.h")).read()
    return {int(v): k for k, v in re.findall(r"(ZO_(?:OK|ERR_[A-Z0-9_]+))\s*=\s*(\d+)", txt)}


class OracleError(Exception):
    def __init__(self, code, stage=0):
        self.code, self.stage = code, stage
        super().__init__(f"{error_names().get(code, code)} (stage {stage})")


def xxh64(data):
    return lib().zo_xxh64(bytes(data), len(data))


class _Reader:
    """io::Read over a bytes object or a Python file-like; keeps the ctypes callback alive."""

    def __init__(self, src):
        if isinstance(src, (bytes, bytearray, memoryview)):
            import io
            src = io.BytesIO(bytes(src))
        self.src = src

        def _cb(_user, buf, n):
            b = self.src.read(n)
            if b:
                C.memmove(buf, b, len(b))
            return len(b)
        self.cb = READ_FN(_cb)


ALL, UPTO_BLOCKS, UPTO_BYTES = 0, 1, 2


class FrameDecoder:
    """Mirror of ruzstd::decoding::FrameDecoder backed by the C oracle."""

    def __init__(self):
        self.L = lib()
        self.h = self.L.zo_new()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.zo_free(self.h); self.h = None

    def _chk(self, e):
        if e:
            raise OracleError(e, self.L.zo_last_error_stage(self.h))

    def set_max_window_size(self, n): self.L.zo_set_max_window_size(self.h, n)
    def max_window_size(self): return self.L.zo_max_window_size(self.h)

    def reset(self, reader):
        if not isinstance(reader, _Reader):
            reader = _Reader(reader)
        self._chk(self.L.zo_init(self.h, reader.cb, None))
        return reader
    init = reset

    def add_dict(self, raw): self._chk(self.L.zo_add_dict(self.h, bytes(raw), len(raw)))
    def add_raw_content_dict(self, dict_id, content): self._chk(self.L.zo_add_raw_content_dict(self.h, dict_id, bytes(content), len(content)))
    def force_dict(self, dict_id): self._chk(self.L.zo_force_dict(self.h, dict_id))

    def decode_blocks(self, reader, strategy=ALL, n=0):
        fin = C.c_int(0)
        self._chk(self.L.zo_decode_blocks(self.h, reader.cb, None, strategy, n, C.byref(fin)))
        return bool(fin.value)

    def read(self, n):
        buf = (C.c_uint8 * max(n, 1))()
        r = self.L.zo_read(self.h, buf, n)
        if r < 0:
            raise OracleError(16)
        return bytes(buf[:r])

    def collect(self):
        out = bytearray()
        while True:
            n = self.can_collect()
            if n == 0:
                break
            out += self.read(n)
        return bytes(out)

    def can_collect(self): return self.L.zo_can_collect(self.h)
    def is_finished(self): return bool(self.L.zo_is_finished(self.h))
    def blocks_decoded(self): return self.L.zo_blocks_decoded(self.h)
    def bytes_read_from_source(self): return self.L.zo_bytes_read_from_source(self.h)
    def content_size(self): return self.L.zo_content_size(self.h)
    def window_size(self): return self.L.zo_window_size(self.h)

    def get_checksum_from_data(self):
        v = C.c_uint32()
        return v.value if self.L.zo_get_checksum_from_data(self.h, C.byref(v)) else None

    def get_calculated_checksum(self):
        v = C.c_uint32()
        return v.value if self.L.zo_get_calculated_checksum(self.h, C.byref(v)) else None

    def decode_from_to(self, source, target_len):
        buf = (C.c_uint8 * max(target_len, 1))()
        r, w = C.c_size_t(), C.c_size_t()
        self._chk(self.L.zo_decode_from_to(self.h, bytes(source), len(source), buf, target_len, C.byref(r), C.byref(w)))
        return r.value, bytes(buf[:w.value])

    def decode_all(self, data, out_cap):
        buf = (C.c_uint8 * max(out_cap, 1))()
        w = C.c_size_t()
        self._chk(self.L.zo_decode_all(self.h, bytes(data), len(data), buf, out_cap, C.byref(w)))
        return bytes(buf[:w.value])

    # --- trace of intermediate results (what the CUDA kernels are checked against)
    def trace_enable(self, on=True): self.L.zo_trace_enable(self.h, 1 if on else 0)

    def trace(self):
        nb = self.L.zo_trace_num_blocks(self.h)
        bl = self.L.zo_trace_blocks(self.h)
        n = C.c_size_t()
        lp = self.L.zo_trace_literals(self.h, C.byref(n))
        lits = bytes(C.cast(lp, C.POINTER(C.c_uint8 * n.value)).contents) if n.value else b""
        sp = self.L.zo_trace_sequences(self.h, C.byref(n))
        seqs = np.ctypeslib.as_array(sp, shape=(n.value, 4)).copy() if n.value else np.zeros((0, 4), np.uint32)
        blocks = []
        for i in range(nb):
            b = bl[i]
            blocks.append({k: (list(getattr(b, k)) if k == "offset_hist_after" else getattr(b, k)) for k, _ in BlockTrace._fields_ if k != "pad"})
        return blocks, lits, seqs


class StreamingDecoder:
    """Mirror of ruzstd::decoding::StreamingDecoder (streaming_decoder.rs:45-156) over the oracle."""

    def __init__(self, source, decoder=None, max_window_size=None):
        self.decoder = decoder or FrameDecoder()
        if max_window_size is not None:
            self.decoder.set_max_window_size(max_window_size)
        self.source = _Reader(source)
        self.decoder.reset(self.source)

    def read(self, n):                                            # :118-155
        d = self.decoder
        if d.is_finished() and d.can_collect() == 0:
            return b""
        while d.can_collect() < n and not d.is_finished():
            d.decode_blocks(self.source, UPTO_BYTES, n - d.can_collect())
        return d.read(n)

    def read_to_end(self):
        out = bytearray()
        while True:
            b = self.read(1 << 16)
            if not b:
                return bytes(out)
            out += b

    def into_frame_decoder(self): return self.decoder


def decode_frame(data, dict_raw=None, raw_dict=None):
    """FrameDecoder::reset + decode_blocks(All) + collect, as tests/decode_corpus.rs:76-100 does."""
    d = FrameDecoder()
    if dict_raw is not None:
        d.add_dict(dict_raw)
    r = d.reset(data)
    if raw_dict is not None:
        d.add_raw_content_dict(1, raw_dict); d.force_dict(1)
    d.decode_blocks(r, ALL)
    return d.collect(), d


def bulk_decode(inp, in_off, in_sz, out_off, out_cap, raw_dict=None, nthreads=1, out=None):
    """Decode many independent frames (numpy arrays of offsets/sizes); returns (output ndarray, sizes)."""
    L = lib()
    inp = np.ascontiguousarray(inp, dtype=np.uint8)
    in_off = np.ascontiguousarray(in_off, dtype=np.uint64); in_sz = np.ascontiguousarray(in_sz, dtype=np.uint64)
    out_off = np.ascontiguousarray(out_off, dtype=np.uint64); out_cap = np.ascontiguousarray(out_cap, dtype=np.uint64)
    total = int((out_off + out_cap).max()) if len(out_off) else 0
    if out is None or len(out) < total:
        out = np.empty(total, dtype=np.uint8)
    out_sz = np.zeros(len(in_off), dtype=np.uint64)
    e = L.zo_bulk_decode(inp.ctypes.data, in_off.ctypes.data, in_sz.ctypes.data, len(in_off), out.ctypes.data,
                         out_off.ctypes.data, out_cap.ctypes.data, out_sz.ctypes.data,
                         bytes(raw_dict) if raw_dict is not None else None, len(raw_dict) if raw_dict is not None else 0, nthreads)
    if e:
        raise OracleError(e)
    return out, out_sz

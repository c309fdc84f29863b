"""Known-answer tests restated from the reference's in-file unit tests (SURVEY.md section 4 'Unit KATs')."""
import ctypes as C

import numpy as np


def _rev(oracle, src, counts):
    vals = (C.c_uint64 * len(counts))()
    rem = oracle.lib().zo_kat_bitreader_reversed(bytes(src), len(src), bytes(counts), len(counts), vals)
    return list(vals), rem


def test_bitreader_reversed_zero_fill_and_negative_remaining(oracle):
    """bit_io/bit_reader_reverse.rs:166-184: reading past the start returns zeros and bits_remaining goes negative."""
    vals, rem = _rev(oracle, [0b01010101, 0b11110000], [9, 9, 5])  # 16 bits available, 23 requested
    allbits = (0b11110000 << 8) | 0b01010101
    assert vals[0] == (allbits >> 7) & 0x1FF
    assert vals[1] == ((allbits & 0x7F) << 2) & 0x1FF
    assert vals[2] == 0
    assert rem == -7


def test_bitreader_reversed_matches_big_integer(oracle):
    """tests/bit_reader.rs:45-79 idea: a 128-bit constant read back to front in odd-sized pieces."""
    rng = np.random.Generator(np.random.PCG64(1))
    src = bytes(rng.integers(0, 256, 37, dtype=np.uint8))
    big = int.from_bytes(src, "little")
    counts = [1, 7, 3, 11, 9, 16, 31, 2, 25, 13, 8, 56, 5, 40, 17]
    vals, rem = _rev(oracle, src, counts)
    pos = len(src) * 8
    for c, v in zip(counts, vals):
        lo = pos - c
        expect = (big >> lo) & ((1 << c) - 1) if lo >= 0 else ((big & ((1 << pos) - 1)) << -lo if pos > 0 else 0)
        assert v == expect
        pos = lo
    assert rem == pos


def test_bitreader_forward(oracle):
    """tests/bit_reader.rs:1-43: LSB-first forward reads."""
    src = bytes([0xA5, 0x3C, 0xFF, 0x01, 0x80])
    big = int.from_bytes(src, "little")
    counts = [3, 5, 9, 1, 14, 8]
    vals = (C.c_uint64 * len(counts))()
    assert oracle.lib().zo_kat_bitreader_forward(src, len(src), bytes(counts), len(counts), vals) == 0
    pos = 0
    for c, v in zip(counts, vals):
        assert v == (big >> pos) & ((1 << c) - 1)
        pos += c
    assert oracle.lib().zo_kat_bitreader_forward(src, len(src), bytes([41]), 1, vals) != 0


def _fse(oracle, probs, log, max_symbol):
    p = (C.c_int32 * len(probs))(*probs)
    out = (C.c_uint32 * (3 << log))()
    assert oracle.lib().zo_kat_fse_build(p, len(probs), log, max_symbol, out) == 0
    return np.array(out, dtype=np.uint32).reshape(-1, 3)  # base_line, num_bits, symbol


def test_predefined_ll_table_entries(oracle):
    """sequence_section_decoder.rs:444-487 test_ll_default."""
    ll = [4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1]
    t = _fse(oracle, ll, 6, 35)
    assert len(t) == 64
    # values from the reference test: (index, symbol, num_bits, base_line)
    for idx, sym, nb, bl in [(0, 0, 4, 0), (19, 27, 6, 0), (39, 25, 4, 16), (60, 35, 6, 0), (59, 24, 5, 32)]:
        assert (t[idx][2], t[idx][1], t[idx][0]) == (sym, nb, bl), idx


def test_fse_table_is_a_valid_decoder(oracle):
    """fse/mod.rs:21-43: probabilities [0,0,-1,3,2,2,56] log 6 -- every state's [base, base + 2^nb) stays inside the table and
    each symbol's states partition the table exactly."""
    probs = [0, 0, -1, 3, 2, 2, 56]
    t = _fse(oracle, probs, 6, 255)
    cover = {}
    for bl, nb, sym in t:
        cover.setdefault(int(sym), []).append((int(bl), int(bl) + (1 << int(nb))))
    for sym, spans in cover.items():
        spans.sort()
        assert spans[0][0] == 0 and spans[-1][1] == 64
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0]
        assert len(spans) == (1 if probs[sym] == -1 else probs[sym])


def test_rep_offset_underflow(oracle):
    """sequence_execution.rs:120-133: of==3, ll==0, hist[0]==0 resolves to 0 instead of underflowing."""
    h = (C.c_uint32 * 3)(0, 4, 8)
    assert oracle.lib().zo_kat_do_offset_history(3, 0, h) == 0
    h = (C.c_uint32 * 3)(1, 4, 8)
    assert oracle.lib().zo_kat_do_offset_history(1, 5, h) == 1 and list(h) == [1, 4, 8]
    assert oracle.lib().zo_kat_do_offset_history(2, 5, h) == 4 and list(h) == [4, 1, 8]
    assert oracle.lib().zo_kat_do_offset_history(1, 0, h) == 1 and list(h) == [1, 4, 8]
    assert oracle.lib().zo_kat_do_offset_history(3, 0, h) == 0 and list(h) == [0, 1, 4]
    assert oracle.lib().zo_kat_do_offset_history(20, 0, h) == 17 and list(h) == [17, 0, 1]


DICT_KAT_TABLES = bytes([
    54, 16, 192, 155, 4, 0, 207, 59, 239, 121, 158, 116, 220, 93, 114, 229, 110, 41, 249, 95,
    165, 255, 83, 202, 254, 68, 74, 159, 63, 161, 100, 151, 137, 21, 184, 183, 189, 100, 235,
    209, 251, 174, 91, 75, 91, 185, 19, 39, 75, 146, 98, 177, 249, 14, 4, 35, 0, 0, 0, 40, 40,
    20, 10, 12, 204, 37, 196, 1, 173, 122, 0, 4, 0, 128, 1, 2, 2, 25, 32, 27, 27, 22, 24, 26,
    18, 12, 12, 15, 16, 11, 69, 37, 225, 48, 20, 12, 6, 2, 161, 80, 40, 20, 44, 137, 145, 204,
    46, 0, 0, 0, 0, 0, 116, 253, 16, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0])
DICT_KAT_CONTENT = bytes([1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 1, 1, 123, 3, 234, 23, 234, 34, 23, 234, 34, 34, 234, 234])


def dict_kat_bytes():
    """The hand-built dictionary of ruzstd/src/tests/dict_test.rs:1-64 (golden vector: id, offsets, content)."""
    return (bytes([0x37, 0xA4, 0x30, 0xEC, 0x01, 0x21, 0x23, 0x47]) + DICT_KAT_TABLES + bytes([3, 0, 0, 0, 10, 0, 0, 0, 0xEF, 0xCD, 0xAB, 0])
            + DICT_KAT_CONTENT)


def test_dictionary_header_kat(oracle):
    """dict_test.rs:1-75 / dictionary.rs:129-163: id, offset history [3,10,0xABCDEF], content; bad magic; truncation."""
    raw = dict_kat_bytes()
    i, offs, n = C.c_uint32(), (C.c_uint32 * 3)(), C.c_size_t()
    e = oracle.lib().zo_kat_decode_dict(raw, len(raw), C.byref(i), offs, C.byref(n))
    assert e == 0, oracle.error_names().get(e)
    assert i.value == 0x47232101 and list(offs) == [3, 10, 0xABCDEF] and n.value == len(DICT_KAT_CONTENT)
    bad = b"\x01\x01\x01\x01" + raw[4:]
    assert oracle.error_names()[oracle.lib().zo_kat_decode_dict(bad, len(bad), C.byref(i), offs, C.byref(n))] == "ZO_ERR_DICT_BAD_MAGIC_NUM"
    for cut in range(len(raw) - len(DICT_KAT_CONTENT)):   # truncated dictionaries must error, never crash
        assert oracle.lib().zo_kat_decode_dict(raw[:cut], cut, C.byref(i), offs, C.byref(n)) != 0


def test_xxh64_known_answers(oracle):
    assert oracle.xxh64(b"") == 0xEF46DB3751D8E999
    assert oracle.xxh64(b"a") == 0xD24EC4F1A98C6E5B
    assert oracle.xxh64(b"abc") == 0x44BC2CF5AD770999
    assert oracle.xxh64(b"Nobody inspects the spammish repetition") == 0xFBCEA83C8A378BF1

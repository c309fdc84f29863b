"""CPU tests of the execution-kernel LOGIC (tests/emul_exec_cta.py models zstd-rs_b200/csrc/exec_cta.cuh bit by bit):
the oracle's per-block traces ({ll, ml, offset_value}, literals) are turned into what k_fse hands over -- prefix-form
records with symbolic offsets -- and executed by the model; the result must be the reference's output of that block."""
import numpy as np
import pytest

from conftest import read_golden
import emul_exec_cta as E


def _run_frame(oracle, data, max_blocks=None, dict_bytes=None):
    d = oracle.FrameDecoder(); d.trace_enable()
    r = d.reset(data); d.decode_blocks(r); out = d.collect()
    blocks, lits, seqs = d.trace()
    hist = [1, 4, 8]
    pos = 0
    nblk = nseq = 0
    for tb in blocks:
        size = tb["out_size"]
        if tb["block_type"] == 2:
            n = tb["num_sequences"]
            s = seqs[tb["seq_offset"]:tb["seq_offset"] + n]
            lit = lits[tb["lit_offset"]:tb["lit_offset"] + tb["regenerated_size"]]
            sym, h_after = E.symbolic_offsets(s[:, 0], s[:, 2])
            # symbols resolve to the reference's actual offsets, and the history carries over symbolically
            got_actual = np.array([E.sym_resolve(int(v), hist) for v in sym], dtype=np.uint64)
            assert np.array_equal(got_actual, s[:, 3].astype(np.uint64))
            if max_blocks is None or nblk < max_blocks:
                rec = E.to_prefix(s[:, 0], s[:, 1], sym)
                try:
                    got = E.exec_block(rec, hist, lit, out[:pos], 0x1000 + pos + 5)
                    assert got == out[pos:pos + size], (nblk, size)
                except E.Bail:
                    pass
                nblk += 1; nseq += n
            if n:
                hist = [E.sym_resolve(int(v), hist) for v in h_after]
                assert hist == list(tb["offset_hist_after"])
        pos += size
    assert pos == len(out)
    return nblk, nseq


def test_model_on_corpus_frames(oracle, manifest):
    """a handful of decodecorpus frames (multi-block, repeat offsets, RLE/raw literals, far matches)"""
    names = sorted(manifest["corpus"])
    picked = [n for n in names if manifest["corpus"][n]["size"] < 40000][:14] + ["z000003.zst"]   # z000003: 137 small blocks, far sources into row 0
    nb = ns = 0
    for n in picked:
        b, s = _run_frame(oracle, read_golden("decodecorpus", n))
        nb += b; ns += s
    assert nb > 20 and ns > 2000


def test_model_on_synthetic(oracle):
    """libzstd level-3 frames: text-like (many short sequences), long overlapping runs, a 2-block chained frame"""
    import datagen as G
    rng = np.random.Generator(np.random.PCG64(5))
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(300)]
    text = b" ".join(words[int(i)] for i in rng.integers(0, 300, 9000))[:48000]
    runs = b"".join(bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 700)) for _ in range(60)) + b"abcabcabc" * 500 + b"xy" * 3000
    chained = (text * 4)[:150000]   # > 128 KiB: two chained blocks, matches reach into the first block
    # a block that is nearly all literals (low-entropy bytes without matches): the literals do not fit beside the window
    lowent = bytes(rng.choice(np.arange(40, 56, dtype=np.uint8), 126000, p=np.arange(1, 17) / 136.0)) + text[:4000]
    for plain, mb in ((text, None), (runs, None), (chained, None), (lowent, None)):
        frame = G.compress(np.frombuffer(plain, dtype=np.uint8), level=3)
        nb, ns = _run_frame(oracle, frame, mb)
        assert nb >= 1


def test_symbolic_history_matches_reference_rule():
    """sequence_execution.rs:125-133: of == 3, ll == 0, hist[0] == 0 -> 0 (saturating), through the symbolic form"""
    sym, h = E.symbolic_offsets([0], [3])
    assert E.sym_resolve(int(sym[0]), [0, 4, 8]) == 0
    assert E.sym_resolve(int(sym[0]), [1, 4, 8]) == 0
    assert E.sym_resolve(int(sym[0]), [7, 4, 8]) == 6
    # three decrements in a row stay saturating
    sym, h = E.symbolic_offsets([0, 0, 0], [3, 3, 3])
    assert [E.sym_resolve(int(v), [2, 9, 9]) for v in sym] == [1, 0, 0]

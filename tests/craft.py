"""Hand-crafted zstd frames for edge cases the reference accepts but a stock encoder never produces (test helpers).

`irregular_huffman_split(frame)` takes a single-block frame whose literals are 4-stream Huffman-compressed and re-encodes the four
streams with an UNEVEN split of the literals.  ruzstd decodes each stream until its bits run out and only checks the total
(literals_section_decoder.rs:94-122,150-155), so it accepts such a frame and yields the same plaintext; the GPU
fast path (which assumes the standard (n+3)/4 split) must detect the anomaly and replay the block with the exact semantics.
"""
import ctypes as C


def _parse_single_block(frame):
    assert frame[:4] == bytes([0x28, 0xB5, 0x2F, 0xFD])
    desc = frame[4]
    single = (desc >> 5) & 1
    pos = 5 + (0 if single else 1) + [0, 1, 2, 4][desc & 3]
    fl = desc >> 6
    pos += ([1 if single else 0, 2, 4, 8][fl])
    bh = int.from_bytes(frame[pos:pos + 3], "little")
    assert bh & 1 and ((bh >> 1) & 3) == 2, "need a single, last, compressed block"
    size = bh >> 3
    return pos, size


def irregular_huffman_split(oracle, frame, shift=7):
    hdr_end, bsize = _parse_single_block(frame)
    blk = frame[hdr_end + 3: hdr_end + 3 + bsize]
    b0 = blk[0]
    assert (b0 & 3) == 2 and ((b0 >> 2) & 3) >= 1, "need Compressed literals with 4 streams"
    sf = (b0 >> 2) & 3
    if sf == 1:
        regen = (b0 >> 4) + ((blk[1] & 0x3F) << 4); comp = (blk[1] >> 6) + (blk[2] << 2); lh = 3
    elif sf == 2:
        regen = (b0 >> 4) + (blk[1] << 4) + ((blk[2] & 3) << 12); comp = (blk[2] >> 2) + (blk[3] << 6); lh = 4
    else:
        regen = (b0 >> 4) + (blk[1] << 4) + ((blk[2] & 0x3F) << 12); comp = (blk[2] >> 6) + (blk[3] << 2) + (blk[4] << 10); lh = 5
    lit_payload = blk[lh:lh + comp]
    rest = blk[lh + comp:]
    # Huffman table from the tree description (through the oracle's table builder)
    L = oracle.lib()
    mb = C.c_uint8()
    ent = (C.c_uint16 * 2048)()
    used = L.zo_kat_huf_build(bytes(lit_payload), len(lit_payload), C.byref(mb), ent, 2048)
    assert used > 0
    mb = mb.value
    code = {}
    for idx in range(1 << mb):
        sym, nb = ent[idx] & 0xFF, ent[idx] >> 8
        if sym not in code:
            code[sym] = (idx >> (mb - nb), nb)
    # literals of the block (oracle trace)
    d = oracle.FrameDecoder(); d.trace_enable()
    r = d.reset(frame); d.decode_blocks(r); plain = d.collect()
    blocks, lits, _ = d.trace()
    lit = lits[blocks[0]["lit_offset"]:blocks[0]["lit_offset"] + regen]
    assert len(lit) == regen
    S = (regen + 3) // 4
    cuts = [0, S + shift, 2 * S - shift, 3 * S + 2 * shift, regen]   # uneven, still ascending
    streams = []
    for a, b in zip(cuts, cuts[1:]):
        val, nbits = 1, 0
        for s in lit[a:b]:
            c, n = code[s]
            val = (val << n) | c
            nbits += n
        streams.append(val.to_bytes((nbits + 1 + 7) // 8, "little"))
    assert all(len(s) < 65536 for s in streams[:3])
    jump = b"".join(len(s).to_bytes(2, "little") for s in streams[:3])
    new_payload = bytes(lit_payload[:used]) + jump + b"".join(streams)
    ncomp = len(new_payload)
    # literals header, size format 3 (18-bit sizes)
    h = (2) | (3 << 2) | ((regen & 0xF) << 4)
    hdr = bytes([h & 0xFF, (regen >> 4) & 0xFF, ((regen >> 12) & 0x3F) | ((ncomp & 3) << 6), (ncomp >> 2) & 0xFF, (ncomp >> 10) & 0xFF])
    nblk = hdr + new_payload + bytes(rest)
    bh = (1) | (2 << 1) | (len(nblk) << 3)
    out = bytes(frame[:hdr_end]) + bh.to_bytes(3, "little") + nblk + bytes(frame[hdr_end + 3 + bsize:])
    return out, plain

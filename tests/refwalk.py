"""What a host that keeps the reference's own header parsing would hand to b200z_decode_blocks_batch: a Python restatement of
read_frame_header (decoding/frame.rs:6-85,116-139), read_block_header (decoding/block_decoder.rs:201-283),
LiteralsSection::parse_from_header (blocks/literals_section.rs:117-223) and SequencesHeader::parse_from_header
(blocks/sequence_section.rs:108-167), valid frames only (test helper)."""
import numpy as np


def frame_header(data):
    assert data[:4] == bytes([0x28, 0xB5, 0x2F, 0xFD])
    desc = data[4]
    single = (desc >> 5) & 1
    pos = 5
    window = None
    if not single:
        wd = data[pos]; pos += 1
        exp, mant = wd >> 3, wd & 7
        base = 1 << (10 + exp)
        window = base + (base // 8) * mant
    dlen = [0, 1, 2, 4][desc & 3]
    pos += dlen
    flen = [1 if single else 0, 2, 4, 8][desc >> 6]
    fcs = int.from_bytes(data[pos:pos + flen], "little") if flen else None
    if flen == 2:
        fcs += 256
    pos += flen
    if single:
        window = fcs
    return pos, window, (desc >> 2) & 1


def literals_header(c):
    b0 = c[0]
    lt, sf = b0 & 3, (b0 >> 2) & 3
    if lt in (0, 1):
        if sf in (0, 2):
            regen, n = b0 >> 3, 1
        elif sf == 1:
            regen, n = (b0 >> 4) + (c[1] << 4), 2
        else:
            regen, n = (b0 >> 4) + (c[1] << 4) + (c[2] << 12), 3
        return lt, regen, 0, 0, n, (1 if lt == 1 else regen)
    streams = 1 if sf == 0 else 4
    if sf <= 1:
        regen, comp, n = (b0 >> 4) + ((c[1] & 0x3F) << 4), (c[1] >> 6) + (c[2] << 2), 3
    elif sf == 2:
        regen, comp, n = (b0 >> 4) + (c[1] << 4) + ((c[2] & 3) << 12), (c[2] >> 2) + (c[3] << 6), 4
    else:
        regen, comp, n = (b0 >> 4) + (c[1] << 4) + ((c[2] & 0x3F) << 12), (c[2] >> 6) + (c[3] << 2) + (c[4] << 10), 5
    return lt, regen, comp, streams, n, comp


def sequences_header(s):
    if s[0] == 0:
        return 0, 0
    if s[0] < 128:
        return s[0], s[1]
    if s[0] < 255:
        return ((s[0] - 128) << 8) + s[1], s[2]
    return s[1] + (s[2] << 8) + 0x7F00, s[3]


def walk(frames, out_offs, out_caps, dtypes):
    """frames: list of bytes; returns (compressed bytes, block descriptor array, frame array) for decode_blocks"""
    BD, BF = dtypes
    blocks, fr = [], []
    base = 0
    for data, oo, oc in zip(frames, out_offs, out_caps):
        pos, window, has_chk = frame_header(data)
        first = len(blocks)
        while True:
            bh = int.from_bytes(data[pos:pos + 3], "little"); pos += 3
            last, bt, size = bh & 1, (bh >> 1) & 3, bh >> 3
            content = 1 if bt == 1 else size
            d = np.zeros(1, dtype=BD)[0]
            d["src_off"], d["content_size"], d["block_type"], d["last_block"] = base + pos, content, bt, last
            d["decompressed_size"] = size if bt in (0, 1) else 0
            if bt == 2:
                c = data[pos:pos + size]
                lt, regen, comp, streams, n, payload = literals_header(c)
                d["literals_type"], d["regenerated_size"], d["compressed_size"], d["num_streams"] = lt, regen, comp, streams
                nseq, modes = sequences_header(c[n + payload:])
                d["num_sequences"], d["modes"] = nseq, modes
            blocks.append(d)
            pos += content
            if last:
                break
        f = np.zeros(1, dtype=BF)[0]
        f["out_off"], f["out_cap"], f["window_size"], f["first_block"], f["num_blocks"] = oo, oc, window, first, len(blocks) - first
        fr.append(f)
        base += len(data)
    return np.frombuffer(b"".join(frames), dtype=np.uint8), np.array(blocks, dtype=BD), np.array(fr, dtype=BF)

// plan.cpp -- host-side framing (see plan.h for the reference functions each piece mirrors).
#include "plan.h"

#include <string.h>

#include <algorithm>
#include <utility>

namespace b200z {

static inline uint32_t rd32le(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint64_t rd64le(const uint8_t *p) { return (uint64_t)rd32le(p) | ((uint64_t)rd32le(p + 4) << 32); }

// frame.rs:6-85.  Each read_exact that would fail maps to the matching *ReadError leaf.
int parse_frame_header(const uint8_t *p, size_t len, FrameHeader &h, uint32_t &skip_len, size_t &consumed) {
    consumed = 0;
    if (len < 4) { consumed = len; return B200Z_ERR_MAGIC_NUMBER_READ; }
    uint32_t magic = rd32le(p);
    size_t pos = 4;
    if (magic >= 0x184D2A50u && magic <= 0x184D2A5Fu) {
        if (len < 8) { consumed = len; return B200Z_ERR_FRAME_DESCRIPTOR_READ; }
        skip_len = rd32le(p + 4);
        consumed = 8;
        return B200Z_ERR_SKIP_FRAME;
    }
    if (magic != 0xFD2FB528u) { consumed = 4; return B200Z_ERR_BAD_MAGIC_NUMBER; }
    if (len < pos + 1) { consumed = len; return B200Z_ERR_FRAME_DESCRIPTOR_READ; }
    h = FrameHeader();
    h.descriptor = p[pos++];
    if (!h.single_segment()) {
        if (len < pos + 1) { consumed = len; return B200Z_ERR_WINDOW_DESCRIPTOR_READ; }
        h.window_descriptor = p[pos++];
    }
    static const uint8_t did_len[4] = {0, 1, 2, 4};
    uint32_t dl = did_len[h.descriptor & 3];
    if (dl) {
        if (len < pos + dl) { consumed = len; return B200Z_ERR_DICTIONARY_ID_READ; }
        uint32_t id = 0;
        for (uint32_t i = 0; i < dl; i++) id += (uint32_t)p[pos + i] << (8 * i);
        pos += dl;
        if (id != 0) { h.has_dict_id = true; h.dict_id = id; }
    }
    uint32_t flag = h.descriptor >> 6;
    uint32_t fl = flag == 0 ? (h.single_segment() ? 1 : 0) : (flag == 1 ? 2 : (flag == 2 ? 4 : 8));
    if (fl) {
        if (len < pos + fl) { consumed = len; return B200Z_ERR_FRAME_CONTENT_SIZE_READ; }
        uint64_t fcs = 0;
        for (uint32_t i = 0; i < fl; i++) fcs += (uint64_t)p[pos + i] << (8 * i);
        if (fl == 2) fcs += 256;
        h.frame_content_size = fcs;
        pos += fl;
    }
    h.header_size = (uint8_t)pos;
    consumed = pos;
    return 0;
}

// frame.rs:116-139
int frame_window_size(const FrameHeader &h, uint64_t &out) {
    if (h.single_segment()) { out = h.frame_content_size; return 0; }
    uint64_t exp = h.window_descriptor >> 3, mant = h.window_descriptor & 7;
    uint64_t base = 1ull << (10 + exp);
    uint64_t w = base + (base / 8) * mant;
    const uint64_t MAXW = (1ull << 41) + 7 * (1ull << 38);
    if (w >= 1024) { if (w < MAXW) { out = w; return 0; } return B200Z_ERR_WINDOW_TOO_BIG; }
    return B200Z_ERR_WINDOW_TOO_SMALL;
}

// block_decoder.rs:201-283
int parse_block_header(const uint8_t b[3], BlockHeader &h) {
    uint32_t t = (b[0] >> 1) & 3;
    if (t == 3) return B200Z_ERR_FOUND_RESERVED_BLOCK;
    uint32_t size = (uint32_t)(b[0] >> 3) | ((uint32_t)b[1] << 5) | ((uint32_t)b[2] << 13);
    if (size > 128u * 1024u) return B200Z_ERR_BLOCK_SIZE_TOO_LARGE;
    h.type = t;
    h.decompressed_size = (t == BT_RAW || t == BT_RLE) ? size : 0;
    h.content_size = t == BT_RLE ? 1 : size;
    h.last = b[0] & 1;
    return 0;
}

static TabRef one_mode(uint32_t mode, TabRef cur, int32_t slot_idx) {
    TabRef r;
    switch (mode) {
        case MODE_PREDEFINED: r.kind = TabRef::PREDEF; return r;
        case MODE_RLE:
        case MODE_FSE: r.kind = TabRef::SLOT; r.idx = (uint32_t)slot_idx; return r;
        default: return cur;  // Repeat: whatever was last (table or RLE byte), sequence_section_decoder.rs:333-336
    }
}

void plan_compressed_block(const uint8_t *c, uint32_t size, BlockDesc &d, BlockRefs &r, TableCursor &cur,
                           uint32_t &n_huf_slots, uint32_t &n_fse_slots, uint64_t &lit_bytes, uint64_t &nseq_total) {
    plan_compressed_block_view(c, nullptr, size, d, r, cur, n_huf_slots, n_fse_slots, lit_bytes, nseq_total);
}

// `c`: the first bytes of the block content (the literals section header: at most 5 are read); `seq`: the first bytes of the sequences
// section header (at most 4 are read), or null when `c` is the whole content.  The device-side header walk (k_walk) hands over
// exactly these bytes per block.
void plan_compressed_block_view(const uint8_t *c, const uint8_t *seq, uint32_t size, BlockDesc &d, BlockRefs &r, TableCursor &cur,
                                uint32_t &n_huf_slots, uint32_t &n_fse_slots, uint64_t &lit_bytes, uint64_t &nseq_total) {
    // ---- literals section header (literals_section.rs:117-223)
    if (size == 0) { d.host_status = host_status(B200Z_ERR_LITSEC_GET_BITS, B200Z_STAGE_BLOCK_BODY, 1); return; }
    uint32_t lt = c[0] & 3, sf = (c[0] >> 2) & 3, need;
    if (lt == LT_RAW || lt == LT_RLE) need = (sf == 0 || sf == 2) ? 1 : (sf == 1 ? 2 : 3);
    else need = sf <= 1 ? 3 : (sf == 2 ? 4 : 5);
    if (size < need) { d.host_status = host_status(B200Z_ERR_LITSEC_NOT_ENOUGH_BYTES, B200Z_STAGE_BLOCK_BODY, 1); return; }
    uint32_t regen = 0, comp = 0, nstreams = 0;
    if (lt == LT_RAW || lt == LT_RLE) {
        if (sf == 0 || sf == 2) regen = c[0] >> 3;
        else if (sf == 1) regen = (c[0] >> 4) + ((uint32_t)c[1] << 4);
        else regen = (c[0] >> 4) + ((uint32_t)c[1] << 4) + ((uint32_t)c[2] << 12);
    } else {
        nstreams = sf == 0 ? 1 : 4;
        if (sf <= 1) { regen = (c[0] >> 4) + (((uint32_t)c[1] & 0x3f) << 4); comp = (c[1] >> 6) + ((uint32_t)c[2] << 2); }
        else if (sf == 2) { regen = (c[0] >> 4) + ((uint32_t)c[1] << 4) + (((uint32_t)c[2] & 3) << 12); comp = (c[2] >> 2) + ((uint32_t)c[3] << 6); }
        else { regen = (c[0] >> 4) + ((uint32_t)c[1] << 4) + (((uint32_t)c[2] & 0x3f) << 12); comp = (c[2] >> 6) + ((uint32_t)c[3] << 2) + ((uint32_t)c[4] << 10); }
    }
    d.lit_type = lt; d.nstreams = nstreams; d.regen_size = regen; d.lit_comp_size = comp; d.lit_off = need;
    // block_decoder.rs:120-134
    uint32_t upper = (lt == LT_COMPRESSED || lt == LT_TREELESS) ? comp : (lt == LT_RLE ? 1 : regen);
    if (size - need < upper) { d.host_status = host_status(B200Z_ERR_MALFORMED_SECTION_HEADER, B200Z_STAGE_BLOCK_BODY, 1); return; }
    if (lt == LT_COMPRESSED) {
        r.build_huf = (int32_t)n_huf_slots++;
        cur.huf.kind = TabRef::SLOT; cur.huf.idx = (uint32_t)r.build_huf;
        r.huf = cur.huf;
    } else if (lt == LT_TREELESS) r.huf = cur.huf;
    if (lt == LT_COMPRESSED || lt == LT_TREELESS) {
        d.lit_buf_off = lit_bytes;
        lit_bytes += ((uint64_t)regen + 31) & ~15ull;  // 16-byte aligned, >= 16 bytes of slack for vector stores
    }
    // ---- sequences section header (sequence_section.rs:108-167)
    const uint8_t *s = seq ? seq : c + need + upper;
    uint32_t rem = size - need - upper, hdr = 0, nseq = 0, modes = 0;
    bool bad = false;
    if (rem == 0) bad = true;
    else if (s[0] == 0) { hdr = 1; }
    else if (s[0] < 128) { if (rem < 2) bad = true; else { nseq = s[0]; modes = s[1]; hdr = 2; } }
    else if (s[0] < 255) {
        if (rem < 2) bad = true;
        else {
            nseq = (((uint32_t)s[0] - 128) << 8) + s[1]; hdr = 2;
            if (nseq != 0) { if (rem < 3) bad = true; else { modes = s[2]; hdr = 3; } }
        }
    } else { if (rem < 4) bad = true; else { nseq = (uint32_t)s[1] + ((uint32_t)s[2] << 8) + 0x7F00; modes = s[3]; hdr = 4; } }
    if (bad) { d.host_status = host_status(B200Z_ERR_SEQHDR_NOT_ENOUGH_BYTES, B200Z_STAGE_BLOCK_BODY, 2); return; }
    d.nseq = nseq; d.modes = modes; d.seq_off = need + upper + hdr;
    if (nseq == 0) {
        // block_decoder.rs:185-191: bytes after an empty sequences section
        if (rem - hdr != 0) d.host_status = host_status(B200Z_ERR_SEQ_EXTRA_BITS, B200Z_STAGE_SEQUENCES, 2);
        return;
    }
    uint32_t ml_ = (modes >> 6) & 3, mo = (modes >> 4) & 3, mm = (modes >> 2) & 3;
    bool needs_slot = ml_ == MODE_RLE || ml_ == MODE_FSE || mo == MODE_RLE || mo == MODE_FSE || mm == MODE_RLE || mm == MODE_FSE;
    if (needs_slot) r.build_fse = (int32_t)n_fse_slots++;
    cur.ll = one_mode(ml_, cur.ll, r.build_fse);
    cur.of = one_mode(mo, cur.of, r.build_fse);
    cur.ml = one_mode(mm, cur.ml, r.build_fse);
    r.ll = cur.ll; r.of = cur.of; r.ml = cur.ml;
    d.seq_buf_off = nseq_total;
    nseq_total += (nseq + 3) & ~3ull;   // keeps every block's sequence array 16-byte aligned (3 x u32 x 4)
}

// ---- XXH64 ------------------------------------------------------------------------------------------------
static const uint64_t P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full, P3 = 0x165667B19E3779F9ull, P4 = 0x85EBCA77C2B2AE63ull, P5 = 0x27D4EB2F165667C5ull;
static inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rnd(uint64_t acc, uint64_t in) { acc += in * P2; acc = rotl(acc, 31); return acc * P1; }
static inline uint64_t mrg(uint64_t acc, uint64_t v) { v = rnd(0, v); acc ^= v; return acc * P1 + P4; }
void XXH64State::reset() { v[0] = P1 + P2; v[1] = P2; v[2] = 0; v[3] = 0ull - P1; total = 0; memsize = 0; }
void XXH64State::update(const uint8_t *p, size_t len) {
    total += len;
    if (memsize + len < 32) { if (len) memcpy(mem + memsize, p, len); memsize += (uint32_t)len; return; }
    const uint8_t *end = p + len;
    if (memsize) {
        size_t fill = 32 - memsize;
        memcpy(mem + memsize, p, fill);
        for (int i = 0; i < 4; i++) v[i] = rnd(v[i], rd64le(mem + 8 * i));
        p += fill; memsize = 0;
    }
    while (p + 32 <= end) { for (int i = 0; i < 4; i++) v[i] = rnd(v[i], rd64le(p + 8 * i)); p += 32; }
    if (p < end) { memcpy(mem, p, (size_t)(end - p)); memsize = (uint32_t)(end - p); }
}
uint64_t XXH64State::digest() const {
    uint64_t h;
    if (total >= 32) { h = rotl(v[0], 1) + rotl(v[1], 7) + rotl(v[2], 12) + rotl(v[3], 18); for (int i = 0; i < 4; i++) h = mrg(h, v[i]); }
    else h = v[2] + P5;
    h += total;
    const uint8_t *p = mem, *end = mem + memsize;
    while (p + 8 <= end) { h ^= rnd(0, rd64le(p)); h = rotl(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= end) { h ^= (uint64_t)rd32le(p) * P1; h = rotl(h, 23) * P2 + P3; p += 4; }
    while (p < end) { h ^= (*p) * P5; h = rotl(h, 11) * P1; p++; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

// ---------------------------------------------------------------------------------------------------------------
// Scheduling decisions (plan.h).  k_exec_cta puts a whole CTA on a frame's chain of blocks (~15x faster per frame), k_exec one warp --
// but k_exec keeps ~4,700 frames in flight and needs fewer instructions per byte, so it wins once there are enough frames to fill
// the machine.  Measured, 1 MiB Silesia-mix frames of 8 blocks, execution kernel alone: 512 frames 4.0 ms with CTAs / 8.7 ms with
// warps, 2048 frames 14.3 / 10.2 ms, 4096 frames 28.0 / 12.5 ms; 64 chained 16 MiB text frames 17.4 ms with CTAs.  Model, fitted to
// those:   t_cta  = max(sum w / SMs, largest w) * kCtaMs            (one CTA per frame, a frame at a time)
//          t_warp = largest w * kWarpChainMs + sum w * kWarpMs      (the longest chain, slowed down by everything else in flight)
// The eligible frames are sorted by work; the k largest would go to CTAs and the rest to warps (the two kernels run one after the
// other: each wants whole SMs).  A real split is taken only when the model promises a clear gain over both pure choices (measured on
// 4096 similar frames: the 171 largest on CTAs shortened k_exec by 0.9 ms and cost 2.4 ms of k_exec_cta).
// ---------------------------------------------------------------------------------------------------------------
void route_exec_frames(const uint64_t *work, const uint8_t *eligible, size_t nframes, uint32_t sms_u, std::vector<uint32_t> &cta_frames) {
    cta_frames.clear();
    std::vector<std::pair<uint64_t, uint32_t>> cand;   // (work, frame)
    uint64_t warp_only_work = 0;
    for (size_t f = 0; f < nframes; f++) {
        if (eligible[f]) cand.emplace_back(work[f], (uint32_t)f);
        else warp_only_work += work[f];
    }
    if (cand.empty()) return;
    constexpr double kCtaMs = 1.06e-5, kWarpChainMs = 5.8e-5, kWarpMs = 1.07e-8;
    const double sms = (double)std::max<uint32_t>(1u, sms_u);
    std::sort(cand.begin(), cand.end(), [](const std::pair<uint64_t, uint32_t> &x, const std::pair<uint64_t, uint32_t> &y) {
        return x.first != y.first ? x.first > y.first : x.second < y.second;
    });
    std::vector<uint64_t> suffix(cand.size() + 1, 0);
    for (size_t i = cand.size(); i-- > 0;) suffix[i] = suffix[i + 1] + cand[i].first;
    size_t best_k = 0;
    double best_t = 0, t_none = 0, t_all = 0;
    uint64_t prefix = 0;
    for (size_t k = 0; k <= cand.size(); k++) {   // the k largest on CTAs
        const double t_cta = k ? std::max((double)prefix / sms, (double)cand[0].first) * kCtaMs : 0.0;
        const double t_warp = (k < cand.size() ? (double)cand[k].first * kWarpChainMs : 0.0) + (double)(suffix[k] + warp_only_work) * kWarpMs;
        if (k == 0) t_none = t_cta + t_warp;
        if (k == cand.size()) t_all = t_cta + t_warp;
        if (k == 0 || t_cta + t_warp < best_t) { best_t = t_cta + t_warp; best_k = k; }
        if (k < cand.size()) prefix += cand[k].first;
    }
    if (best_k != 0 && best_k != cand.size() && best_t > 0.7 * std::min(t_none, t_all)) best_k = t_none <= t_all ? 0 : cand.size();
    for (size_t k = 0; k < best_k; k++) cta_frames.push_back(cand[k].second);
}

// k_exec runs beside k_fse and walks every frame's blocks in order: in descriptor order (frame after frame) the frames at the end of
// the list sit idle until k_fse's last wave and the 16 chains of a k_fse warp are 2 frames x 8 blocks of very different lengths.
// Rows -- block 0 of every frame, then block 1, ... -- sorted by sequence count keep every chain fed and a warp's chains alike
// (a warp takes as long as its longest chain).
void build_fse_order(const uint32_t *first_block, const uint32_t *nblocks, const uint8_t *on_cta, size_t nframes, const uint32_t *nseq,
                     size_t nblocks_total, std::vector<uint32_t> &order) {
    order.clear();
    uint32_t maxb = 0;
    size_t covered = 0;
    bool want = false;
    for (size_t f = 0; f < nframes; f++) {
        covered += nblocks[f];
        if (on_cta[f]) continue;
        maxb = std::max(maxb, nblocks[f]);
        if (nblocks[f] >= 2) want = true;
    }
    if (!want || covered != nblocks_total) return;
    std::vector<uint8_t> seen(nblocks_total, 0);   // the frames must cover every block exactly once
    for (size_t f = 0; f < nframes; f++)
        for (uint32_t k = 0; k < nblocks[f]; k++) {
            const uint64_t b = (uint64_t)first_block[f] + k;
            if (b >= nblocks_total || seen[b]) return;
            seen[b] = 1;
        }
    std::vector<uint32_t> start(maxb + 2, 0);   // start[bi + 1] = blocks with block-in-frame index bi among the warp kernel's frames
    for (size_t f = 0; f < nframes; f++)
        if (!on_cta[f]) for (uint32_t k = 0; k < nblocks[f]; k++) start[k + 1]++;
    for (uint32_t k = 0; k <= maxb; k++) start[k + 1] += start[k];
    uint32_t tail = start[maxb + 1];   // k_exec_cta's frames follow, frame after frame
    order.assign(nblocks_total, 0u);
    for (size_t f = 0; f < nframes; f++)
        for (uint32_t k = 0; k < nblocks[f]; k++) {
            if (on_cta[f]) order[tail++] = first_block[f] + k;
            else order[start[k]++] = first_block[f] + k;
        }
    uint32_t lo = 0;
    for (uint32_t k = 0; k < maxb; k++) {
        const uint32_t hi = start[k];   // (start[k] has advanced to the end of row k)
        std::sort(order.begin() + lo, order.begin() + hi, [&](uint32_t x, uint32_t y) { return nseq[x] != nseq[y] ? nseq[x] > nseq[y] : x < y; });
        lo = hi;
    }
}

}  // namespace b200z

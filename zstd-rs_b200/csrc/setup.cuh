// setup.cuh -- device-only, warp-cooperative expansion of entropy tables (one warp per zstd block).
//
// Same results as the serial builders in tables.cuh (which restate FSETable::build_decoding_table,
// fse_decoder.rs:141-220, and HuffmanTable::build_table_from_weights, huff0_decoder.rs:284-377), reorganised so
// that the O(table size) parts run on 32 lanes out of shared memory:
//   * symbol spreading: the k-th placed cell is the k-th position of the walk p -> (p + step) & mask that lies
//     below the "less than one" region; a ballot/popc scan over the walk gives k for every position at once;
//   * state numbering: a cell's state number is its rank among the cells of the same symbol in index order --
//     __match_any_sync inside each 32-cell chunk plus a per-symbol running count;
//   * Huffman ranges: per-length start indices from ballot counts, every symbol fills its own range.
// The bit-serial parts (probability parsing, FSE-compressed Huffman weights) stay on lane 0; they touch a few
// hundred bits per block.
#pragma once
#include "tables.cuh"

namespace b200z {

struct alignas(16) SetupScratch {  // per warp, shared memory
    uint8_t nb4[HUF_TABLE_ENTRIES / 2];  // Huffman code lengths staged here, then copied out coalesced
    int16_t probs[256];
    uint8_t weights[264];
    uint8_t sym_of_rank[FSE_MAX_ENTRIES];
    uint8_t cell_sym[FSE_MAX_ENTRIES];
    uint16_t count[64];
    uint16_t start[264];           // Huffman: first table index of each symbol
    uint32_t wtab[64];             // FSE table of the compressed Huffman weights (lane 0)
    uint16_t wcount[256];          // its state counters
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// probs[0..nprobs) in shared memory, nprobs <= 64, 5 <= log <= 9; writes 1 << log compact entries to `tab` (global).
// Called by a full warp.  Returns 0 or an error (uniform across the warp).
__device__ int fse_build_warp(SetupScratch &sc, uint32_t nprobs, uint32_t log, uint32_t max_symbol, FseTab *tab) {
    const uint32_t lane = lane_id(), size = 1u << log, lt = lanemask_lt();
    if (nprobs > max_symbol + 1) return B200Z_ERR_FSE_TOO_MANY_SYMBOLS;
    // ---- ranks: positive symbols get consecutive rank ranges in symbol order; -1 symbols go to the top, in order
    uint32_t total = 0, nneg = 0;
    for (uint32_t base = 0; base < nprobs; base += 32) {
        uint32_t s = base + lane;
        int32_t p = s < nprobs ? sc.probs[s] : 0;
        uint32_t negm = __ballot_sync(0xffffffffu, p == -1);
        if (p == -1) {
            uint32_t r = nneg + __popc(negm & lt);
            if (r < size) sc.cell_sym[size - 1 - r] = (uint8_t)s;
        }
        nneg += __popc(negm);
        uint32_t cnt = p > 0 ? (uint32_t)p : 0u, incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += v; }
        uint32_t first = total + incl - cnt;
        // short runs by their own lane, the rest of a long run (a dominant symbol owns hundreds of cells) by the whole warp
        const uint32_t own = cnt < 8u ? cnt : 8u;
        for (uint32_t k = 0; k < own && first + k < size; k++) sc.sym_of_rank[first + k] = (uint8_t)s;
        uint32_t big = __ballot_sync(0xffffffffu, cnt > 8u);
        while (big) {
            const int j = __ffs((int)big) - 1;
            big &= big - 1;
            const uint32_t f = __shfl_sync(0xffffffffu, first, j), c = __shfl_sync(0xffffffffu, cnt, j);
            for (uint32_t k = 8u + lane; k < c && f + k < size; k += 32) sc.sym_of_rank[f + k] = (uint8_t)(base + (uint32_t)j);
        }
        total += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (nneg > size || total + nneg != size) return B200Z_ERR_REFERENCE_WOULD_PANIC;  // cannot happen: the parser checked the sum
    const uint32_t negative_idx = size - nneg;
    __syncwarp();
    // ---- spreading (fse_decoder.rs:178-197)
    const uint32_t step = (size >> 1) + (size >> 3) + 3, msk = size - 1;
    uint32_t placed = 0;
    for (uint32_t w = 0; w < size; w += 32) {
        uint32_t pos = ((w + lane) * step) & msk;
        bool valid = pos < negative_idx;
        uint32_t vm = __ballot_sync(0xffffffffu, valid);
        if (valid) sc.cell_sym[pos] = sc.sym_of_rank[placed + __popc(vm & lt)];
        placed += __popc(vm);
    }
    if (lane < 32) { sc.count[lane] = 0; sc.count[lane + 32] = 0; }
    __syncwarp();
    // ---- state numbers, baselines, num_bits (fse_decoder.rs:200-218, 340-366)
    for (uint32_t c = 0; c < size; c += 32) {
        uint32_t i = c + lane;
        uint32_t sym = sc.cell_sym[i];
        uint16_t ent;
        bool live = i < negative_idx;
        uint32_t key = live ? sym : 0xffffu;
        uint32_t m = __match_any_sync(0xffffffffu, key);
        uint32_t cnt = 0;
        if (live) cnt = sc.count[sym & 63u] + __popc(m & lt);
        __syncwarp();
        if (live && (m & lt) == 0) sc.count[sym & 63u] += (uint16_t)__popc(m);
        __syncwarp();
        if (live) {
            uint32_t prob = (uint32_t)(int32_t)sc.probs[sym], bl = 0, nb = 0;
            uint32_t h = hbs(prob);
            uint32_t slices = ((1u << (h - 1)) == prob) ? prob : (1u << h);
            uint32_t n_double = slices - prob, n_single = prob - n_double;
            uint32_t width = size / slices;
            uint32_t b = hbs(width) - 1;
            if (cnt < n_double) { bl = n_single * width + cnt * width * 2; nb = b + 1; }
            else { bl = (cnt - n_double) * width; nb = b; }
            ent = fse_pack16(log, bl, nb, sym);
        } else ent = fse_pack16(log, 0, log, sym);
        tab->e[i] = ent;
    }
    if (lane == 0) { tab->log = log; tab->valid = 1; tab->is_rle = 0; }
    return 0;
}

// weights[0..nweights) in shared memory -> HufSlot in global memory.  Full warp; uniform return value.
__device__ int huf_build_warp(SetupScratch &sc, uint32_t nweights, HufSlot *slot) {
    const uint32_t lane = lane_id(), lt = lanemask_lt();
    // weight checks + sum (huff0_decoder.rs:290-296)
    uint32_t sum = 0, bad = 0;
    for (uint32_t base = 0; base < nweights; base += 32) {
        uint32_t w = base + lane < nweights ? sc.weights[base + lane] : 0;
        uint32_t bm = __ballot_sync(0xffffffffu, w > HUF_MAX_BITS);
        if (bm && !bad) bad = 1;
        sum += w > 0 && w <= HUF_MAX_BITS ? 1u << (w - 1) : 0u;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (bad) return B200Z_ERR_HUF_WEIGHT_BIGGER_THAN_MAX_NUM_BITS;
    if (sum == 0) return B200Z_ERR_HUF_MISSING_WEIGHTS;
    const uint32_t max_bits = hbs(sum);
    const uint32_t left_over = (1u << max_bits) - sum;
    if (left_over == 0 || (left_over & (left_over - 1))) return B200Z_ERR_HUF_LEFTOVER_NOT_POWER_OF_2;
    const uint32_t last_weight = hbs(left_over);
    if (max_bits > HUF_MAX_BITS) return B200Z_ERR_HUF_MAX_BITS_TOO_HIGH;
    if (lane == 0) sc.weights[nweights] = (uint8_t)last_weight;   // the implicit last symbol
    __syncwarp();
    const uint32_t nsym = nweights + 1;
    // symbols per code length; lane b (1..11) owns length b
    uint32_t my_count = 0;
    for (uint32_t base = 0; base < nsym; base += 32) {
        uint32_t w = base + lane < nsym ? sc.weights[base + lane] : 0;
        uint32_t bits = w ? max_bits + 1 - w : 0;
        for (uint32_t b = 1; b <= max_bits; b++) {
            uint32_t c = __popc(__ballot_sync(0xffffffffu, bits == b));
            if (lane == b) my_count += c;
        }
    }
    // start index per code length, longest first (:344-351): start[b] = sum_{b' > b} count[b'] << (max_bits - b')
    uint32_t my_span = (lane >= 1 && lane <= max_bits) ? my_count << (max_bits - lane) : 0;
    uint32_t suffix = my_span;  // inclusive suffix sum over lanes
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_down_sync(0xffffffffu, suffix, d); if (lane + d < 32) suffix += v; }
    uint32_t my_start = suffix - my_span;   // for length == lane
    // per symbol start: start[bits] + (number of earlier symbols with the same length) << (max_bits - bits)
    uint32_t running = 0;   // lane b keeps how many symbols of length b were seen in earlier chunks
    for (uint32_t base = 0; base < nsym; base += 32) {
        uint32_t s = base + lane;
        uint32_t w = s < nsym ? sc.weights[s] : 0;
        uint32_t bits = w ? max_bits + 1 - w : 0;
        uint32_t m = __match_any_sync(0xffffffffu, bits);
        uint32_t before = __shfl_sync(0xffffffffu, running, bits) + __popc(m & lt);
        uint32_t st = __shfl_sync(0xffffffffu, my_start, bits) + (before << (max_bits - bits));
        if (s < nsym) sc.start[s] = bits ? (uint16_t)st : (uint16_t)0xffff;
        for (uint32_t b = 1; b <= max_bits; b++) {
            uint32_t c = __popc(__ballot_sync(0xffffffffu, bits == b));
            if (lane == b) running += c;
        }
    }
    __syncwarp();
    // fill (:360-374): symbol s owns [start, start + 2^(max_bits - bits)); symbols wrap to u8 like `symbol as u8`.
    // Symbols go straight to global memory (byte stores, nothing reads them back here); the 4-bit lengths are
    // merged in shared memory and copied out as 16-byte vectors.
    for (uint32_t s = 0; s < nsym; s++) {
        uint32_t st = sc.start[s];
        if (st == 0xffffu) continue;
        uint32_t bits = max_bits + 1 - sc.weights[s], n = 1u << (max_bits - bits);
        for (uint32_t k = lane; k < n; k += 32) slot->sym[st + k] = (uint8_t)s;
        if (n >= 2) { for (uint32_t k = lane; k < (n >> 1); k += 32) sc.nb4[(st >> 1) + k] = (uint8_t)(bits | (bits << 4)); }
        else if (lane == 0) {
            uint8_t o = sc.nb4[st >> 1];
            sc.nb4[st >> 1] = (st & 1u) ? (uint8_t)((o & 0x0Fu) | (bits << 4)) : (uint8_t)((o & 0xF0u) | bits);
        }
        __syncwarp();
    }
    {
        const uint4 *src4 = reinterpret_cast<const uint4 *>(sc.nb4);
        uint4 *dst4 = reinterpret_cast<uint4 *>(slot->nb4);
        uint32_t n16 = ((1u << max_bits) / 2 + 15) / 16;
        for (uint32_t k = lane; k < n16; k += 32) dst4[k] = src4[k];
    }
    if (lane == 0) { slot->max_bits = max_bits; slot->status = 0; }
    return 0;
}

}  // namespace b200z

// Host-side unit check of the description bit readers in zstd-rs_b200/csrc/tables.cuh (they run on lane 0 of k_setup and on the
// host for formatted dictionaries): the word-based FwdBits::get / RevBitsSmall::get must agree, on random slices and positions,
// with bit-serial restatements of the reference readers -- BitReader::get_bits (bit_io/bit_reader.rs:28-91, LSB first) and
// BitReaderReversed::get_bits (bit_io/bit_reader_reverse.rs:92-113: MSB first from the end, zeros below the start, bits_remaining
// going negative).  Guard bytes behind every slice differ from the data, so an over-read would change a result.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../zstd-rs_b200/csrc/tables.cuh"
using namespace b200z;

static uint32_t serial_rev(const uint8_t *src, int32_t &p, uint32_t n) {
    uint32_t v = 0;
    for (uint32_t k = 0; k < n; k++) {
        const int32_t bit = p - 1 - (int32_t)k;
        v = (v << 1) | (bit >= 0 ? ((src[bit >> 3] >> (bit & 7)) & 1u) : 0u);
    }
    p -= (int32_t)n;
    return v;
}
static bool serial_fwd(const uint8_t *src, uint32_t len, uint32_t &idx, uint32_t n, uint32_t &out) {
    if (len * 8u - idx < n) return false;
    uint32_t v = 0;
    for (uint32_t k = 0; k < n; k++) { const uint32_t i = idx + k; v |= ((src[i >> 3] >> (i & 7u)) & 1u) << k; }
    idx += n; out = v;
    return true;
}
int main() {
    srand(20260923);
    for (int it = 0; it < 100000; it++) {
        const uint32_t len = 1 + rand() % 40;
        std::vector<uint8_t> buf(len + 8, 0xEE);
        for (uint32_t i = 0; i < len; i++) buf[i] = (uint8_t)rand();
        const int32_t p0 = rand() % (len * 8 + 1);
        RevBitsSmall a{buf.data(), p0};
        int32_t pb = p0;
        for (int k = 0; k < 30; k++) {
            const uint32_t n = rand() % 17, x = a.get(n), y = serial_rev(buf.data(), pb, n);
            if (x != y || a.p != pb) { printf("RevBitsSmall mismatch: len %u p0 %d n %u: %x vs %x\n", len, p0, n, x, y); return 1; }
        }
        FwdBits f{buf.data(), len, 0};
        uint32_t ib = 0;
        for (int k = 0; k < 30; k++) {
            uint32_t n = rand() % 25, x = 0, y = 0;
            const bool r1 = f.get(n, x), r2 = serial_fwd(buf.data(), len, ib, n, y);
            if (r1 != r2 || (r1 && x != y) || f.idx != ib) { printf("FwdBits mismatch: len %u n %u: %d %d %x vs %x\n", len, n, r1, r2, x, y); return 1; }
        }
    }
    puts("ok");
    return 0;
}

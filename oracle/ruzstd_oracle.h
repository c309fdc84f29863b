/*
 * ruzstd_oracle.h -- CPU restatement of the ruzstd (KillingSpark/zstd-rs @ eb7e03cc, v0.9.1) decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may build, load or call this.  The product (zstd-rs_b200/, include/b200zstd.h)
 * never links, imports or falls back to it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against every golden vector
 * the reference's own tests hold for the path: 101 decodecorpus frames incl. XXH64
 * checksums and bytes-consumed, 207 dictionary frames, 4 window fixtures, 49 fuzz artifacts (must not crash),
 * the bit-reader / predefined-table / rep-offset / dictionary-header KATs -- and, as a secondary oracle,
 * against the system libzstd 1.5.5.
 *
 * All citations are relative to /root/reference/ruzstd/src/.
 */
#ifndef RUZSTD_ORACLE_H
#define RUZSTD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flat error codes: one per leaf variant of the reference's nested error enums (decoding/errors.rs).
 * The numeric values are shared with include/b200zstd.h (B200Z_ERR_*); tests/test_abi.py asserts that. */
typedef enum zo_error {
    ZO_OK = 0,
    /* ReadFrameHeaderError (errors.rs:95-104) */
    ZO_ERR_MAGIC_NUMBER_READ = 1,
    ZO_ERR_BAD_MAGIC_NUMBER = 2,
    ZO_ERR_FRAME_DESCRIPTOR_READ = 3,
    ZO_ERR_INVALID_FRAME_DESCRIPTOR = 4,
    ZO_ERR_WINDOW_DESCRIPTOR_READ = 5,
    ZO_ERR_DICTIONARY_ID_READ = 6,
    ZO_ERR_FRAME_CONTENT_SIZE_READ = 7,
    ZO_ERR_SKIP_FRAME = 8,
    /* FrameHeaderError (errors.rs:34-42) / FrameDecoderError (errors.rs:472-486) */
    ZO_ERR_WINDOW_TOO_BIG = 10,
    ZO_ERR_WINDOW_TOO_SMALL = 11,
    ZO_ERR_WINDOW_SIZE_TOO_BIG = 12,
    ZO_ERR_DICT_NOT_PROVIDED = 13,
    ZO_ERR_NOT_YET_INITIALIZED = 14,
    ZO_ERR_FAILED_TO_READ_CHECKSUM = 15,
    ZO_ERR_FAILED_TO_DRAIN_DECODEBUFFER = 16,
    ZO_ERR_FAILED_TO_SKIP_FRAME = 17,
    ZO_ERR_TARGET_TOO_SMALL = 18,
    /* BlockHeaderReadError (errors.rs:156-161) */
    ZO_ERR_BLOCK_HEADER_READ = 20,
    ZO_ERR_FOUND_RESERVED_BLOCK = 21,
    ZO_ERR_BLOCK_SIZE_TOO_LARGE = 22,
    /* DecodeBlockContentError / DecompressBlockError (errors.rs:345-350, 256-267) */
    ZO_ERR_DECODER_STATE_IS_FAILED = 30,
    ZO_ERR_EXPECTED_HEADER_OF_PREVIOUS_BLOCK = 31,
    ZO_ERR_BLOCK_BODY_READ = 32,
    ZO_ERR_BLOCK_CONTENT_READ = 33,
    ZO_ERR_MALFORMED_SECTION_HEADER = 34,
    /* LiteralsSectionParseError / SequencesHeaderParseError (errors.rs:816-820, 870-872) */
    ZO_ERR_LITSEC_ILLEGAL_TYPE = 40,
    ZO_ERR_LITSEC_GET_BITS = 41,
    ZO_ERR_LITSEC_NOT_ENOUGH_BYTES = 42,
    ZO_ERR_SEQHDR_NOT_ENOUGH_BYTES = 45,
    /* DecompressLiteralsError (errors.rs:586-598) */
    ZO_ERR_LIT_MISSING_COMPRESSED_SIZE = 50,
    ZO_ERR_LIT_MISSING_NUM_STREAMS = 51,
    ZO_ERR_LIT_GET_BITS = 52,
    ZO_ERR_LIT_UNINITIALIZED_HUFFMAN_TABLE = 55,
    ZO_ERR_LIT_MISSING_BYTES_FOR_JUMP_HEADER = 56,
    ZO_ERR_LIT_MISSING_BYTES_FOR_LITERALS = 57,
    ZO_ERR_LIT_EXTRA_PADDING = 58,
    ZO_ERR_LIT_BITSTREAM_READ_MISMATCH = 59,
    ZO_ERR_LIT_DECODED_LITERAL_COUNT_MISMATCH = 60,
    /* HuffmanTableError (errors.rs:991-1028) */
    ZO_ERR_HUF_GET_BITS = 70,
    ZO_ERR_HUF_FSE_DECODER = 71,
    ZO_ERR_HUF_SOURCE_IS_EMPTY = 72,
    ZO_ERR_HUF_NOT_ENOUGH_BYTES_FOR_WEIGHTS = 73,
    ZO_ERR_HUF_EXTRA_PADDING = 74,
    ZO_ERR_HUF_TOO_MANY_WEIGHTS = 75,
    ZO_ERR_HUF_MISSING_WEIGHTS = 76,
    ZO_ERR_HUF_LEFTOVER_NOT_POWER_OF_2 = 77,
    ZO_ERR_HUF_NOT_ENOUGH_BYTES_TO_DECOMPRESS_WEIGHTS = 78,
    ZO_ERR_HUF_FSE_TABLE_USED_TOO_MANY_BYTES = 79,
    ZO_ERR_HUF_NOT_ENOUGH_BYTES_IN_SOURCE = 80,
    ZO_ERR_HUF_WEIGHT_BIGGER_THAN_MAX_NUM_BITS = 81,
    ZO_ERR_HUF_MAX_BITS_TOO_HIGH = 82,
    /* FSETableError (errors.rs:892-908); the same leaf can surface under Huffman weights, sequence
     * tables or a dictionary -- the code is the leaf, zo_last_error_stage() tells where. */
    ZO_ERR_FSE_ACC_LOG_IS_ZERO = 90,
    ZO_ERR_FSE_ACC_LOG_TOO_BIG = 91,
    ZO_ERR_FSE_GET_BITS = 92,
    ZO_ERR_FSE_PROBABILITY_COUNTER_MISMATCH = 93,
    ZO_ERR_FSE_TOO_MANY_SYMBOLS = 94,
    ZO_ERR_FSE_TABLE_IS_UNINITIALIZED = 95, /* FSEDecoderError::TableIsUninitialized (errors.rs:957-960) */
    /* DecodeSequenceError (errors.rs:726-739) */
    ZO_ERR_SEQ_EXTRA_PADDING = 100,
    ZO_ERR_SEQ_UNSUPPORTED_OFFSET = 101,
    ZO_ERR_SEQ_ZERO_OFFSET = 102,
    ZO_ERR_SEQ_NOT_ENOUGH_BYTES_FOR_NUM_SEQUENCES = 103,
    ZO_ERR_SEQ_EXTRA_BITS = 104,
    ZO_ERR_SEQ_MISSING_COMPRESSION_MODE = 105,
    ZO_ERR_SEQ_MISSING_BYTE_FOR_RLE_LL_TABLE = 106,
    ZO_ERR_SEQ_MISSING_BYTE_FOR_RLE_OF_TABLE = 107,
    ZO_ERR_SEQ_MISSING_BYTE_FOR_RLE_ML_TABLE = 108,
    /* ExecuteSequencesError / DecodeBufferError (errors.rs:683-687, 393-396) */
    ZO_ERR_EXEC_NOT_ENOUGH_BYTES_FOR_SEQUENCE = 110,
    ZO_ERR_EXEC_ZERO_OFFSET = 111,
    ZO_ERR_EXEC_NOT_ENOUGH_BYTES_IN_DICTIONARY = 112,
    ZO_ERR_EXEC_OFFSET_TOO_BIG = 113,
    /* DictionaryDecodeError (errors.rs:419-424) */
    ZO_ERR_DICT_NOT_ENOUGH_BYTES = 120,
    ZO_ERR_DICT_BAD_MAGIC_NUM = 121,
    /* Places where the reference would panic (assert!/unreachable!/index out of bounds) instead of
     * returning an error; SURVEY.md Appendix B.8.  Never produced by spec-valid input. */
    ZO_ERR_REFERENCE_WOULD_PANIC = 200,
    /* Resource limits of this build (documented deviations, never hit by spec-valid input). */
    ZO_ERR_BLOCK_OUTPUT_LIMIT = 210,
    ZO_ERR_INVALID_ARGUMENT = 220,
    ZO_ERR_OUT_OF_MEMORY = 221
} zo_error;

/* Which reference stage raised the last error (the enum nesting path in errors.rs, flattened). */
typedef enum zo_stage {
    ZO_STAGE_NONE = 0,
    ZO_STAGE_FRAME_HEADER = 1,
    ZO_STAGE_BLOCK_HEADER = 2,
    ZO_STAGE_BLOCK_BODY = 3,
    ZO_STAGE_LITERALS = 4,
    ZO_STAGE_SEQUENCES = 5,
    ZO_STAGE_EXECUTE = 6,
    ZO_STAGE_CHECKSUM = 7,
    ZO_STAGE_DICTIONARY = 8,
    ZO_STAGE_DRAIN = 9
} zo_stage;

/* io::Read::read semantics (io_nostd.rs:97-130): return bytes read (0 = EOF), negative = error. */
typedef long (*zo_read_fn)(void *user, uint8_t *buf, size_t len);
/* io::Write::write semantics: return bytes written (0 = cannot accept), negative = error. */
typedef long (*zo_write_fn)(void *user, const uint8_t *buf, size_t len);

/* BlockDecodingStrategy (decoding/frame_decoder.rs:96-100) */
enum { ZO_STRATEGY_ALL = 0, ZO_STRATEGY_UPTO_BLOCKS = 1, ZO_STRATEGY_UPTO_BYTES = 2 };

typedef struct zo_decoder zo_decoder;

/* ---- FrameDecoder mirror (decoding/frame_decoder.rs:154-627) ---- */
zo_decoder *zo_new(void);
void zo_free(zo_decoder *d);
void zo_set_max_window_size(zo_decoder *d, uint64_t max_window_size);
uint64_t zo_max_window_size(const zo_decoder *d);
int zo_init(zo_decoder *d, zo_read_fn read, void *user); /* == reset */
int zo_add_dict(zo_decoder *d, const uint8_t *raw, size_t len); /* Dictionary::decode_dict + add_dict */
/* Extension (absent from the reference, SURVEY.md 8c gap): raw-content dictionary = content only,
 * offset history [1,4,8], no entropy tables; selected with zo_force_dict(id). */
int zo_add_raw_content_dict(zo_decoder *d, uint32_t id, const uint8_t *content, size_t len);
int zo_force_dict(zo_decoder *d, uint32_t dict_id);
int zo_decode_blocks(zo_decoder *d, zo_read_fn read, void *user, int strategy, size_t n, int *finished);
long zo_read(zo_decoder *d, uint8_t *target, size_t len);           /* impl Read for FrameDecoder */
long zo_collect_to_writer(zo_decoder *d, zo_write_fn write, void *user);
size_t zo_can_collect(const zo_decoder *d);
int zo_is_finished(const zo_decoder *d);
size_t zo_blocks_decoded(const zo_decoder *d);
uint64_t zo_bytes_read_from_source(const zo_decoder *d);
uint64_t zo_content_size(const zo_decoder *d);
int zo_get_checksum_from_data(const zo_decoder *d, uint32_t *out); /* 1 = Some, 0 = None */
int zo_get_calculated_checksum(const zo_decoder *d, uint32_t *out);
int zo_decode_from_to(zo_decoder *d, const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_len,
                      size_t *read, size_t *written);
int zo_decode_all(zo_decoder *d, const uint8_t *input, size_t in_len, uint8_t *output, size_t out_cap,
                  size_t *written);
int zo_last_error_stage(const zo_decoder *d);
/* frame header facts of the current frame (frame.rs:88-150) */
uint64_t zo_window_size(const zo_decoder *d);
int zo_frame_dict_id(const zo_decoder *d, uint32_t *out);

/* ---- per-block trace: intermediate results the CUDA kernels are checked against ---- */
typedef struct zo_block_trace {
    uint32_t block_type;      /* 0 raw, 1 rle, 2 compressed */
    uint32_t literals_type;   /* 0 raw, 1 rle, 2 compressed, 3 treeless (compressed blocks only) */
    uint32_t num_streams;     /* 0, 1 or 4 */
    uint32_t regenerated_size;
    uint32_t num_sequences;
    uint32_t huf_max_bits;
    uint64_t lit_offset;      /* into trace literals pool */
    uint64_t seq_offset;      /* into trace sequence pool, in sequences */
    uint64_t out_offset;      /* frame output position where this block starts */
    uint64_t out_size;
    uint32_t offset_hist_after[3];
    uint32_t pad;
} zo_block_trace;

typedef struct zo_seq_trace {
    uint32_t ll, ml, of;      /* Sequence (blocks/sequence_section.rs:21-37) */
    uint32_t actual_offset;   /* after do_offset_history (sequence_execution.rs:59-118) */
} zo_seq_trace;

void zo_trace_enable(zo_decoder *d, int on);
size_t zo_trace_num_blocks(const zo_decoder *d);
const zo_block_trace *zo_trace_blocks(const zo_decoder *d);
const uint8_t *zo_trace_literals(const zo_decoder *d, size_t *len);
const zo_seq_trace *zo_trace_sequences(const zo_decoder *d, size_t *count);

/* ---- primitives exposed for KATs (tests/test_oracle_kat.py) ---- */
/* BitReaderReversed (bit_io/bit_reader_reverse.rs): read bits `counts[i]` in turn, return values and the
 * final bits_remaining(). */
long zo_kat_bitreader_reversed(const uint8_t *src, size_t len, const uint8_t *counts, size_t n, uint64_t *values);
/* BitReader forward (bit_io/bit_reader.rs) */
int zo_kat_bitreader_forward(const uint8_t *src, size_t len, const uint8_t *counts, size_t n, uint64_t *values);
/* FSETable::build_from_probabilities (fse/fse_decoder.rs:126-220): fills entries[1<<acc_log] as
 * {base_line, num_bits, symbol} triples of u32. */
int zo_kat_fse_build(const int32_t *probs, size_t nprobs, uint8_t acc_log, uint8_t max_symbol, uint32_t *entries3);
/* FSETable::build_decoder from a serialized description; returns bytes used or -error */
long zo_kat_fse_read(const uint8_t *src, size_t len, uint8_t max_log, uint8_t max_symbol, uint8_t *acc_log,
                     uint32_t *entries3, size_t cap_entries);
/* HuffmanTable::build_decoder (huff0/huff0_decoder.rs:117-123): returns bytes used or -error;
 * entries2[i] = symbol | num_bits<<8 */
long zo_kat_huf_build(const uint8_t *src, size_t len, uint8_t *max_bits, uint16_t *entries, size_t cap_entries);
/* do_offset_history (decoding/sequence_execution.rs:59-118) */
uint32_t zo_kat_do_offset_history(uint32_t offset_value, uint32_t lit_len, uint32_t hist[3]);
/* Dictionary::decode_dict (decoding/dictionary.rs:45-126): returns 0 or error; outputs id, offsets, content len */
int zo_kat_decode_dict(const uint8_t *raw, size_t len, uint32_t *id, uint32_t offs[3], size_t *content_len);
/* XXH64 seed 0 (twox-hash call sites decode_buffer.rs:42,225,290,301) */
uint64_t zo_xxh64(const uint8_t *data, size_t len);

/* ---- bulk helper for the CPU baseline: decode `nframes` independent frames (offsets into one buffer) with
 * one decoder per thread over `nthreads` pthreads; returns 0 or the first error.  out_offsets[i] gives where
 * frame i's plaintext goes; out_caps[i] its capacity. */
int zo_bulk_decode(const uint8_t *input, const uint64_t *in_offsets, const uint64_t *in_sizes, size_t nframes,
                   uint8_t *output, const uint64_t *out_offsets, const uint64_t *out_caps, uint64_t *out_sizes,
                   const uint8_t *raw_dict, size_t raw_dict_len, int nthreads);

#ifdef __cplusplus
}
#endif
#endif

/*
 * oracle/lz4_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference LZ4 block codec (lz4/lz4 v1.10.0, lib/lz4.c) used as the
 * parity checker for the CUDA path.  Nothing in the product (lz4_b200/, include/) may include,
 * link or call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do.
 *
 * Parity pin: oracle/_ref/libref_lz4.so (the unmodified reference lib/lz4.c compiled in place
 * by oracle/Makefile) -- tests/test_oracle_vs_ref.py checks byte-identical compressed output and
 * identical decoder return values / bytes against it, and tests/golden/ holds vectors generated
 * from it (tests/golden/make_golden.py).
 */
#ifndef LZ4_ORACLE_H
#define LZ4_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lz4.h:214-215, lz4.c:751 */
int oracle_lz4_compress_bound(int srcSize);

/* lz4.c:1453 (LZ4_compress_fast) -> :1382 (extState) -> :1344 (generic) -> :930 (generic_validated),
 * noDict / noDictIssue instantiation, 64-bit little-endian host. Byte-identical output. */
int oracle_lz4_compress_fast(const uint8_t* src, uint8_t* dst, int srcSize, int dstCapacity, int acceleration);

/* lz4.c:2451 (LZ4_decompress_safe) -> :2022 (LZ4_decompress_generic, decode_full_block, noDict),
 * x86-64 build (LZ4_FAST_DEC_LOOP=1): same accepted set, same return value (size or -(ip)-1),
 * same bytes in dst[0, ret). */
int oracle_lz4_decompress_safe(const uint8_t* src, uint8_t* dst, int compressedSize, int dstCapacity);

/* tests/datagen.c:154-160 (RDG_genBuffer): synthetic compressible data, byte-identical. */
void oracle_datagen(uint8_t* buffer, size_t size, double matchProba, double litProba, unsigned seed);

/* Walk a valid block and report sequence statistics (used by tests / DESIGN numbers). */
typedef struct {
    uint32_t n_sequences;      /* tokens, including the final literal-only one */
    uint32_t literal_bytes;
    uint32_t match_bytes;
    uint32_t overlap_matches;  /* offset < match length */
} oracle_block_stats;
int oracle_lz4_block_stats(const uint8_t* src, int compressedSize, oracle_block_stats* st);

/* ---- CPU timing harness (bench.py cpu_baseline / --impl reference) ------------------------- */
typedef int (*oracle_decomp_fn)(const char* src, char* dst, int compressedSize, int dstCapacity);
typedef int (*oracle_comp_fn)(const char* src, char* dst, int srcSize, int dstCapacity, int acceleration);

/* Decompress nBlocks blocks (src + srcOff[i], size srcSize[i]) into dst + i*dstStride with
 * `threads` pthreads (static partition, as BASELINE.md section 3); returns wall seconds of the
 * pass, or <0 if any block failed.  outSizes[i] receives each return value. */
double oracle_time_decompress(oracle_decomp_fn fn, const uint8_t* src, const int64_t* srcOff,
                              const int32_t* srcSize, uint8_t* dst, int64_t dstStride, int32_t dstCap,
                              int32_t* outSizes, int64_t nBlocks, int threads);
double oracle_time_compress(oracle_comp_fn fn, const uint8_t* src, int64_t srcStride, int32_t srcSize,
                            int64_t lastSize, uint8_t* dst, int64_t dstStride, int32_t dstCap, int accel,
                            int32_t* outSizes, int64_t nBlocks, int threads);
/* multi-threaded datagen: segment k of segBytes gets seed seed0+k */
void oracle_first_touch(uint8_t* dst, int64_t dstStride, int64_t nBlocks, int threads);
void oracle_pack(const uint8_t* slots, int64_t slotStride, const int32_t* sizes, const int64_t* packOff,
                 uint8_t* packed, int64_t nBlocks, int threads);
void oracle_datagen_mt(uint8_t* buffer, size_t size, size_t segBytes, double matchProba, unsigned seed0, int threads);

#ifdef __cplusplus
}
#endif
#endif

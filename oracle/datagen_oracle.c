/*
 * oracle/datagen_oracle.c -- TEST INFRASTRUCTURE ONLY (see lz4_oracle.h).
 *
 * Restatement of the reference's synthetic-data generator (tests/datagen.c:60-160,
 * RDG_genBuffer), the input named by BASELINE.json ("tests/datagen P50 / P90").  Byte-identical
 * to the compiled reference generator (checked in tests/test_oracle_vs_ref.py and pinned by
 * SHA-256 digests in tests/golden/).
 */
#include "lz4_oracle.h"
#include <string.h>

#define LIT_TABLE_LOG 13
#define LIT_TABLE_SIZE (1u << LIT_TABLE_LOG)

/* datagen.c:60-68: multiplicative/xor/rotate generator */
static uint32_t next_rand(uint32_t* state)
{
    uint32_t r = *state;
    r *= 2654435761u;
    r ^= 2246822519u;
    r = (r << 13) | (r >> 19);
    *state = r;
    return r;
}

/* datagen.c:71-88: skewed literal alphabet '('..'}' starting at '0' (or 0..255 when ld <= 0) */
static void fill_literal_table(uint8_t* lt, double ld)
{
    const uint8_t first = ld <= 0.0 ? 0 : '(';
    const uint8_t last = ld <= 0.0 ? 255 : '}';
    uint8_t ch = ld <= 0.0 ? 0 : '0';
    uint32_t u = 0;
    while (u < LIT_TABLE_SIZE) {
        uint32_t weight = (uint32_t)((double)(LIT_TABLE_SIZE - u) * ld) + 1;
        uint32_t end = u + weight;
        if (end > LIT_TABLE_SIZE) end = LIT_TABLE_SIZE;
        while (u < end) lt[u++] = ch;
        if (ch == last) ch = first; else ch++;
    }
}

static uint8_t gen_literal(uint32_t* seed, const uint8_t* lt) { return lt[next_rand(seed) & (LIT_TABLE_SIZE - 1)]; }
static uint32_t rand15(uint32_t* seed) { return (next_rand(seed) >> 3) & 32767; }
/* datagen.c:99-100: 7/8 short (0..15), 1/8 long (15..526) */
static uint32_t rand_length(uint32_t* seed)
{
    if ((next_rand(seed) >> 7) & 7) return next_rand(seed) & 15;
    return (next_rand(seed) & 511) + 15;
}

/* datagen.c:101-151 with prefixSize == 0 */
static void gen_block(uint8_t* buf, size_t size, double matchProba, const uint8_t* lt, uint32_t* seed)
{
    const uint32_t matchProba32 = (uint32_t)(32768 * matchProba);
    size_t pos = 0;

    while (matchProba >= 1.0) {                       /* datagen.c:109-120: zero runs */
        size_t size0 = next_rand(seed) & 3;
        size0 = (size_t)1 << (16 + size0 * 2);
        size0 += next_rand(seed) & (size0 - 1);
        if (size < pos + size0) { memset(buf + pos, 0, size - pos); return; }
        memset(buf + pos, 0, size0);
        pos += size0;
        buf[pos - 1] = gen_literal(seed, lt);
    }

    if (size == 0) return;
    buf[0] = gen_literal(seed, lt);
    pos = 1;

    while (pos < size) {
        if (rand15(seed) < matchProba32) {            /* datagen.c:129-141: copy within 32K */
            size_t length = (size_t)rand_length(seed) + 4;
            uint32_t offset = rand15(seed) + 1;
            size_t from, end;
            if (offset > pos) offset = (uint32_t)pos;
            from = pos - offset;
            end = pos + length;
            if (end > size) end = size;
            while (pos < end) buf[pos++] = buf[from++];
        } else {                                      /* datagen.c:142-149: literal noise */
            size_t length = rand_length(seed);
            size_t end = pos + length;
            if (end > size) end = size;
            while (pos < end) buf[pos++] = gen_literal(seed, lt);
        }
    }
}

void oracle_datagen(uint8_t* buffer, size_t size, double matchProba, double litProba, unsigned seed)
{
    uint8_t lt[LIT_TABLE_SIZE];
    uint32_t s = seed;
    if (litProba == 0.0) litProba = matchProba / 4.5;  /* datagen.c:157 */
    fill_literal_table(lt, litProba);
    gen_block(buffer, size, matchProba, lt, &s);
}

/*
 * lz4_phaseb_v2.h -- EXPERIMENTAL variant of phase B of lz4_expand_fast_kernel ("uniform body").
 *
 * Not part of the default build: lz4_kernels.cu includes it only with -DLZ4K_PHASEB_V2
 * (LZ4K_PHASEB_V2=1 python -m lz4_b200.build --force).  Written in round 1 after the GPU budget was
 * spent; its ARITHMETIC is checked on the CPU by tests/test_phaseb_v2_emul.py, which compiles this
 * header with g++ and replays the loop lane by lane on reference-compressed blocks, but it has not
 * run on a GPU yet -- measure it before making it the default.
 *
 * What changes against the shipped loop (same work distribution, same done-flag protocol):
 *   - literal and match pieces share ONE code path: the source is an index into a single shared-memory
 *     window (the staged input and the output window are members of the same struct), the done-flag
 *     indices of a literal piece point at a sentinel flag that is always set;
 *   - the sequence record is re-read from shared memory at the top of every iteration instead of in
 *     two divergent places (chunk hand-out and "next sequence starts inside this chunk");
 *   - only short offsets (< 8, incl. the invalid-but-accepted 0) keep a separate, rarely taken path.
 * Instruction count per loop iteration when every path is live (cuobjdump -sass of both builds):
 * shipped 10 header + 45 hand-out + 18 literal + 28 match + 28 tail + 12 store + 18 next-sequence
 * = ~161; this variant 18 header/ballots + 40 hand-out + 17 record + 33 uniform source/flags/load +
 * 15 tail + 13 store (or 6 sequence step) = ~144, i.e. about 10 % fewer warp-instructions for the
 * same number of iterations -- a modest gain, to be confirmed on the GPU.
 *
 * Everything here is plain C++ on plain pointers so that the same text compiles for the device and
 * for the CPU emulator.
 */
#ifndef LZ4_PHASEB_V2_H
#define LZ4_PHASEB_V2_H

#include <stdint.h>

#if defined(__CUDACC__)
#define PB_FN __device__ __forceinline__
#define PB_POPC(x) __popc(x)
#define PB_FUNNEL_R(lo, hi, s) __funnelshift_r((lo), (hi), (s))
#define PB_FENCE() asm volatile("fence.acq_rel.cta;" ::: "memory")
#else
#define PB_FN static inline
#define PB_POPC(x) __builtin_popcount(x)
static inline uint32_t pb_funnel_r_host(uint32_t lo, uint32_t hi, uint32_t s)
{
    s &= 31u;
    return s ? (lo >> s) | (hi << (32u - s)) : lo;
}
#define PB_FUNNEL_R(lo, hi, s) pb_funnel_r_host((lo), (hi), (s))
#define PB_FENCE() ((void)0)
#endif

struct pb_rec { uint32_t x, y; };       /* {matchStart | nextStart<<16, (litSrc-outStart)&0xFFFF | offset<<16} */

enum { kPbSentinel = 8192 };            /* done8[kPbSentinel] is always 1 */

struct PBView {
    const uint8_t* window;              /* 4-byte aligned; staged input at window[0..], output at window[outDelta..] */
    uint8_t* out;                       /* == window + outDelta */
    int outDelta;                       /* multiple of 4 */
    const pb_rec* rec;
    const uint32_t* bits;               /* bit p: a sequence starts at output byte p */
    const uint16_t* seqbase;            /* number of start bits before bits[i] */
    volatile uint8_t* done8;            /* done8[c] != 0: output bytes [8c, 8c+8) are final */
    int head;                           /* staged input: byte i of the block is window[head + i] */
    int total;                          /* decoded size of the block */
};

struct PBLane {
    int p, pe, k, pos;
    uint64_t acc;
    bool needNew, exhausted;
};

/* unaligned 64-bit read at byte index idx (>= -4) of a 4-aligned array */
PB_FN uint64_t pb_lds64u(const uint8_t* base, int idx)
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base + (idx & ~3));
    const uint32_t sh = (uint32_t)(idx & 3) * 8u;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    return (uint64_t)PB_FUNNEL_R(w0, w1, sh) | ((uint64_t)PB_FUNNEL_R(w1, w2, sh) << 32);
}

PB_FN void pb_init(PBLane& L)
{
    L.p = L.pe = L.k = L.pos = 0;
    L.acc = 0;
    L.needNew = true;
    L.exhausted = false;
}

/* chunk c of warp `warp`'s list (strips warp, warp+32, ... of 256 bytes) */
PB_FN void pb_take(PBLane& L, const PBView& V, int warp, int c, int warpChunks)
{
    if (c >= warpChunks) { L.exhausted = true; return; }
    const int p = ((warp + ((c >> 5) << 5)) << 8) + ((c & 31) << 3);
    if (p >= V.total) return;                               /* chunks past the end of the block are skipped */
    const uint32_t bw = V.bits[p >> 5];
    L.p = p;
    L.pe = (p + 8 < V.total) ? p + 8 : V.total;
    L.k = (int)V.seqbase[p >> 5] + PB_POPC(bw & (0xFFFFFFFFu >> (31 - (p & 31)))) - 1;
    L.pos = p;
    L.acc = 0;
    L.needNew = false;
}

/* one piece of the lane's chunk; returns false when the piece is blocked on an unfinished source */
PB_FN bool pb_body(PBLane& L, const PBView& V)
{
    const pb_rec r = V.rec[L.k];
    int m = (int)(r.x & 0xFFFFu), e = (int)(r.x >> 16);
    const int off = (int)(r.y >> 16);
    const uint32_t d = r.y & 0xFFFFu;
    if (m == 0 && L.k != 0) m = 65536;                      /* 16-bit wrap of 65536 */
    if (e == 0) e = 65536;
    const bool isLit = L.pos < m;
    const int lim = isLit ? m : e;
    const int end = lim < L.pe ? lim : L.pe;
    bool ok = true;
    uint64_t v = 0;
    if (!isLit && off < 8) {                                /* short period, or offset 0 (zero bytes, lz4.c:2407) */
        if (off != 0) {
            for (int x = L.pos; x < end; x++) {
                int sidx = x - off;
                if (sidx >= m) sidx = m - off + ((x - m) % off);        /* always before the match */
                uint32_t byte;
                if (sidx >= L.p) {
                    byte = (uint32_t)((L.acc >> (8 * (sidx - L.p))) & 0xFFu);
                } else {
                    if (!V.done8[sidx >> 3]) { ok = false; break; }
                    PB_FENCE();
                    byte = V.out[sidx];
                }
                v |= (uint64_t)byte << (8 * (x - L.pos));
            }
        }
    } else {
        const int src = L.pos - off;
        const int f0 = isLit ? (int)kPbSentinel : (src >> 3);
        const int f1 = isLit ? (int)kPbSentinel : ((end - 1 - off) >> 3);
        const int idx = isLit ? V.head + (int)(((uint32_t)L.pos + d) & 0xFFFFu) : V.outDelta + src;
        ok = (V.done8[f0] & V.done8[f1]) != 0;
        if (ok) {
            if (!isLit) PB_FENCE();                         /* flags before data */
            v = pb_lds64u(V.window, idx);
        }
    }
    if (!ok) return false;
    const int len = end - L.pos;
    v &= 0xFFFFFFFFFFFFFFFFull >> (64 - 8 * len);
    L.acc |= v << (8 * (L.pos - L.p));
    L.pos = end;
    if (L.pos >= L.pe) {                                    /* chunk complete: publish it */
        *reinterpret_cast<uint64_t*>(V.out + L.p) = L.acc;
        PB_FENCE();                                         /* data before flag */
        V.done8[L.p >> 3] = 1;
        L.needNew = true;
    } else {
        L.k += (L.pos == e) ? 1 : 0;                        /* next sequence starts inside this chunk */
    }
    return true;
}

#endif /* LZ4_PHASEB_V2_H */

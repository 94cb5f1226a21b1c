/*
 * lz4_scan_v2.h -- EXPERIMENTAL intra-block parallel scan: one WARP per block instead of one thread.
 *
 * Not part of the default build (lz4_kernels.cu uses it only with -DLZ4K_SCAN_V2).  Written in
 * round 1 after the GPU budget was spent: its logic is checked on the CPU by
 * tests/test_scan_v2_emul.py (this header compiled by g++, the 32 lanes run phase by phase) against
 * the one-thread scan of lz4_scan_core.h -- same return value, same sequence count, same marks -- but
 * it has not run on a GPU yet.
 *
 * Why: the one-thread scan walks ~2 400 dependent sequences per 64 KB block, so its duration does not
 * shrink with the batch (2.5 ms for 4 096 blocks as for 65 536) and a 4 MB block costs a single thread
 * ~150 000 steps.  A token chain started at an ARBITRARY byte of LZ4 data falls back onto the true
 * chain quickly (tests/perf/sync_study.py: 74 bytes median, 98 % within 1 KB on P50 data), which
 * allows a speculative split:
 *
 *   1. the input range of the front loop (token positions <= n-26) is cut in 32 segments; lane l
 *      parses from the first byte of segment l as if it were a token, up to the first token at or
 *      past the next segment (COUNT pass: exit position, #sequences, #output bytes; only the checks
 *      that depend on input positions);
 *   2. fix-up: lane l's true entry is lane l-1's exit.  Lanes whose entry differs from what they
 *      parsed from parse again; repeated until nothing changes (lane 0 starts at 0, so by induction
 *      every lane ends up parsing from its true entry; usually one extra round);
 *   3. exclusive sums give every lane its first sequence index and output position;
 *   4. WRITE pass: every lane re-walks its segment with absolute positions, writes the marks and
 *      applies the output-dependent rules of scan_front (capacity, offset before the start);
 *   5. the first lane (in order) whose walk ends -- leaving the front region, a rule of scan_front that
 *      hands over to the byte-wise code, or the offset error -- owns the rest: it returns the error
 *      or runs scan_tail from exactly the state the one-thread front loop would have reached.
 *
 * The result is bit-identical to scan_block(): same return value, nSeq and marks[0, nSeq).
 * The lanes communicate through a small per-warp record in shared memory, phase by phase
 * (__syncwarp between phases on the device; the CPU emulator simply runs the 32 lanes of a phase one
 * after the other), so the device kernel and the emulator execute the same text.
 */
#ifndef LZ4_SCAN_V2_H
#define LZ4_SCAN_V2_H

#include "lz4_scan_core.h"

enum { SV2_RAN = 0, SV2_END = 1, SV2_ERR = 2 };
constexpr int kSv2MinBytes = 2048;         /* smaller inputs: lane 0 runs the one-thread scan */

struct SV2Res {                            /* COUNT pass of one lane */
    int exitPos;                           /* first token position at or past the segment end (or where the walk stopped) */
    int stop;                              /* 1: the front region ends at exitPos (later lanes have nothing) */
    uint32_t count, olen;                  /* sequences committed, output bytes they produce */
};
struct SV2End {                            /* WRITE pass of one lane */
    int kind;                              /* SV2_RAN: ran into the next segment; SV2_END: front loop ends here; SV2_ERR */
    int ip;                                /* SV2_END: token position to resume at; SV2_ERR: error position */
    uint32_t op, nseq;                     /* SV2_END: output position / sequence index at ip */
    int nextEvt;
};
struct SV2Shared {
    SV2Res res[32];
    SV2End end[32];
    int changed;
    int ret;
    uint32_t nseq;
};
struct SV2Lane {                           /* registers of one lane */
    int segStart, segEnd, from, isVoid;
    int newFrom, newVoid, need;
    uint32_t seqBase, outBase;
};

/* One walk over [from, segEnd): scan_front's loop body with the output position relative (COUNT) or
 * absolute (WRITE).  Position-only rules apply in both passes, output-dependent rules in WRITE only. */
template <bool WRITE>
SC_FN void sv2_walk(const uint8_t* __restrict__ src, int nI, int capI, int from, int segEnd,
                    uint32_t opBase, uint32_t seqBase, uint32_t* marks, SV2Res& R, SV2End& E)
{
    int fip = from, nextEvt = 0, stop = 0, kind = SV2_RAN, errIp = 0;
    uint32_t fop = opBase, cnt = 0;
    while (fip < segEnd) {
        if (fip > nI - 26) { stop = 1; kind = SV2_END; break; }
        if (fip >= nextEvt) {                                      // L1 prefetch, once per 128 input bytes
            if (fip + 128 < nI) prefetch_l1(src + fip + 128);
            nextEvt = ((fip >> 7) + 1) << 7;
        }
        if (WRITE) { const uint32_t nseq = seqBase + cnt; MARK_VISIT(fip, fop); }
        const uint32_t v = ld32u(src + fip);
        const int mcode = (int)(v & 15u);
        int lit = (int)((v >> 4) & 15u), q = 1;
        if (lit == 15) {
            uint32_t b = (v >> 8) & 0xFFu;
            lit += (int)b; q = 2;
            while (b == 255u && fip + q <= nI - 15 && lit < (1 << 28)) { b = ldb(src + fip + q); q++; lit += (int)b; }
            if (b == 255u || fip + q > nI - 15) { stop = 1; kind = SV2_END; break; }
            if ((uint32_t)(fip + q) + (uint32_t)lit + 32u > (uint32_t)nI) { stop = 1; kind = SV2_END; break; }
            if (WRITE && fop + (uint32_t)lit > (uint32_t)(capI - 32)) { kind = SV2_END; break; }
        }
        const int offPos = fip + q + lit;
        const uint32_t v3 = ld32u(src + offPos);
        const uint32_t off16 = v3 & 0xFFFFu;
        int mlen = mcode + kMinMatch, ipn = offPos + 2;
        if (mcode == 15) {
            uint32_t b = (v3 >> 16) & 0xFFu;
            ipn++; mlen += (int)b;
            while (b == 255u && ipn <= nI - 4 && mlen < (1 << 28)) { b = ldb(src + ipn); ipn++; mlen += (int)b; }
            if (b == 255u || ipn > nI - 4) { stop = 1; kind = SV2_END; break; }
        }
        const uint32_t opn = fop + (uint32_t)lit;
        if (WRITE) {
            if (opn + (uint32_t)mlen >= (uint32_t)(capI - 64)) { kind = SV2_END; break; }
            if (off16 > opn) { kind = SV2_ERR; errIp = ipn; break; }
        }
        fip = ipn; fop = opn + (uint32_t)mlen; cnt++;
    }
    if (WRITE) {
        E.kind = kind; E.ip = (kind == SV2_ERR) ? errIp : fip; E.op = fop; E.nseq = seqBase + cnt; E.nextEvt = nextEvt;
    } else {
        R.exitPos = fip; R.stop = stop; R.count = cnt; R.olen = fop - opBase;
    }
}

/* ---- phase 0: segments + first speculative walk ---- */
SC_FN void sv2_phase0(int lane, SV2Lane& L, SV2Shared& S, const uint8_t* src, int nI, int capI)
{
    const int lim = nI - 26;                                   /* last token position of the front region */
    int seg = (lim + 32) / 32;
    if (seg < 64) seg = 64;
    L.segStart = lane * seg;
    L.segEnd = (lane == 31) ? 0x7FFFFFFF : (lane + 1) * seg;
    L.from = L.segStart;
    L.isVoid = 0;
    SV2End unused;
    sv2_walk<false>(src, nI, capI, L.from, L.segEnd, 0u, 0u, nullptr, S.res[lane], unused);
    if (lane == 0) S.changed = 0;
}

/* ---- fix-up round, part 1 (read): where does my segment really start? ---- */
SC_FN void sv2_decide(int lane, SV2Lane& L, const SV2Shared& S)
{
    L.need = 0;
    if (lane == 0) { L.newFrom = 0; L.newVoid = 0; return; }     /* (the caller resets S.changed between rounds) */
    const SV2Res prev = S.res[lane - 1];
    L.newVoid = prev.stop;
    L.newFrom = prev.exitPos;
    L.need = (L.newVoid != L.isVoid) || (L.newFrom != L.from);
}

/* ---- fix-up round, part 2 (write): walk again from the new entry ---- */
SC_FN void sv2_redo(int lane, SV2Lane& L, SV2Shared& S, const uint8_t* src, int nI, int capI)
{
    if (!L.need) return;
    L.from = L.newFrom;
    L.isVoid = L.newVoid;
    if (L.isVoid) {                                            /* the front region ended in an earlier lane */
        S.res[lane].exitPos = L.from; S.res[lane].stop = 1; S.res[lane].count = 0; S.res[lane].olen = 0;
    } else {
        SV2End unused;
        sv2_walk<false>(src, nI, capI, L.from, L.segEnd, 0u, 0u, nullptr, S.res[lane], unused);
    }
    S.changed = 1;
}

/* ---- bases + WRITE pass ---- */
SC_FN void sv2_write(int lane, SV2Lane& L, SV2Shared& S, const uint8_t* src, int nI, int capI, uint32_t* marks)
{
    uint32_t sb = 0, ob = 0;
    for (int j = 0; j < lane; j++) { sb += S.res[j].count; ob += S.res[j].olen; }
    L.seqBase = sb; L.outBase = ob;
    if (L.isVoid) { S.end[lane].kind = SV2_RAN; return; }
    SV2Res unused;
    sv2_walk<true>(src, nI, capI, L.from, L.segEnd, ob, sb, marks, unused, S.end[lane]);
}

/* ---- the first lane whose walk ended finishes the block ---- */
SC_FN void sv2_finish(int lane, SV2Shared& S, const uint8_t* src, int nI, int capI, uint32_t* marks)
{
    int j = 0;
    while (j < 31 && S.end[j].kind == SV2_RAN) j++;
    if (lane != j) return;
    const SV2End e = S.end[j];
    if (e.kind == SV2_ERR) { S.ret = (int)(-(int64_t)e.ip) - 1; S.nseq = 0; return; }   /* lz4.c:2443 */
    ScanState st;
    st.ip = e.ip; st.op = (int64_t)e.op; st.nseq = e.nseq; st.fast = true;
    st.nextPrefetch = (e.nextEvt > 0) ? (int64_t)e.nextEvt + 128 : 128;
    uint32_t ns = 0;
    S.ret = scan_tail(src, nI, capI, st, &ns, marks);
    S.nseq = ns;
}

#endif /* LZ4_SCAN_V2_H */

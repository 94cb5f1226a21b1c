/*
 * lz4_encode_par.cuh -- the PARALLEL-PARSE compressor: one CTA per block, every phase data-parallel.
 *
 * Included by lz4_kernels.cu.  Replaces, for throughput, the role of LZ4_compress_generic_validated
 * (lz4.c:930-1338) -- 4-byte hash table match finder + greedy parse -- with a formulation that has no
 * sequential walk over the block.  The output is a valid LZ4 block that round-trips exactly, it is
 * deterministic (no result depends on scheduling), but it is NOT byte-identical to the reference's output:
 * the parse differs (every position is a candidate, no skipping), the ratio stays within 2 % of the
 * reference's at acceleration 1 (tests/test_gpu_parity.py).  The byte-identical encoder
 * (lz4_encode_kernel) remains behind LZ4_compress_default / LZ4_compress_fast and LZ4B200_compress_blocks.
 *
 * A block of n <= 65 536 bytes is staged in shared memory by one TMA bulk load, then processed in WINDOWS
 * of 4096 positions; thread t of 1024 owns the 4 consecutive positions 4t .. 4t+3 of the window:
 *
 *   find    every position hashes its 4 bytes (Fibonacci hash, 13 bits: lz4.c:779) and reads T[h] = the
 *           LATEST position of an EARLIER window with that hash; atomicMax publishes, in T2[h], the EARLIEST
 *           position of THIS window with that hash (both orders are scheduling-independent).  A position's
 *           candidate is T2's (if it lies before the position and its 4 bytes match), else T's.
 *           After the look-ups of a window every position enters T (atomicMax).
 *   lengths consecutive positions with the same offset form a RUN: only its first position extends its
 *           match (36 bytes alone; longer ones are queued and finished by a whole warp, 128 bytes per
 *           iteration), the others derive their length from it.
 *   select  the reference's greedy rule -- the first position at or after the end of the previous match
 *           that has a match is taken (lz4.c:1014-1100) -- evaluated by one warp: each lane walks 128
 *           positions speculatively, then lanes whose entry point moved walk again until nothing changes.
 *   emit    <= 1 sequence per thread (matches are >= 4 long): backward extension (lz4.c:1107-1109), sizes,
 *           CTA-wide exclusive sum, token / lengths / literals / offset written straight to the output;
 *           literal runs above 32 bytes are copied by a whole warp.
 *
 * The end-of-block rules of the format are the reference's: no match starts after n-12, the last 5 bytes
 * are literals (lz4.c:963-964, 1233); output that does not fit dstCapacity makes the call return 0.
 * `acceleration` > 1 thins the candidate positions (every `step`-th position is hashed / inserted).
 */
#pragma once

constexpr int kEpThreads = 1024;
constexpr int kEpWin = 4096;                         /* positions per window = 4 per thread */
constexpr int kEpHashLog = 13;
constexpr int kEpSoloLen = 20;                       /* bytes a run start compares alone before queueing the match for a warp */
constexpr int kEpMaxJobs = 128;                      /* long matches a window finishes with a warp each: the first ones by position */
constexpr int kEpInlineLits = 32;                    /* literal runs up to this length are copied by the emitting thread */
constexpr int kEpMaxLitJobs = 1024;
constexpr int kEpStage = 24576;                      /* a window's output is assembled here and written out coalesced when it fits */

struct EncParSmem {
    alignas(16) uint8_t src[65536 + 64];             /* staged block (keeps the source's 16-byte phase) */
    uint32_t T[1 << kEpHashLog];                     /* latest position + 1 of an earlier window, per hash */
    uint32_t T2[1 << kEpHashLog];                    /* (window + 1) << 16 | (0xFFFF - index in window): earliest of this window */
    uint16_t cand[kEpWin];                           /* candidate position, 0xFFFF = none */
    uint16_t len[kEpWin];                            /* run starts: match length */
    uint16_t start[kEpWin];                          /* window index of the run start at or before this position */
    uint16_t litStart[kEpWin];                       /* selected positions: where their literals start */
    uint32_t hasBits[kEpWin / 32], selBits[kEpWin / 32];
    uint16_t jobIdx[kEpMaxJobs];
    uint32_t litJob[kEpMaxLitJobs][3];               /* {source position, output offset, length} */
    alignas(16) uint8_t stage[kEpStage];
    uint32_t warpA[32], warpB[32];
    uint32_t nJobs, nLitJobs, E, O, fail;
    alignas(8) uint64_t mbar;
};
static_assert(sizeof(EncParSmem) <= 232448, "EncParSmem exceeds the shared memory a CTA can opt in to");

__device__ __forceinline__ uint32_t ep_ld32(const uint8_t* base, uint32_t i)      /* unaligned 4 bytes at base + i (base 4-aligned) */
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base) + (i >> 2);
    const uint32_t sh = (i & 3u) * 8u;
    return __funnelshift_r(w[0], sh ? w[1] : 0u, sh);
}
__device__ __forceinline__ uint32_t ep_runlen_bytes(uint32_t x) { return x >= 15u ? 1u + (x - 15u) / 255u : 0u; }

/* length of the window-relative position idx: its run start's length minus the distance to it (>= 4) */
__device__ __forceinline__ int ep_len_of(const EncParSmem& S, int idx)
{
    const int st = S.start[idx];
    const int L = (int)S.len[st] - (idx - st);
    return L < 4 ? 4 : L;
}

__global__ void __launch_bounds__(kEpThreads, 1) lz4_encode_par_kernel(lz4k_encode_args a)
{
    extern __shared__ __align__(16) uint8_t smemRaw[];
    EncParSmem& S = *reinterpret_cast<EncParSmem*>(smemRaw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t parity = 0;
    const int accel = a.acceleration < 1 ? 1 : (a.acceleration > kAccelMax ? kAccelMax : a.acceleration);
    const int step = accel <= 4 ? 1 : 1 + (accel + 3) / 8;       /* every position for acceleration 1..4, then every 2nd (5..12), 3rd (13..20), ... */
    if (tid == 0) mbar_init(&S.mbar, 1);
    __syncthreads();
#ifdef LZ4K_PHASE_TIMING
    long long tPhase = clock64();
#endif

    for (int64_t b = blockIdx.x; b < a.nBlocks; b += gridDim.x) {
        const uint8_t* gsrc = a.src + b * a.srcStride;
        uint8_t* dst = a.dst + b * a.dstStride;
        const int n = a.srcSizeArr ? a.srcSizeArr[b] : a.srcSize;
        const int64_t cap = a.dstCap;
        if (n <= 0 || n > 65536) {                                    /* n == 0: one token byte (lz4.c:1361-1371); larger blocks are not this kernel's */
            if (tid == 0) { int r = 0; if (n == 0 && cap >= 1) { dst[0] = 0; r = 1; } a.outSize[b] = r; }
            continue;
        }
        const int head = (int)(reinterpret_cast<uintptr_t>(gsrc) & 15);
        const uint32_t loadBytes = (uint32_t)((head + n + 15) & ~15);
        if (tid == 0) {
            mbar_expect_tx(&S.mbar, loadBytes);
            for (uint32_t o = 0; o < loadBytes; o += 16384u) tma_load_1d(S.src + o, gsrc - head + o, min(16384u, loadBytes - o), &S.mbar);
            S.E = 0; S.O = 0; S.fail = 0; S.nJobs = 0; S.nLitJobs = 0;
        }
        for (int k = tid; k < (1 << kEpHashLog); k += kEpThreads) { S.T[k] = 0; S.T2[k] = 0; }
        __syncthreads();
        mbar_wait(&S.mbar, parity);
        parity ^= 1;
        const uint8_t* src = S.src;                                   /* byte i of the block = src[head + i] */
        const int mflimit = n - kMfLimit, matchlimit = n - kLastLiterals;
        const int nWin = (n >= kMinLength) ? (mflimit + kEpWin) / kEpWin : 0;      /* windows that hold positions <= mflimit */

        for (int w = 0; w < nWin; w++) {
            const int c0 = w * kEpWin, i0 = 4 * tid, p0 = c0 + i0;
            /* ---------------- find, part 1: hash, old candidate, publish "earliest of this window" ---------------- */
            uint32_t v[4], h[4], old[4], near[4];
            bool probe[4];
            {
                const uint32_t at = (uint32_t)(head + p0);
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(src) + (at >> 2);
                const bool any = p0 <= mflimit;
                const uint32_t wm = (any && p0 >= 4) ? wp[-1] : 0u;
                const uint32_t w0 = any ? wp[0] : 0u, w1 = any ? wp[1] : 0u, w2 = any ? wp[2] : 0u;
                const uint32_t sh = (at & 3u) * 8u;
                const uint32_t pre = __funnelshift_r(wm, w0, sh);                 /* bytes p0-4 .. p0-1 */
                const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
                v[0] = lo; v[1] = __funnelshift_r(lo, hi, 8); v[2] = __funnelshift_r(lo, hi, 16); v[3] = __funnelshift_r(lo, hi, 24);
                /* near[k] = smallest d in 1..4 with bytes [p-d, p-d+4) == [p, p+4) (a run of period d: RLE-like data), else 0.
                 * The hash tables hold ONE position per hash, so inside such a run neighbouring positions would get
                 * unrelated candidates; taking p-d keeps them on one offset (one run, one long match). */
                const uint32_t b[7] = {pre, __funnelshift_r(pre, lo, 8), __funnelshift_r(pre, lo, 16), __funnelshift_r(pre, lo, 24), v[0], v[1], v[2]};
                #pragma unroll
                for (int k = 0; k < 4; k++) {                                     /* b[4 + k - d] = the 4 bytes at p0 + k - d */
                    near[k] = 0;
                    #pragma unroll
                    for (int d = 4; d >= 1; d--)
                        if (p0 + k - d >= 0 && b[4 + k - d] == v[k]) near[k] = (uint32_t)d;
                }
            }
            if (tid < kEpWin / 32) { S.hasBits[tid] = 0; }
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                const int p = p0 + k;
                probe[k] = p <= mflimit && (step == 1 || p % step == 0);
                h[k] = (v[k] * 2654435761u) >> (32 - kEpHashLog);
                old[k] = 0;
                if (probe[k]) {
                    old[k] = S.T[h[k]];
                    atomicMax(&S.T2[h[k]], ((uint32_t)(w + 1) << 16) | (uint32_t)(0xFFFF - (i0 + k)));
                }
            }
            __syncthreads();
            PHASE_MARK(0);                                     // find 1
            /* ---------------- find, part 2: choose the candidate, enter the table ---------------- */
            uint32_t cnd[4];
            uint32_t hasMask = 0;
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                const int p = p0 + k;
                cnd[k] = 0xFFFFu;
                if (probe[k]) {
                    const uint32_t e = S.T2[h[k]];
                    const int q = c0 + (0xFFFF - (int)(e & 0xFFFFu));                 /* earliest position of this window with this hash */
                    if (near[k]) cnd[k] = (uint32_t)p - near[k];
                    else if (q < p && ep_ld32(src, (uint32_t)(head + q)) == v[k]) cnd[k] = (uint32_t)q;
                    else if (old[k] && ep_ld32(src, (uint32_t)head + old[k] - 1u) == v[k]) cnd[k] = old[k] - 1u;
                    atomicMax(&S.T[h[k]], (uint32_t)p + 1u);
                    if (cnd[k] != 0xFFFFu) hasMask |= 1u << k;
                }
                S.cand[i0 + k] = (uint16_t)cnd[k];
            }
            if (hasMask) atomicOr(&S.hasBits[i0 >> 5], hasMask << (i0 & 31));
            __syncthreads();
            PHASE_MARK(1);                                     // find 2
            /* ---------------- lengths: run starts extend, the others point at their run start ---------------- */
            {
                uint32_t prevC = (i0 > 0) ? S.cand[i0 - 1] : 0xFFFFu;             /* (a window's first position always starts a run) */
                uint32_t lastStart = 0;                                           /* window index + 1 of the latest run start in this quad */
                uint32_t longMask = 0;                                            /* run starts of this quad that are still matching after kEpSoloLen bytes */
                #pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int p = p0 + k;
                    const bool has = cnd[k] != 0xFFFFu;
                    const bool isStart = has && !(prevC != 0xFFFFu && cnd[k] == prevC + 1u);
                    if (isStart) {
                        /* longest match allowed here: the format's end-of-block rule, and the end of this window -- the rest of a
                         * longer match is found again by the next window (3 bytes per split), which bounds the work per window */
                        const int limit = min(matchlimit - p, c0 + kEpWin - p + kMinMatch);
                        int L = 4;
                        const uint32_t pa = (uint32_t)(head + p), ca = (uint32_t)head + cnd[k];
                        while (L < kEpSoloLen && L < limit) {
                            const uint32_t x = ep_ld32(src, pa + L) ^ ep_ld32(src, ca + L);
                            if (x) { L += (__ffs(x) - 1) >> 3; break; }
                            L += 4;
                        }
                        if (L > limit) L = limit;
                        if (L >= kEpSoloLen && L < limit) longMask |= 1u << k;    /* still matching: a warp finishes it */
                        S.len[i0 + k] = (uint16_t)L;
                        lastStart = (uint32_t)(i0 + k) + 1u;
                    }
                    prevC = cnd[k];
                }
                /* the long matches of the window, ranked by position (a scheduling-independent order): the first kEpMaxJobs
                 * are finished by a warp each, the others keep kEpSoloLen (valid, shorter) */
                uint32_t nl = (uint32_t)__popc(longMask), inclL = nl;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(kFull, inclL, d); if (lane >= d) inclL += y; }
                if (lane == 31) S.warpB[warp] = inclL;
                /* inclusive max-scan of lastStart over the threads: the run start at or before each quad's end */
                uint32_t m = lastStart;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(kFull, m, d); if (lane >= d) m = max(m, y); }
                if (lane == 31) S.warpA[warp] = m;
                __syncthreads();
                uint32_t before = 0, rankL = inclL - nl;                          /* latest run start / long matches in earlier warps */
                for (int q = 0; q < warp; q++) { before = max(before, S.warpA[q]); rankL += S.warpB[q]; }
                for (int k = 0; k < 4; k++)
                    if ((longMask >> k) & 1u) { if (rankL < (uint32_t)kEpMaxJobs) S.jobIdx[rankL] = (uint16_t)(i0 + k); rankL++; }
                if (tid == kEpThreads - 1) S.nJobs = rankL;
                uint32_t prevT = __shfl_up_sync(kFull, m, 1);
                if (lane == 0) prevT = 0;
                uint32_t run = max(before, prevT);                                /* latest run start before this quad */
                prevC = (i0 > 0) ? S.cand[i0 - 1] : 0xFFFFu;
                #pragma unroll
                for (int k = 0; k < 4; k++) {
                    const bool has = cnd[k] != 0xFFFFu;
                    const bool isStart = has && !(prevC != 0xFFFFu && cnd[k] == prevC + 1u);
                    if (isStart) run = (uint32_t)(i0 + k) + 1u;
                    S.start[i0 + k] = (uint16_t)(run ? run - 1u : (uint32_t)(i0 + k));
                    prevC = cnd[k];
                }
            }
            __syncthreads();
            PHASE_MARK(2);                                     // lengths
            /* ---------------- long matches: one warp per queued run start, 128 bytes per iteration ---------------- */
            {
                const uint32_t nJobs = min(S.nJobs, (uint32_t)kEpMaxJobs);
                for (uint32_t j = warp; j < nJobs; j += 32) {
                    const int idx = S.jobIdx[j], p = c0 + idx, limit = min(matchlimit - p, c0 + kEpWin - p + kMinMatch);
                    const uint32_t pa = (uint32_t)(head + p), ca = (uint32_t)head + S.cand[idx];
                    int L = kEpSoloLen;
                    for (;;) {
                        const int o = L + 4 * lane;
                        uint32_t x = 0;
                        bool stop = o >= limit;                                    /* at or past the allowed end: counts as a mismatch at o */
                        if (!stop) x = ep_ld32(src, pa + o) ^ ep_ld32(src, ca + o);
                        const unsigned mm = __ballot_sync(kFull, stop || x != 0u);
                        if (mm) {
                            const int f = __ffs(mm) - 1;
                            const uint32_t xf = __shfl_sync(kFull, x, f);
                            const bool sf = __shfl_sync(kFull, (int)stop, f) != 0;
                            L = L + 4 * f + (sf ? 0 : ((__ffs(xf) - 1) >> 3));
                            break;
                        }
                        L += 128;
                    }
                    if (L > limit) L = limit;
                    if (lane == 0) S.len[idx] = (uint16_t)L;
                }
            }
            __syncthreads();
            PHASE_MARK(3);                                     // long matches
            /* ---------------- select: the greedy chain through this window (warp 0, 128 positions per lane) ---------------- */
            if (warp == 0) {
                uint32_t has[4], sel[4];
                #pragma unroll
                for (int q = 0; q < 4; q++) has[q] = S.hasBits[4 * lane + q];
                const int segLo = 128 * lane, segHi = segLo + 128;
                const int Ein = (int)S.E;
                int eCur = Ein, exitE = Ein, firstSel = -1;
                auto walk = [&](int e) {                       /* chain enters with "end of the last match" = e */
                    sel[0] = sel[1] = sel[2] = sel[3] = 0; firstSel = -1;
                    int rel = e - c0;
                    if (rel < segLo) rel = segLo;
                    while (rel < segHi) {
                        int q = (rel - segLo) >> 5;
                        uint32_t mword = has[q] & (0xFFFFFFFFu << (rel & 31));
                        while (mword == 0u && ++q < 4) mword = has[q];
                        if (q >= 4) break;
                        const int idx = segLo + 32 * q + (__ffs(mword) - 1);
                        sel[q] |= 1u << (idx & 31);
                        S.litStart[idx] = (uint16_t)e;
                        if (firstSel < 0) firstSel = idx;
                        e = c0 + idx + ep_len_of(S, idx);
                        rel = e - c0;
                    }
                    exitE = e;
                };
                walk(eCur);
                int finalExit = Ein;
                for (;;) {
                    /* exit of lanes 0..j = the end of the last match selected at or before lane j (lanes that select nothing pass
                     * the chain through): a "last valid value" scan instead of one round per lane */
                    int val = (firstSel >= 0) ? exitE : -1;
                    #pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(kFull, val, d); if (lane >= d && val < 0) val = y; }
                    const int incl = (val < 0) ? Ein : val;
                    int eNew = __shfl_up_sync(kFull, incl, 1);
                    if (lane == 0) eNew = Ein;
                    const bool need = max(eNew - c0, segLo) != max(eCur - c0, segLo);      /* the search would start elsewhere */
                    eCur = eNew;
                    if (!__any_sync(kFull, need)) { finalExit = __shfl_sync(kFull, incl, 31); break; }
                    if (need) walk(eNew);
                }
                exitE = finalExit;
                if (firstSel >= 0) S.litStart[firstSel] = (uint16_t)eCur;          /* its literals start at the true entry */
                #pragma unroll
                for (int q = 0; q < 4; q++) S.selBits[4 * lane + q] = sel[q];
                if (lane == 31) S.E = (uint32_t)exitE;
            }
            __syncthreads();
            PHASE_MARK(4);                                     // select
            /* ---------------- emit: <= 1 sequence per thread ---------------- */
            {
                const uint32_t mine = (S.selBits[i0 >> 5] >> (i0 & 31)) & 0xFu;
                int p = 0, A = 0, L = 0, ll = 0;
                uint32_t c = 0, size = 0;
                if (mine) {
                    const int idx = i0 + (__ffs(mine) - 1);
                    p = c0 + idx; A = S.litStart[idx]; c = S.cand[idx]; L = ep_len_of(S, idx);
                    while (p > A && c > 0u && src[head + p - 1] == src[head + c - 1u]) { p--; c--; L++; }      /* lz4.c:1107-1109 */
                    ll = p - A;
                    size = 1u + ep_runlen_bytes((uint32_t)ll) + (uint32_t)ll + 2u + ep_runlen_bytes((uint32_t)(L - kMinMatch));
                }
                uint32_t incl = size;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(kFull, incl, d); if (lane >= d) incl += y; }
                if (lane == 31) S.warpB[warp] = incl;
                __syncthreads();
                const uint32_t O0 = S.O;
                uint32_t base = O0, winTotal = 0;
                for (int q = 0; q < 32; q++) { const uint32_t x = S.warpB[q]; winTotal += x; if (q < warp) base += x; }
                /* scattered byte stores to global memory cost one memory transaction each: the window's output is put
                 * together in shared memory and written out as consecutive bytes (windows whose output does not fit --
                 * a match after a very long literal run -- are written directly) */
                const bool staged = winTotal <= (uint32_t)kEpStage;
                uint8_t* const obase = staged ? S.stage - O0 : dst;             /* output offset o lives at obase + o */
                if (mine) {
                    int64_t o = (int64_t)base + incl - size;
                    if (o + size > cap) S.fail = 1;
                    else {
                        uint8_t* d = obase + o;
                        const uint32_t ml = (uint32_t)(L - kMinMatch);
                        *d++ = (uint8_t)((min((uint32_t)ll, 15u) << 4) | min(ml, 15u));
                        if (ll >= 15) { uint32_t r = (uint32_t)ll - 15u; while (r >= 255u) { *d++ = 255; r -= 255u; } *d++ = (uint8_t)r; }
                        if (ll <= kEpInlineLits) {
                            for (int i = 0; i < ll; i++) d[i] = src[head + A + i];
                        } else {
                            const uint32_t j = atomicAdd(&S.nLitJobs, 1u);         /* (at most 4096/33 such runs end in a window) */
                            S.litJob[j][0] = (uint32_t)A; S.litJob[j][1] = (uint32_t)(d - obase); S.litJob[j][2] = (uint32_t)ll;
                        }
                        d += ll;
                        const uint32_t off = (uint32_t)p - c;
                        *d++ = (uint8_t)off; *d++ = (uint8_t)(off >> 8);
                        if (ml >= 15u) { uint32_t r = ml - 15u; while (r >= 255u) { *d++ = 255; r -= 255u; } *d++ = (uint8_t)r; }
                    }
                }
                __syncthreads();
                if (tid == 0) S.O = O0 + winTotal;
                const uint32_t nLit = S.nLitJobs;
                for (uint32_t j = warp; j < nLit; j += 32) {
                    const uint32_t from = S.litJob[j][0], to = S.litJob[j][1], cnt = S.litJob[j][2];
                    for (uint32_t i = lane; i < cnt; i += 32) obase[to + i] = src[head + from + i];
                }
                __syncthreads();
                if (staged && !S.fail)                                           /* coalesced: consecutive threads, consecutive bytes */
                    for (uint32_t i = tid; i < winTotal; i += kEpThreads) dst[O0 + i] = S.stage[i];
                if (tid == 0) S.nLitJobs = 0;
                __syncthreads();
                PHASE_MARK(5);                                 // emit
            }
        }
        /* ---------------- last literals (lz4.c:1302-1329) ---------------- */
        __syncthreads();
        {
            const uint32_t E = S.E, O = S.O;
            const uint32_t last = (uint32_t)n - E;
            const int64_t total = (int64_t)O + 1 + ep_runlen_bytes(last) + last;
            const bool ok = !S.fail && total <= cap;
            if (ok) {
                uint8_t* d = dst + O;
                const uint32_t skip = 1u + ep_runlen_bytes(last);
                if (tid == 0) {
                    *d++ = (uint8_t)(min(last, 15u) << 4);
                    if (last >= 15u) { uint32_t r = last - 15u; while (r >= 255u) { *d++ = 255; r -= 255u; } *d++ = (uint8_t)r; }
                }
                for (uint32_t i = tid; i < last; i += kEpThreads) dst[O + skip + i] = src[head + E + i];
            }
            if (tid == 0) a.outSize[b] = ok ? (int32_t)total : 0;
        }
        __syncthreads();                                               /* S is reused by the next block */
        PHASE_MARK(6);                                     // load + tables + last literals
    }
}

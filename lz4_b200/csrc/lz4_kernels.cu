/*
 * lz4_kernels.cu -- hand-written sm_100a CUDA kernels of the LZ4 block codec.
 *
 * Decode = two kernels:
 *   scan   : one THREAD per block walks the token chain and applies every acceptance rule of the
 *            reference decoder (lz4.c:2022-2445, x86-64 control flow incl. the fast loop), giving
 *            the exact return value, and leaves a 4-byte mark per sequence.  No data is moved.
 *   expand : moves the bytes of blocks the scan accepted (literal + match copies).
 * Encode = one WARP per block replaying the reference's greedy parse (lz4.c:930-1338) with the
 *   32 lanes probing 32 consecutive search positions per step; output is byte-identical.
 * pack   = exclusive scan of block sizes + gather into one contiguous stream.
 *
 * All arithmetic is integer/byte; no tensor cores (this is an HBM/latency-bound scan codec).
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <stddef.h>
#include "lz4_kernels.h"

namespace {

constexpr int kMinMatch = 4;
constexpr int kLastLiterals = 5;
constexpr int kMfLimit = 12;
constexpr int kMinLength = 13;
constexpr uint32_t kMaxDistance = 65535;
constexpr int kSmallLimit = 65536 + 11;          // lz4.c:710
constexpr int kSkipTrigger = 6;                  // lz4.c:711
constexpr int kAccelMax = 65537;                 // lz4.c:58
constexpr uint32_t kMaxInput = 0x7E000000u;      // lz4.h:214
constexpr unsigned kFull = 0xFFFFFFFFu;

unsigned long long g_launches = 0;

/* optional phase timing of the fast expand kernel (thread 0 of each CTA; enabled with -DLZ4K_PHASE_TIMING) */
__device__ unsigned long long g_phaseCycles[8];
__device__ unsigned long long g_loopStats[4];    // phase B: warp iterations, lane-iterations with a piece, blocked lane-iterations, pieces done
#ifdef LZ4K_PHASE_TIMING
#define PHASE_MARK(i) do { if (tid == 0) { const long long t_ = clock64(); atomicAdd(&g_phaseCycles[i], (unsigned long long)(t_ - tPhase)); tPhase = t_; } } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#endif

/* byte readers + the per-block scan (scan_front / scan_tail / scan_block): shared with the CPU
 * emulators under tests/emul/, see the header */
#define LZ4_SCAN_CORE_CONSTANTS
#include "lz4_scan_core.h"
#include "lz4_scan_par.h"
#include "lz4_scan_split.h"
#include "lz4_rows_core.h"

/* low 5 bytes at p (for the 5-byte hash, lz4.c:785-791) */
__device__ __forceinline__ uint64_t ld40u(const uint8_t* p)
{
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t lo = __ldg(w);
    uint32_t hi = __ldg(w + 1);
    uint32_t v = __funnelshift_r(lo, hi, sh);
    uint32_t b4 = (hi >> sh) & 0xFFu;
    return (uint64_t)v | ((uint64_t)b4 << 32);
}

constexpr int kInBytes = 65536 + 64;           /* staged compressed block: <= 65535 bytes + 16-byte phase + rounding */

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
/* wait for the phase with the given parity; the hardware may suspend the warp for up to `hintNs` per attempt, so that
 * waiting warps do not take issue slots from the warps they are waiting for */
#ifndef LZ4K_WAIT_HINT_NS
#define LZ4K_WAIT_HINT_NS 2000
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)LZ4K_WAIT_HINT_NS) : "memory");
    } while (!ok);
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_1d(void* smemDst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smemDst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* smemSrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(smemSrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }



/* workspace layout (lz4k_decode_workspace_bytes): header | nSeq[N] | slowList[N] | marks[N][markStride] */
struct WsHeader { uint32_t slowCount, pad[3]; };

struct WsView {
    WsHeader* hdr; uint32_t* nSeq; uint32_t* slowList; uint32_t* marks; uint32_t* scratch; uint32_t markStride;
    uint32_t* tileFirst; uint32_t* tileFlag; uint32_t tilesMax;    /* batches of blocks above 64 KB (wide marks): see lz4_expand_tiles_kernel */
};
constexpr int kTile = 61440;                       /* output bytes per tile of a block above 64 KB: 15 waves of 4096 */
constexpr int32_t kTileMaxCap = 64 << 20;          /* larger capacities take the generic kernel (the marks would need 2 x capacity) */
__host__ __device__ inline bool wide_batch(const int32_t* dstCapArr, int32_t dstCap) { return !dstCapArr && dstCap > 65536 && dstCap <= kTileMaxCap; }
__host__ __device__ inline uint32_t tiles_of(int32_t bytes) { return ((uint32_t)bytes + (uint32_t)kTile - 1u) / (uint32_t)kTile; }
/* mark slots per block: a block the shared-memory expand kernel may take (capacity <= 64 KB) has at most
 * capacity/4 + 1 sequences (every sequence but the last produces >= 4 bytes) and the scan visits at most one more;
 * batches of larger blocks need no marks at all */
__host__ __device__ inline uint32_t mark_stride(const int32_t* dstCapArr, int32_t dstCap, bool wide = false)
{
    if (dstCapArr) return (uint32_t)kMaxSeqFast;               /* per-block capacities live on the device: worst case */
    if (wide) return 2u * ((uint32_t)dstCap / 4u + 2u);        /* wide marks: two words per sequence */
    if (dstCap <= 0 || dstCap > 65536) return 0u;
    const uint32_t s = (uint32_t)dstCap / 4u + 2u;
    return s < (uint32_t)kMaxSeqFast ? s : (uint32_t)kMaxSeqFast;
}
__host__ __device__ inline WsView ws_view(void* ws, int64_t n, uint32_t markStride, uint32_t tilesMax = 0)
{
    WsView v;
    uint8_t* p = reinterpret_cast<uint8_t*>(ws);
    v.hdr = reinterpret_cast<WsHeader*>(p); p += 256;
    v.nSeq = reinterpret_cast<uint32_t*>(p); p += ((n * 4 + 255) / 256) * 256;
    v.slowList = reinterpret_cast<uint32_t*>(p); p += ((n * 4 + 255) / 256) * 256;
    v.marks = reinterpret_cast<uint32_t*>(p); p += (((size_t)n * markStride * 4 + 255) / 256) * 256;
    v.scratch = reinterpret_cast<uint32_t*>(p);                /* split scan: 2 x markStride words per block */
    v.markStride = markStride;
    /* wide batches have no scratch lists: per block tilesMax + 1 first-sequence indices, tilesMax done flags + 1 "gave up" word */
    v.tileFirst = reinterpret_cast<uint32_t*>(p); p += (((size_t)n * (tilesMax + 1) * 4 + 255) / 256) * 256;
    v.tileFlag = reinterpret_cast<uint32_t*>(p);
    v.tilesMax = tilesMax;
    return v;
}
__host__ __device__ inline size_t ws_bytes(int64_t nBlocks, uint32_t markStride, uint32_t tilesMax = 0)
{
    const size_t lst = (((size_t)nBlocks * 4 + 255) / 256) * 256;
    const size_t mk = (((size_t)nBlocks * markStride * sizeof(uint32_t) + 255) / 256) * 256;
    if (tilesMax) {                                            /* header | nSeq | slowList | wide marks | tileFirst | tileFlag */
        const size_t tl = (((size_t)nBlocks * (tilesMax + 1) * 4 + 255) / 256) * 256;
        return 256 + 2 * lst + mk + 2 * tl + 256;
    }
    return 256 + 2 * lst + 3 * mk + 256;                      /* header | nSeq | slowList | marks | scratch (2 x marks) */
}
/* A batch of blocks above 64 KB is decoded in tiles when the caller's workspace has room for the wide marks
 * (LZ4B200_decompress_workspace_bytes_for says how much); with the small workspace it takes the generic kernel. */
__host__ __device__ inline bool use_wide(const int32_t* dstCapArr, int32_t dstCap, int64_t nBlocks, size_t workspaceBytes)
{
    return wide_batch(dstCapArr, dstCap) && workspaceBytes >= ws_bytes(nBlocks, mark_stride(dstCapArr, dstCap, true), tiles_of(dstCap));
}
__device__ __forceinline__ WsView ws_view(const lz4k_decode_args& a)
{
    const bool wide = use_wide(a.dstCapArr, a.dstCap, a.nBlocks, a.workspaceBytes);
    return ws_view(a.workspace, a.nBlocks, mark_stride(a.dstCapArr, a.dstCap, wide), wide ? tiles_of(a.dstCap) : 0u);
}

/* a block the shared-memory expand kernel takes: input and output fit its 64 KB windows, marks for every sequence */
__device__ __forceinline__ bool rows_eligible(int n, int cap, uint32_t nseq, uint32_t markStride)
{
    return n > 0 && n <= 65535 && cap > 0 && cap <= 65536 && nseq <= (uint32_t)kMaxSeqFast && nseq <= markStride;
}

/* ---- scan, one THREAD per block: reads through the read-only data cache with L1 prefetch hints ---- */

__global__ void __launch_bounds__(128) lz4_scan_kernel(lz4k_decode_args a)
{
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.nBlocks) return;
    const WsView w = ws_view(a);
    const uint8_t* src = a.src + a.srcOff[b];
    const int n = a.srcSize[b];
    int cap = a.dstCapArr ? a.dstCapArr[b] : a.dstCap;
    uint32_t ns = 0;
    const bool wantMarks = (n > 0 && n <= 65535 && cap > 0 && cap <= 65536 && w.markStride > 0);
    uint32_t* marks = wantMarks ? (w.marks + b * w.markStride) : nullptr;
    MemPtr<true> mem{src};
    const int r = scan_block(mem, n, cap, &ns, marks, w.markStride);
    a.outSize[b] = r;
    w.nSeq[b] = ns;
    if (r > 0 && !rows_eligible(n, cap, ns, w.markStride)) w.slowList[atomicAdd(&w.hdr->slowCount, 1u)] = (uint32_t)b;
}

/* ---- scan, kSsLanes lanes per block that merge (lz4_scan_split.h): the chain of dependent steps is kSsLanes times shorter ----
 * The lanes of a block are neighbours in a warp (32 / kSsLanes blocks per warp); phases are separated by __syncwarp.
 * Lists go to the scratch part of the workspace (2 x markStride words per block), the per-block records to shared memory. */
constexpr int kSplitThreads = 128;
constexpr int kSplitBlocksPerCta = kSplitThreads / kSsLanes;
constexpr int64_t kSplitMaxBlocks = 16384;        /* largest batch the split scan is the default for */

__global__ void __launch_bounds__(kSplitThreads, 10) lz4_scan_split_kernel(lz4k_decode_args a)
{
    __shared__ SsBlock sh[kSplitBlocksPerCta];
    const int lane = threadIdx.x % kSsLanes, slot = threadIdx.x / kSsLanes;
    const int64_t b = (int64_t)blockIdx.x * kSplitBlocksPerCta + slot;
    if (b >= a.nBlocks) return;                                 /* (whole groups of kSsLanes lanes leave together) */
    const unsigned grp = ((1u << kSsLanes) - 1u) << ((threadIdx.x & 31) / kSsLanes * kSsLanes);   /* this block's lanes in the warp */
    const WsView w = ws_view(a);
    const uint8_t* src = a.src + a.srcOff[b];
    const int n = a.srcSize[b];
    const int cap = a.dstCapArr ? a.dstCapArr[b] : a.dstCap;
    const bool wantMarks = (n > 0 && n <= 65535 && cap > 0 && cap <= 65536 && w.markStride > 0);
    uint32_t* marks = wantMarks ? (w.marks + b * w.markStride) : nullptr;
    SsBlock& S = sh[slot];
    MemPtr<true> mem{src};
    int r = 0;
    uint32_t ns = 0;
    bool serial = !(wantMarks && cap >= 64 && n >= kSsMinBytes);
    if (!serial) {
        const uint32_t R = w.markStride / kSsLanes;
        uint32_t* sc = w.scratch + b * 2 * (int64_t)w.markStride;
        uint32_t* A = sc + (size_t)lane * R;
        uint32_t* B = sc + w.markStride + (size_t)lane * R;
        ss_p1(lane, S, mem, n, A, B, R);
        __syncwarp(grp);
        ss_p2(lane, S, mem, n, A, B, R, [&](int t) { return sc + w.markStride + (size_t)t * R; });
        __syncwarp(grp);
        if (lane == 0) ss_p3(S, [&](int t) { return sc + (size_t)t * R; });
        __syncwarp(grp);
        serial = S.fallback != 0;
        if (!serial) {
            uint32_t e, c, co;
            ss_p4(lane, S, cap, A, B, marks, w.markStride, e, c, co);
            #pragma unroll
            for (int m = 1; m < kSsLanes; m <<= 1) {
                const uint32_t e2 = __shfl_xor_sync(grp, e, m), c2 = __shfl_xor_sync(grp, c, m), co2 = __shfl_xor_sync(grp, co, m);
                if (e2 < e) e = e2;
                if (c2 < c) { c = c2; co = co2; }
            }
            if (lane == 0) { S.errIdx = e; S.capIdx = c; S.capOpn = co; }
            __syncwarp(grp);                                        /* the final marks of every lane are written */
            if (lane == 0) r = ss_p5(S, mem, n, cap, &ns, marks, w.markStride);
        }
    }
    if (serial && lane == 0) r = scan_block(mem, n, cap, &ns, marks, w.markStride);
    if (lane == 0) {
        a.outSize[b] = r;
        w.nSeq[b] = ns;
        if (r > 0 && !rows_eligible(n, cap, ns, w.markStride)) w.slowList[atomicAdd(&w.hdr->slowCount, 1u)] = (uint32_t)b;
    }
}

/* ---- scan, one CTA of NL lanes per block (lz4_scan_par.h): same outputs as lz4_scan_kernel ----
 * Blocks of up to 65 535 bytes are staged in shared memory by one TMA bulk load (the walks are chains of
 * dependent 4-byte reads: ~30 cycles each from shared memory instead of an L2 / HBM round trip); larger
 * blocks (lz4frame's 4 MB blocks) are walked in global memory by the same lanes. */
#ifndef LZ4K_SCAN_LANES
#define LZ4K_SCAN_LANES 128
#endif
constexpr int kScanLanes = LZ4K_SCAN_LANES;
static_assert(kScanLanes % 32 == 0 && kScanLanes <= kSpMaxLanes, "scan lanes");

struct ScanParSmem {
    alignas(16) uint8_t in[65536 + 64];
    SpShared sp;
    uint32_t warpCnt[kSpMaxLanes / 32], warpLen[kSpMaxLanes / 32];
    int first;
    alignas(8) uint64_t mbar;
};

template <class M>
__device__ __forceinline__ void scan_par_block(ScanParSmem& S, M& mem, int n, int cap, uint32_t* marks, uint32_t markCap,
                                               int& ret, uint32_t& nseq)
{
    const int lane = threadIdx.x, nl = kScanLanes;
    if (cap < 64 || n < kSpMinBytes) {
        if (lane == 0) { uint32_t ns = 0; S.sp.ret = scan_block(mem, n, cap, &ns, marks, markCap); S.sp.nseq = ns; }
        __syncthreads();
        ret = S.sp.ret; nseq = S.sp.nseq;
        __syncthreads();
        return;
    }
    SpLane L;
    sp_phase0(lane, nl, L, S.sp, mem, n, cap);
    __syncthreads();
    for (;;) {
        sp_decide(lane, L, S.sp);
        __syncthreads();                                   /* everybody has read S.changed's previous value and its neighbour's result */
        if (lane == 0) S.sp.changed = 0;
        __syncthreads();
        sp_redo(lane, L, S.sp, mem, n, cap);
        __syncthreads();
        if (!S.sp.changed) break;
    }
    /* exclusive sums of (count, olen) over the lanes */
    uint32_t c = S.sp.res[lane].count, o = S.sp.res[lane].olen, ci = c, oi = o;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t yc = __shfl_up_sync(kFull, ci, d), yo = __shfl_up_sync(kFull, oi, d);
        if ((lane & 31) >= d) { ci += yc; oi += yo; }
    }
    if ((lane & 31) == 31) { S.warpCnt[lane >> 5] = ci; S.warpLen[lane >> 5] = oi; }
    if (lane == 0) S.first = nl - 1;
    __syncthreads();
    uint32_t cb = ci - c, ob = oi - o;
    for (int wq = 0; wq < (lane >> 5); wq++) { cb += S.warpCnt[wq]; ob += S.warpLen[wq]; }
    sp_write(lane, L, S.sp, mem, n, cap, cb, ob, marks, markCap);
    if (S.sp.end[lane].kind != SP_RAN) atomicMin(&S.first, lane);
    __syncthreads();
    sp_finish(lane, S.first, S.sp, mem, n, cap, marks, markCap);
    __syncthreads();
    ret = S.sp.ret; nseq = S.sp.nseq;
    __syncthreads();
}

__global__ void __launch_bounds__(kScanLanes) lz4_scan_par_kernel(lz4k_decode_args a)
{
    extern __shared__ __align__(16) uint8_t smemRaw[];
    ScanParSmem& S = *reinterpret_cast<ScanParSmem*>(smemRaw);
    const WsView w = ws_view(a);
    const int tid = threadIdx.x;
    uint32_t parity = 0;
    if (tid == 0) mbar_init(&S.mbar, 1);
    __syncthreads();
    for (int64_t b = blockIdx.x; b < a.nBlocks; b += gridDim.x) {
        const uint8_t* src = a.src + a.srcOff[b];
        const int n = a.srcSize[b];
        const int cap = a.dstCapArr ? a.dstCapArr[b] : a.dstCap;
        const bool wide = w.tilesMax != 0;                            /* a batch of blocks above 64 KB: two-word marks for the tiles kernel */
        const bool wantMarks = wide ? n > 0 : (n > 0 && n <= 65535 && cap > 0 && cap <= 65536 && w.markStride > 0);
        uint32_t* marks = wantMarks ? (w.marks + (size_t)b * w.markStride) : nullptr;
        const uint32_t markCap = wide ? w.markStride / 2u : w.markStride;
        int r = -1;
        uint32_t ns = 0;
        if (wide && n > 0 && n <= 65535) {
            const int head = (int)(reinterpret_cast<uintptr_t>(src) & 15);
            const uint32_t loadBytes = (uint32_t)((head + n + 15) & ~15);
            if (tid == 0) {
                mbar_expect_tx(&S.mbar, loadBytes);
                for (uint32_t o = 0; o < loadBytes; o += 16384u)
                    tma_load_1d(S.in + o, src - head + o, min(16384u, loadBytes - o), &S.mbar);
            }
            while (!mbar_try_wait(&S.mbar, parity)) { }
            parity ^= 1;
            MemPtr<false, true> mem{S.in + head};
            scan_par_block(S, mem, n, cap, marks, markCap, r, ns);
        } else if (wide && n > 65535) {
            MemPtr<true, true> mem{src};
            scan_par_block(S, mem, n, cap, marks, markCap, r, ns);
        } else if (n > 0 && n <= 65535 && cap > 0) {
            const int head = (int)(reinterpret_cast<uintptr_t>(src) & 15);
            const uint32_t loadBytes = (uint32_t)((head + n + 15) & ~15);
            if (tid == 0) {
                mbar_expect_tx(&S.mbar, loadBytes);
                for (uint32_t o = 0; o < loadBytes; o += 16384u)
                    tma_load_1d(S.in + o, src - head + o, min(16384u, loadBytes - o), &S.mbar);
            }
            while (!mbar_try_wait(&S.mbar, parity)) { }
            parity ^= 1;
            MemPtr<false> mem{S.in + head};
            scan_par_block(S, mem, n, cap, marks, w.markStride, r, ns);
        } else if (n > 65535 && cap >= 64) {
            MemPtr<true> mem{src};
            scan_par_block(S, mem, n, cap, marks, w.markStride, r, ns);
        } else {                                               /* degenerate arguments: the one-thread code decides (lz4.c:2036, :2064-2069) */
            if (tid == 0) { uint32_t q = 0; MemPtr<true> mem{src}; S.sp.ret = scan_block(mem, n, cap, &q, nullptr, 0u); S.sp.nseq = q; }
            __syncthreads();
            r = S.sp.ret; ns = S.sp.nseq;
            __syncthreads();
        }
        if (tid == 0) {
            a.outSize[b] = r;
            w.nSeq[b] = ns;
            const bool tiled = wide && ns <= markCap;                 /* lz4_expand_tiles_kernel takes it */
            if (r > 0 && !tiled && !rows_eligible(n, cap, ns, w.markStride)) w.slowList[atomicAdd(&w.hdr->slowCount, 1u)] = (uint32_t)b;
        }
    }
}

/* ---- wide batches: per tile of kTile output bytes the first sequence whose match starts in or beyond it ----
 * tileFirst[b][t] = min { k : matchStart(k) >= t * kTile } for t = 0 .. tiles(b) (nseq if none); done flags cleared.
 * The match starts are non-decreasing in k and the last mark holds the decoded size, so sequence k's neighbours decide. */
__global__ void __launch_bounds__(256) lz4_tile_index_kernel(lz4k_decode_args a)
{
    const WsView w = ws_view(a);
    for (int64_t b = blockIdx.x; b < a.nBlocks; b += gridDim.x) {
        uint32_t* first = w.tileFirst + (size_t)b * (w.tilesMax + 1);
        uint32_t* flag = w.tileFlag + (size_t)b * (w.tilesMax + 1);
        for (uint32_t t = threadIdx.x; t <= w.tilesMax; t += blockDim.x) flag[t] = 0u;
        const int total = a.outSize[b];
        const uint32_t ns = w.nSeq[b];
        if (total <= 0 || ns == 0 || ns > w.markStride / 2u) continue;
        const uint32_t nT = tiles_of(total);
        const uint32_t* marks = w.marks + (size_t)b * w.markStride;
        for (uint32_t k = threadIdx.x; k < ns; k += blockDim.x) {
            const uint32_t m = marks[2 * (size_t)k + 1];
            const uint32_t t1 = min(m / (uint32_t)kTile, nT);             /* tiles t <= t1 start at or below m */
            uint32_t t0 = 0;                                                /* first tile not yet served by k - 1 */
            if (k) t0 = min(marks[2 * (size_t)k - 1] / (uint32_t)kTile, nT) + 1u;
            for (uint32_t t = t0; t <= t1; t++) first[t] = k;
            if (k + 1 == ns) for (uint32_t t = t1 + 1; t <= nT; t++) first[t] = ns;
        }
    }
}

/* =============================================================================================
 * expand (generic): one warp per accepted block, any block size, straight to global memory
 * ============================================================================================= */
__global__ void __launch_bounds__(128) lz4_expand_generic_kernel(lz4k_decode_args a)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const WsView w = ws_view(a);
    const int64_t count = w.hdr->slowCount;       // only blocks the scan accepted and left to this kernel
    for (int64_t idx = warp0; idx < count; idx += nWarps) {
    const int64_t b = w.slowList[idx];
    const uint8_t* __restrict__ src = a.src + a.srcOff[b];
    uint8_t* dst = a.dst + (a.dstOff ? a.dstOff[b] : b * a.dstStride);
    const int64_t n = a.srcSize[b];
    int64_t ip = 0, op = 0;

    for (;;) {
        uint32_t token = ldb<true>(src + ip); ip++;
        int64_t ll = token >> 4;
        if (ll == 15) { uint32_t x; do { x = ldb<true>(src + ip); ip++; ll += x; } while (x == 255); }
        for (int64_t k = lane; k < ll; k += 32) dst[op + k] = (uint8_t)ldb<true>(src + ip + k);
        ip += ll; op += ll;
        if (ip >= n) break;
        uint32_t offset = ld16<true>(src + ip); ip += 2;
        int64_t ml = token & 15;
        if (ml == 15) { uint32_t x; do { x = ldb<true>(src + ip); ip++; ml += x; } while (x == 255); }
        ml += kMinMatch;
        __syncwarp();                             // everything before `op` is now visible to all lanes
        if (offset == 0) {                        // reference zero-fills (lz4.c:2407, :500)
            for (int64_t k = lane; k < ml; k += 32) dst[op + k] = 0;
        } else if ((int64_t)offset >= ml) {
            const volatile uint8_t* from = dst + op - offset;
            for (int64_t k = lane; k < ml; k += 32) dst[op + k] = from[k];
        } else {                                  // self-overlapping match: period `offset`
            const volatile uint8_t* from = dst + op - offset;
            for (int64_t k = lane; k < ml; k += 32) dst[op + k] = from[k % offset];
        }
        op += ml;
    }
    __syncwarp();
    }
}

/* =============================================================================================
 * expand (rows): one CTA per 64 KB block, one output BYTE per thread, waves + run table
 *
 * See lz4_rows_core.h for the formulation.  Per block:
 *   TMA bulk load  : compressed block HBM -> smem (`in`, placed below `out` in the window)
 *   runs, pass 1   : one lane per sequence re-reads its token (marks from the scan) and sets one bit
 *                    per run start (literal run, match run, pieces of self-overlapping matches)
 *   rank           : rows[r].y = (run starts before row r) - 1
 *   runs, pass 2   : the same lanes write delta(run) at the run's rank
 *   waves          : kWave bytes per CTA barrier; thread t handles bytes t, t+1024, ... of the wave:
 *                    rank -> delta -> (follow sources inside the wave) -> LDS.U8 -> STS.U8
 *   TMA bulk store : decoded block smem -> HBM
 * ============================================================================================= */
#ifndef LZ4K_ROWS_RPT
#define LZ4K_ROWS_RPT 4                      /* rows per thread per wave: wave = 1024 * RPT bytes */
#endif
constexpr int kRowsThreads = 1024;
constexpr int kRowsRpt = LZ4K_ROWS_RPT;
constexpr int kWave = kRowsThreads * kRowsRpt;
constexpr int kRowsCache = 3;                /* sequences per thread whose parse is kept in registers between the passes */

struct RowsDesc {                            /* one block's arguments, fetched a block ahead */
    const uint8_t* src;
    uint8_t* dst;
    int n, total, nseq;
    int64_t b;
};

struct RowsSmem {
    alignas(16) uint8_t zero[16];            /* always-zero cell: source of offset-0 matches (lz4.c:2407) */
    alignas(16) uint8_t in[kInBytes];        /* staged compressed block; keeps the source's 16-byte phase */
    alignas(16) uint8_t out[65536];          /* output window; MUST lie above `zero` and `in` */
    alignas(16) uint32_t tab[kRowsMaxRuns];  /* delta per run */
    alignas(8) uint2 rows[2048];             /* {run-start bits of the row, run starts before the row - 1} */
    uint32_t warpSum[32];
    alignas(8) uint64_t mbar;                /* TMA load of `in` */
    alignas(8) uint64_t wbar;                /* wave barrier: one arrival per warp */
    RowsDesc desc[2];
    uint32_t nRuns;
};
static_assert(sizeof(RowsSmem) <= 232448, "RowsSmem exceeds the 227 KB of shared memory a CTA can opt in to");

__device__ __forceinline__ uint2 lds_u64(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ uint32_t lanemask_le()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}

/* Source addresses of one thread's bytes of a wave: bytes p0 + r*1024 (r < kRowsRpt) of the wave that starts at shared
 * address waveS.  FULL: the whole wave lies inside the block (no bounds checks).  All addresses are 32-bit
 * shared-window addresses.  Nothing here reads the output window: the run table alone decides where a byte comes
 * from, so a warp resolves wave w+1 while other warps still copy wave w. */
template <bool FULL>
__device__ __forceinline__ void rows_resolve(uint32_t (&sa)[LZ4K_ROWS_RPT], const uint32_t p0, const uint32_t waveS,
                                             const uint32_t lim, const uint32_t outS, const uint32_t rowsS,
                                             const uint32_t tabS, const uint32_t zeroS, const uint32_t le)
{
    const uint32_t rowAddr = rowsS + ((p0 >> 5) << 3);
    #pragma unroll
    for (int r = 0; r < LZ4K_ROWS_RPT; r++) {
        sa[r] = zeroS;
        if (FULL || p0 + (uint32_t)(r * 1024) < lim) {
            const uint2 row = lds_u64(rowAddr + (uint32_t)(r * 32 * 8));
            const uint32_t j = row.y + (uint32_t)__popc(row.x & le);
            sa[r] = outS + p0 + (uint32_t)(r * 1024) + lds_u32(tabS + (j << 2));
        }
    }
    uint32_t hi = sa[0];
    #pragma unroll
    for (int r = 1; r < LZ4K_ROWS_RPT; r++) hi = max(hi, sa[r]);
    while (hi >= waveS) {                                      /* sources inside this wave: follow them, all rows of the thread at once */
        hi = 0;
        #pragma unroll
        for (int r = 0; r < LZ4K_ROWS_RPT; r++) {
            uint32_t x = sa[r];
            if (x >= waveS) {
                const uint32_t q = x - outS;
                const uint2 row = lds_u64(rowsS + ((q >> 5) << 3));
                const uint32_t j = row.y + (uint32_t)__popc(row.x & (0xFFFFFFFFu >> (31u - (q & 31u))));
                x += lds_u32(tabS + (j << 2));
                sa[r] = x;
            }
            hi = max(hi, x);
        }
    }
}

__global__ void __launch_bounds__(kRowsThreads, 1) lz4_expand_rows_kernel(lz4k_decode_args a)
{
    static_assert(kRowsThreads == 1024, "rows_resolve assumes 1024 threads (32 rows per wave slice)");
    extern __shared__ __align__(16) uint8_t smemRaw[];
    RowsSmem& S = *reinterpret_cast<RowsSmem*>(smemRaw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WsView w = ws_view(a);
    const uint32_t le = lanemask_le();
    const uint32_t outA = (uint32_t)offsetof(RowsSmem, out);
    const int zeroDelta0 = (int)offsetof(RowsSmem, zero) - (int)outA;
    const uint32_t sBase = smem_u32(smemRaw);
    const uint32_t outS = sBase + outA, zeroS = sBase + (uint32_t)offsetof(RowsSmem, zero);
    const uint32_t rowsS = sBase + (uint32_t)offsetof(RowsSmem, rows), tabS = sBase + (uint32_t)offsetof(RowsSmem, tab);
    uint32_t parity = 0, wpar = 0;

    /* a block's arguments; n = 0 marks "nothing to do here" (past the end, rejected by the scan, or a block of the
     * generic kernel: input > 65535 bytes, capacity > 64 KB, more than kMaxSeqFast sequences).  The raw loads are
     * issued a block ahead (fetchRaw) and only looked at after the waves (finish), so nobody waits for them. */
    struct Raw { int64_t b, off, dOff; int n, total, cap; uint32_t ns; };
    auto fetchRaw = [&](int64_t b) {
        Raw r;
        r.b = b; r.off = 0; r.dOff = 0; r.n = 0; r.total = 0; r.cap = 0; r.ns = 0;
        if (b < a.nBlocks) {
            r.n = a.srcSize[b]; r.total = a.outSize[b]; r.ns = w.nSeq[b]; r.off = a.srcOff[b];
            r.cap = a.dstCapArr ? a.dstCapArr[b] : a.dstCap;
            r.dOff = a.dstOff ? a.dstOff[b] : b * a.dstStride;
        }
        return r;
    };
    auto finish = [&](const Raw& r) {
        RowsDesc d;
        d.b = r.b; d.src = nullptr; d.dst = nullptr; d.n = 0; d.total = 0; d.nseq = 0;
        if (r.b < a.nBlocks && r.total > 0 && rows_eligible(r.n, r.cap, r.ns, w.markStride)) {
            d.src = a.src + r.off; d.dst = a.dst + r.dOff;
            d.n = r.n; d.total = r.total; d.nseq = (int)r.ns;
        }
        return d;
    };
    auto issueLoad = [&](const RowsDesc& d) {                     /* thread 0: TMA bulk load of the compressed block */
        const int head = (int)(reinterpret_cast<uintptr_t>(d.src) & 15);
        const uint32_t loadBytes = (uint32_t)((head + d.n + 15) & ~15);
        mbar_expect_tx(&S.mbar, loadBytes);
        for (uint32_t o = 0; o < loadBytes; o += 16384u)
            tma_load_1d(S.in + o, d.src - head + o, min(16384u, loadBytes - o), &S.mbar);
    };

    if (tid == 0) { mbar_init(&S.mbar, 1); mbar_init(&S.wbar, kRowsThreads / 32); }
    if (tid < 4) reinterpret_cast<uint32_t*>(S.zero)[tid] = 0;
    if (tid == 0) {
        S.desc[0] = finish(fetchRaw(blockIdx.x));
        if (S.desc[0].n > 0) issueLoad(S.desc[0]);
    }
    __syncthreads();
#ifdef LZ4K_PHASE_TIMING
    long long tPhase = clock64();
#endif

    /* static assignment: CTA c takes blocks c, c + grid, c + 2 grid, ... (blocks of a batch cost about the same) */
    for (uint32_t it = 0;; it++) {
        const int64_t b = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
        if (b >= a.nBlocks) break;
        const RowsDesc d = S.desc[it & 1];
        Raw rawNext;                                           /* the next block's arguments: loads issued now, used after the waves */
        if (tid == 0) rawNext = fetchRaw(b + gridDim.x);
        if (d.n == 0) {                                        /* not a block of this kernel */
            __syncthreads();
            if (tid == 0) { const RowsDesc dn = finish(rawNext); S.desc[(it + 1) & 1] = dn; if (dn.n > 0) issueLoad(dn); }
            __syncthreads();
            continue;
        }
        const int n = d.n, total = d.total, nseq = d.nseq;
        const int head = (int)(reinterpret_cast<uintptr_t>(d.src) & 15);

        for (int k = tid; k < 2048; k += kRowsThreads) S.rows[k] = make_uint2(0u, 0u);
        /* this thread's first marks, fetched while the TMA load is in flight */
        const uint32_t* marks = w.marks + d.b * w.markStride;
        uint32_t mk[kRowsCache];
        #pragma unroll
        for (int i = 0; i < kRowsCache; i++) {
            const int k = tid + i * kRowsThreads;
            mk[i] = (k < nseq) ? marks[k] : 0u;
        }
        __syncthreads();                                   /* rows are zero */
        PHASE_MARK(0);                                     // zeroing + marks
        mbar_wait(&S.mbar, parity);
        parity ^= 1;
        PHASE_MARK(1);                                     // TMA load wait

        /* ---- runs, pass 1: one bit per run start ---- */
        const uint8_t* in = S.in + head;
        const uint32_t inA = (uint32_t)offsetof(RowsSmem, in) + (uint32_t)head;
        auto setBit = [&](int s, int) { atomicOr(&S.rows[s >> 5].x, 1u << (s & 31)); };
        RwSeq sq[kRowsCache];
        #pragma unroll
        for (int i = 0; i < kRowsCache; i++) {
            const int k = tid + i * kRowsThreads;
            if (k < nseq) {
                sq[i] = rw_parse(in, mk[i], k, k + 1 == nseq);
                if (sq[i].ll > 0) setBit(sq[i].op, 0);
                if (sq[i].mlen > 0) rw_match_runs(sq[i].m, sq[i].off, sq[i].mlen, zeroDelta0, setBit);
            }
        }
        for (int k = tid + kRowsCache * kRowsThreads; k < nseq; k += kRowsThreads) {
            const RwSeq s = rw_parse(in, marks[k], k, k + 1 == nseq);
            if (s.ll > 0) setBit(s.op, 0);
            if (s.mlen > 0) rw_match_runs(s.m, s.off, s.mlen, zeroDelta0, setBit);
        }
        __syncthreads();
        PHASE_MARK(2);                                     // pass 1

        /* ---- rank: rows[r].y = (number of run starts in rows [0, r)) - 1 ---- */
        {
            constexpr int WPT = 2048 / kRowsThreads;
            uint32_t cnt[WPT], x = 0;
            #pragma unroll
            for (int j = 0; j < WPT; j++) { cnt[j] = __popc(S.rows[tid * WPT + j].x); x += cnt[j]; }
            uint32_t incl = x;
            #pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) { uint32_t y = __shfl_up_sync(kFull, incl, dd); if (lane >= dd) incl += y; }
            if (lane == 31) S.warpSum[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                uint32_t v = S.warpSum[lane];
                #pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) { uint32_t y = __shfl_up_sync(kFull, v, dd); if (lane >= dd) v += y; }
                S.warpSum[lane] = v;
                if (lane == 31) S.nRuns = v;
            }
            __syncthreads();
            uint32_t ex = incl - x + (warp ? S.warpSum[warp - 1] : 0);
            #pragma unroll
            for (int j = 0; j < WPT; j++) { S.rows[tid * WPT + j].y = ex - 1u; ex += cnt[j]; }
        }
        const uint32_t nRuns = S.nRuns;
        if (tid == 0) tma_wait_read0();                    /* the previous block's bulk store has finished reading S.out */
        __syncthreads();
        PHASE_MARK(3);                                     // rank
        const bool tooMany = nRuns > (uint32_t)kRowsMaxRuns;   /* (pathological) more runs than the table holds: generic kernel */
        if (tooMany && tid == 0) w.slowList[atomicAdd(&w.hdr->slowCount, 1u)] = (uint32_t)d.b;

        if (!tooMany) {
            /* ---- runs, pass 2: delta of every run at its rank ---- */
            auto emit = [&](const RwSeq& s) {
                if (s.ll > 0) S.tab[rw_rank(S.rows, (uint32_t)s.op)] = (inA + (uint32_t)s.ls) - (outA + (uint32_t)s.op);
                if (s.mlen > 0) {
                    uint32_t j = rw_rank(S.rows, (uint32_t)s.m);
                    rw_match_runs(s.m, s.off, s.mlen, zeroDelta0, [&](int, int dlt) { S.tab[j++] = (uint32_t)dlt; });
                }
            };
            #pragma unroll
            for (int i = 0; i < kRowsCache; i++) {
                const int k = tid + i * kRowsThreads;
                if (k < nseq) emit(sq[i]);
            }
            for (int k = tid + kRowsCache * kRowsThreads; k < nseq; k += kRowsThreads) {
                emit(rw_parse(in, marks[k], k, k + 1 == nseq));
            }
            __syncthreads();
            PHASE_MARK(4);                                     // pass 2

            /* ---- waves: resolve (run table only) | wait for the previous wave | copy | arrive ---- */
            const int nWaves = (total + kWave - 1) / kWave;
            uint32_t sa[kRowsRpt];
            auto resolve = [&](int wv) {
                const uint32_t p0 = (uint32_t)(wv * kWave + tid);
                if ((wv + 1) * kWave <= total) rows_resolve<true>(sa, p0, outS + (uint32_t)(wv * kWave), 0xFFFFFFFFu, outS, rowsS, tabS, zeroS, le);
                else rows_resolve<false>(sa, p0, outS + (uint32_t)(wv * kWave), (uint32_t)total, outS, rowsS, tabS, zeroS, le);
            };
#ifdef LZ4K_WAVE_BARSYNC
            /* debug build: the same waves with a plain CTA barrier instead of the split mbarrier (no overlap of the
             * resolve with the wait).  compute-sanitizer's racecheck does not model mbarrier arrive / wait as
             * synchronisation; this variant shows that the wave structure itself is hazard free. */
            for (int wv = 0; wv < nWaves; wv++) {
                resolve(wv);
                const uint32_t p0 = (uint32_t)(wv * kWave + tid);
                const uint32_t lim = ((wv + 1) * kWave <= total) ? 0xFFFFFFFFu : (uint32_t)total;
                uint32_t v[kRowsRpt];
                #pragma unroll
                for (int r = 0; r < kRowsRpt; r++) v[r] = lds_u8(sa[r]);
                #pragma unroll
                for (int r = 0; r < kRowsRpt; r++)
                    if (p0 + (uint32_t)(r * kRowsThreads) < lim) sts_u8(outS + p0 + (uint32_t)(r * kRowsThreads), v[r]);
                if (wv == nWaves - 1) fence_proxy_async();
                __syncthreads();
            }
#else
            resolve(0);
            for (int wv = 0; wv < nWaves; wv++) {
                if (wv > 0) { mbar_wait(&S.wbar, wpar); wpar ^= 1; }   /* every warp has copied wave wv-1 */
                const uint32_t p0 = (uint32_t)(wv * kWave + tid);
                const uint32_t lim = ((wv + 1) * kWave <= total) ? 0xFFFFFFFFu : (uint32_t)total;
                uint32_t v[kRowsRpt];
                #pragma unroll
                for (int r = 0; r < kRowsRpt; r++) v[r] = lds_u8(sa[r]);
                #pragma unroll
                for (int r = 0; r < kRowsRpt; r++)
                    if (p0 + (uint32_t)(r * kRowsThreads) < lim) sts_u8(outS + p0 + (uint32_t)(r * kRowsThreads), v[r]);
                if (wv == nWaves - 1) fence_proxy_async();     /* generic-proxy writes of `out` before the bulk store reads them */
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.wbar);           /* release: this warp's bytes of wave wv are written */
                if (wv + 1 < nWaves) resolve(wv + 1);
            }
            mbar_wait(&S.wbar, wpar);
            wpar ^= 1;
#endif
            PHASE_MARK(5);                                     // waves
        }

        /* `in`, `tab` and `rows` are dead: start the next block's load before anything else */
        if (tid == 0) {
            const RowsDesc dn = finish(rawNext);
            S.desc[(it + 1) & 1] = dn;
            if (dn.n > 0) issueLoad(dn);
        }
        /* ---- store: smem -> HBM ---- */
        if (!tooMany) {
            uint8_t* dst = d.dst;
            if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                const uint32_t bulk = (uint32_t)total & ~15u;
                if (tid == 0 && bulk) {
                    for (uint32_t o = 0; o < bulk; o += 16384u) tma_store_1d(dst + o, S.out + o, min(16384u, bulk - o));
                    tma_commit();
                }
                if (tid < (total & 15)) dst[bulk + tid] = S.out[bulk + tid];
            } else {
                for (int k = tid; k < total; k += kRowsThreads) dst[k] = S.out[k];
            }
        }
        __syncthreads();                                   /* desc[(it+1)&1] is visible; `out` tail reads are done */
        PHASE_MARK(6);                                     // store issue
    }
    if (tid == 0) tma_wait_all0();
}

/* =============================================================================================
 * expand (tiles): blocks ABOVE 64 KB (lz4frame's 256 KB .. 4 MB blocks), one CTA per 60 KB OUTPUT TILE
 *
 * The rows formulation applied to a window of a big block.  The scan (lz4_scan_par_kernel) leaves wide marks
 * {token position, match start} per sequence and lz4_tile_index_kernel the first sequence of every tile; a unit of work is
 * (block b, tile t) = output bytes [t * kTile, (t + 1) * kTile) of block b.  Differences to lz4_expand_rows_kernel:
 *   - the sequences that overlap the tile are parsed from GLOBAL memory (their tokens may lie megabytes apart from the
 *     tile's literals) and their runs are clipped to the tile (rw_tile_runs);
 *   - the compressed block is not staged: a literal run's delta maps into a virtual range [kLitBase, ...) that means
 *     "byte x - kLitBase of the compressed block", read from global memory through the read-only cache;
 *   - a match source below the tile start lies in an EARLIER tile of the same block: it is read from the destination
 *     in global memory (L2).  Tile t therefore waits, after its passes 1-2 and before its waves, for tile t - 1's flag,
 *     which that tile's CTA sets once its bulk store has completed.  Units are numbered tile-major (all blocks' tile 0,
 *     then tile 1, ...) and taken in ascending order by a grid of resident CTAs, so a tile's predecessor was started
 *     earlier: no deadlock; with >= 148 blocks in the batch the wait is over before it begins.
 *   - a tile with more runs than the table holds hands its block to the generic kernel (which rewrites the whole block).
 * ============================================================================================= */
constexpr uint32_t kLitBase = 1u << 28;                   /* virtual window addresses >= kLitBase: compressed byte (address - kLitBase) */
constexpr uint32_t kZeroV = kLitBase + (1u << 27);         /* the always-zero cell (offset 0, lz4.c:2407) */

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool FULL>
__device__ __forceinline__ void tiles_resolve(uint32_t (&sa)[LZ4K_ROWS_RPT], const uint32_t p0, const uint32_t waveS,
                                              const uint32_t lim, const uint32_t outS, const uint32_t rowsS,
                                              const uint32_t tabS, const uint32_t le)
{
    const uint32_t rowAddr = rowsS + ((p0 >> 5) << 3);
    bool any = false;
    #pragma unroll
    for (int r = 0; r < LZ4K_ROWS_RPT; r++) {
        sa[r] = outS + 65536u;                                  /* (a byte past the tile: never read, never "inside the wave") */
        if (FULL || p0 + (uint32_t)(r * 1024) < lim) {
            const uint2 row = lds_u64(rowAddr + (uint32_t)(r * 32 * 8));
            const uint32_t j = row.y + (uint32_t)__popc(row.x & le);
            sa[r] = outS + p0 + (uint32_t)(r * 1024) + lds_u32(tabS + (j << 2));
        }
        any |= (sa[r] - waveS) < (uint32_t)kWave;
    }
    while (any) {                                               /* sources inside this wave: follow them */
        any = false;
        #pragma unroll
        for (int r = 0; r < LZ4K_ROWS_RPT; r++) {
            uint32_t x = sa[r];
            if ((x - waveS) < (uint32_t)kWave) {
                const uint32_t q = x - outS;
                const uint2 row = lds_u64(rowsS + ((q >> 5) << 3));
                const uint32_t j = row.y + (uint32_t)__popc(row.x & (0xFFFFFFFFu >> (31u - (q & 31u))));
                x += lds_u32(tabS + (j << 2));
                sa[r] = x;
            }
            any |= (x - waveS) < (uint32_t)kWave;
        }
    }
}

__global__ void __launch_bounds__(kRowsThreads, 1) lz4_expand_tiles_kernel(lz4k_decode_args a)
{
    extern __shared__ __align__(16) uint8_t smemRaw[];
    RowsSmem& S = *reinterpret_cast<RowsSmem*>(smemRaw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WsView w = ws_view(a);
    const uint32_t le = lanemask_le();
    const uint32_t sBase = smem_u32(smemRaw);
    const uint32_t outS = sBase + (uint32_t)offsetof(RowsSmem, out);
    const uint32_t rowsS = sBase + (uint32_t)offsetof(RowsSmem, rows), tabS = sBase + (uint32_t)offsetof(RowsSmem, tab);
    const uint32_t tilesMax = w.tilesMax;
    uint32_t wpar = 0;
    uint32_t* pendFlag = nullptr;                               /* thread 0: flag of the unit whose bulk store is in flight */
    auto publishPending = [&]() {                               /* thread 0 */
        tma_wait_all0();                                        /* the store has completed (not only finished reading `out`) */
        if (pendFlag) { __threadfence(); st_release_gpu(pendFlag, 1u); pendFlag = nullptr; }
    };
    if (tid == 0) mbar_init(&S.wbar, kRowsThreads / 32);
    __syncthreads();

    const int64_t units = a.nBlocks * (int64_t)tilesMax;
    for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
        const int64_t b = u % a.nBlocks;
        const uint32_t t = (uint32_t)(u / a.nBlocks);
        const int total = a.outSize[b];
        const uint32_t nseq = w.nSeq[b];
        uint32_t* flags = w.tileFlag + (size_t)b * (tilesMax + 1);
        if (!(total > 0 && nseq > 0 && nseq <= w.markStride / 2u && t < tiles_of(total))) {     /* (the same for every thread) */
            if (tid == 0 && pendFlag) publishPending();
            continue;
        }
        const uint8_t* src = a.src + a.srcOff[b];
        uint8_t* dstB = a.dst + (a.dstOff ? a.dstOff[b] : b * a.dstStride);
        const uint32_t* marks = w.marks + (size_t)b * w.markStride;
        const uint32_t* first = w.tileFirst + (size_t)b * (tilesMax + 1);
        const int os = (int)(t * (uint32_t)kTile), oe = min(os + kTile, total), len = oe - os;
        const uint32_t f0 = first[t], f1 = first[t + 1];
        const int k0 = f0 ? (int)f0 - 1 : 0, k1 = (int)min(f1 + 1u, nseq);
        const uint32_t litBase = kLitBase - outS + (uint32_t)os;
        const int zeroDelta0 = (int)(kZeroV - outS + (uint32_t)os);
        auto parse = [&](int k) { return rw_parse_wide(src, marks[2 * (size_t)k], marks[2 * (size_t)k + 1], k + 1 == (int)nseq); };

        for (int k = tid; k < 2048; k += kRowsThreads) S.rows[k] = make_uint2(0u, 0u);
        __syncthreads();
        /* ---- runs, pass 1 (a thread's first sequences stay in registers for pass 2: parsing reads global memory) ---- */
        RwSeq sq[kRowsCache];
        #pragma unroll
        for (int i = 0; i < kRowsCache; i++) {
            const int k = k0 + tid + i * kRowsThreads;
            if (k < k1) {
                sq[i] = parse(k);
                rw_tile_runs(sq[i], os, oe, litBase, zeroDelta0, [&](int st, int) { atomicOr(&S.rows[st >> 5].x, 1u << (st & 31)); });
            }
        }
        for (int k = k0 + tid + kRowsCache * kRowsThreads; k < k1; k += kRowsThreads) {
            const RwSeq s = parse(k);
            rw_tile_runs(s, os, oe, litBase, zeroDelta0, [&](int st, int) { atomicOr(&S.rows[st >> 5].x, 1u << (st & 31)); });
        }
        __syncthreads();
        /* ---- rank ---- */
        {
            constexpr int WPT = 2048 / kRowsThreads;
            uint32_t cnt[WPT], x = 0;
            #pragma unroll
            for (int j = 0; j < WPT; j++) { cnt[j] = __popc(S.rows[tid * WPT + j].x); x += cnt[j]; }
            uint32_t incl = x;
            #pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) { uint32_t y = __shfl_up_sync(kFull, incl, dd); if (lane >= dd) incl += y; }
            if (lane == 31) S.warpSum[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                uint32_t v = S.warpSum[lane];
                #pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) { uint32_t y = __shfl_up_sync(kFull, v, dd); if (lane >= dd) v += y; }
                S.warpSum[lane] = v;
                if (lane == 31) S.nRuns = v;
            }
            __syncthreads();
            uint32_t ex = incl - x + (warp ? S.warpSum[warp - 1] : 0);
            #pragma unroll
            for (int j = 0; j < WPT; j++) { S.rows[tid * WPT + j].y = ex - 1u; ex += cnt[j]; }
        }
        const uint32_t nRuns = S.nRuns;
        if (tid == 0) publishPending();                        /* the previous unit's store is complete: its flag goes up, `out` is free */
        __syncthreads();
        if (nRuns > (uint32_t)kRowsMaxRuns) {                  /* (pathological) the generic kernel redoes the whole block */
            if (tid == 0) {
                if (atomicExch(&flags[tilesMax], 1u) == 0u) w.slowList[atomicAdd(&w.hdr->slowCount, 1u)] = (uint32_t)b;
                __threadfence();
                st_release_gpu(&flags[t], 1u);                 /* whoever waits for this tile may go on: its bytes will be rewritten */
            }
            __syncthreads();
            continue;
        }
        /* ---- runs, pass 2 ---- */
        #pragma unroll
        for (int i = 0; i < kRowsCache; i++) {
            const int k = k0 + tid + i * kRowsThreads;
            if (k < k1) rw_tile_runs(sq[i], os, oe, litBase, zeroDelta0, [&](int st, int d) { S.tab[rw_rank(S.rows, (uint32_t)st)] = (uint32_t)d; });
        }
        for (int k = k0 + tid + kRowsCache * kRowsThreads; k < k1; k += kRowsThreads) {
            const RwSeq s = parse(k);
            rw_tile_runs(s, os, oe, litBase, zeroDelta0, [&](int st, int d) { S.tab[rw_rank(S.rows, (uint32_t)st)] = (uint32_t)d; });
        }
        __syncthreads();
        /* ---- the tile before this one must be in global memory ---- */
        if (t > 0) {
            if (tid == 0) { while (ld_acquire_gpu(&flags[t - 1]) == 0u) __nanosleep(200); }
            __syncthreads();
        }
        /* ---- waves ---- */
        {
            const int nWaves = (len + kWave - 1) / kWave;
            uint32_t sa[kRowsRpt];
            auto resolve = [&](int wv) {
                const uint32_t p0 = (uint32_t)(wv * kWave + tid);
                if ((wv + 1) * kWave <= len) tiles_resolve<true>(sa, p0, outS + (uint32_t)(wv * kWave), 0xFFFFFFFFu, outS, rowsS, tabS, le);
                else tiles_resolve<false>(sa, p0, outS + (uint32_t)(wv * kWave), (uint32_t)len, outS, rowsS, tabS, le);
            };
#ifdef LZ4K_WAVE_BARSYNC
            /* debug build (see lz4_expand_rows_kernel): a plain CTA barrier between the waves, for compute-sanitizer's racecheck */
            for (int wv = 0; wv < nWaves; wv++) {
                resolve(wv);
                const uint32_t p0 = (uint32_t)(wv * kWave + tid);
                const uint32_t lim = ((wv + 1) * kWave <= len) ? 0xFFFFFFFFu : (uint32_t)len;
                uint32_t v[kRowsRpt];
                #pragma unroll
                for (int r = 0; r < kRowsRpt; r++) {
                    const uint32_t x = sa[r];
                    v[r] = 0u;
                    if (p0 + (uint32_t)(r * kRowsThreads) < lim) {
                        if (x >= kLitBase) v[r] = (x == kZeroV) ? 0u : (uint32_t)__ldg(src + (x - kLitBase));
                        else if (x >= outS) v[r] = lds_u8(x);
                        else v[r] = (uint32_t)__ldcg(dstB + (os - (int)(outS - x)));
                    }
                }
                #pragma unroll
                for (int r = 0; r < kRowsRpt; r++)
                    if (p0 + (uint32_t)(r * kRowsThreads) < lim) sts_u8(outS + p0 + (uint32_t)(r * kRowsThreads), v[r]);
                if (wv == nWaves - 1) fence_proxy_async();
                __syncthreads();
            }
#else
            resolve(0);
            for (int wv = 0; wv < nWaves; wv++) {
                if (wv > 0) { mbar_wait(&S.wbar, wpar); wpar ^= 1; }   /* every warp has copied wave wv-1 */
                const uint32_t p0 = (uint32_t)(wv * kWave + tid);
                const uint32_t lim = ((wv + 1) * kWave <= len) ? 0xFFFFFFFFu : (uint32_t)len;
                uint32_t v[kRowsRpt];
                #pragma unroll
                for (int r = 0; r < kRowsRpt; r++) {
                    const uint32_t x = sa[r];
                    v[r] = 0u;
                    if (p0 + (uint32_t)(r * kRowsThreads) < lim) {
                        if (x >= kLitBase) v[r] = (x == kZeroV) ? 0u : (uint32_t)__ldg(src + (x - kLitBase));
                        else if (x >= outS) v[r] = lds_u8(x);
                        else v[r] = (uint32_t)__ldcg(dstB + (os - (int)(outS - x)));
                    }
                }
                #pragma unroll
                for (int r = 0; r < kRowsRpt; r++)
                    if (p0 + (uint32_t)(r * kRowsThreads) < lim) sts_u8(outS + p0 + (uint32_t)(r * kRowsThreads), v[r]);
                if (wv == nWaves - 1) fence_proxy_async();     /* generic-proxy writes of `out` before the bulk store reads them */
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.wbar);
                if (wv + 1 < nWaves) resolve(wv + 1);
            }
            mbar_wait(&S.wbar, wpar);
            wpar ^= 1;
#endif
        }
        /* ---- store: smem -> HBM ---- */
        {
            uint8_t* dst = dstB + os;
            if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                const uint32_t bulk = (uint32_t)len & ~15u;
                if (tid == 0 && bulk) {
                    for (uint32_t o = 0; o < bulk; o += 16384u) tma_store_1d(dst + o, S.out + o, min(16384u, bulk - o));
                    tma_commit();
                }
                if (tid < (len & 15)) { dst[bulk + tid] = S.out[bulk + tid]; __threadfence(); }
            } else {
                for (int k = tid; k < len; k += kRowsThreads) dst[k] = S.out[k];
                __threadfence();
            }
            if (tid == 0) pendFlag = &flags[t];
        }
        __syncthreads();                                       /* `out` tail reads are done; direct stores are fenced */
    }
    if (tid == 0) publishPending();
}

/* ---- ceiling of the rows kernel's skeleton (developer tool; bench.py --ceiling) ----
 * Same persistent structure -- one CTA of 1024 threads per SM, TMA bulk load of the compressed block into `in`, TMA bulk
 * store of 64 KB from `out` -- with the decode replaced by
 *   mode 0: nothing (load b+1 and store b are both in flight): what HBM and the TMA path give this structure;
 *   mode 1: one LDS.U8 + STS.U8 per output byte, one byte per thread, 4096-byte waves with a CTA barrier each: the floor
 *           of moving every byte through shared memory at byte granularity, without any address resolution.
 * The bytes written are meaningless; the timing is the point. */
__global__ void __launch_bounds__(kRowsThreads, 1) lz4_ceiling_kernel(lz4k_decode_args a, int mode)
{
    extern __shared__ __align__(16) uint8_t smemRaw[];
    RowsSmem& S = *reinterpret_cast<RowsSmem*>(smemRaw);
    const int tid = threadIdx.x;
    const uint32_t sBase = smem_u32(smemRaw);
    const uint32_t outS = sBase + (uint32_t)offsetof(RowsSmem, out), inS = sBase + (uint32_t)offsetof(RowsSmem, in);
    uint32_t parity = 0;
    auto issueLoad = [&](int64_t b) {
        const uint8_t* src = a.src + a.srcOff[b];
        const int n = a.srcSize[b];
        const int head = (int)(reinterpret_cast<uintptr_t>(src) & 15);
        const uint32_t loadBytes = (uint32_t)((head + n + 15) & ~15);
        mbar_expect_tx(&S.mbar, loadBytes);
        for (uint32_t o = 0; o < loadBytes; o += 16384u) tma_load_1d(S.in + o, src - head + o, min(16384u, loadBytes - o), &S.mbar);
    };
    if (tid == 0) { mbar_init(&S.mbar, 1); }
    __syncthreads();
    if (tid == 0 && blockIdx.x < a.nBlocks) issueLoad(blockIdx.x);
    for (int64_t b = blockIdx.x; b < a.nBlocks; b += gridDim.x) {
        mbar_wait(&S.mbar, parity);
        parity ^= 1;
        if (mode >= 1) {
            if (tid == 0) tma_wait_read0();                    /* the previous store has finished reading `out` */
            __syncthreads();
            for (int wv = 0; wv < 65536 / kWave; wv++) {
                uint32_t v[kRowsRpt];
                #pragma unroll
                for (int r = 0; r < kRowsRpt; r++) v[r] = lds_u8(inS + (uint32_t)((wv * kWave + r * kRowsThreads + tid) & 0x7FFF));
                #pragma unroll
                for (int r = 0; r < kRowsRpt; r++) sts_u8(outS + (uint32_t)(wv * kWave + r * kRowsThreads + tid), v[r]);
                __syncthreads();
            }
            fence_proxy_async();
            __syncthreads();
        }
        if (tid == 0) {
            if (b + gridDim.x < a.nBlocks) issueLoad(b + gridDim.x);
            uint8_t* dst = a.dst + b * a.dstStride;
            for (uint32_t o = 0; o < 65536u; o += 16384u) tma_store_1d(dst + o, S.out + o, 16384u);
            tma_commit();
        }
        if (mode >= 1) __syncthreads();
    }
    if (tid == 0) tma_wait_all0();
}

/* =============================================================================================
 * encode: one warp per block, byte-identical replay of LZ4_compress_generic_validated
 * ============================================================================================= */

template <bool SMALL> struct Table;
template <> struct Table<true> {       // byU16, 13-bit 4-byte hash (lz4.c:779-780)
    uint16_t* t;
    __device__ __forceinline__ uint32_t hash(const uint8_t* p) const { return (ld32u<true>(p) * 2654435761u) >> 19; }
    __device__ __forceinline__ uint32_t hashv(const uint8_t* p, uint32_t& v32) const { v32 = ld32u<true>(p); return (v32 * 2654435761u) >> 19; }
    __device__ __forceinline__ uint32_t get(uint32_t h) const { return t[h]; }
    __device__ __forceinline__ void put(uint32_t h, uint32_t pos) const { t[h] = (uint16_t)pos; }
};
template <> struct Table<false> {      // byU32, 12-bit 5-byte hash (lz4.c:785-791,799)
    uint32_t* t;
    __device__ __forceinline__ static uint32_t h5(uint64_t v) { return (uint32_t)(((v << 24) * 889523592379ull) >> 52); }
    __device__ __forceinline__ uint32_t hash(const uint8_t* p) const { return h5(ld40u(p)); }
    __device__ __forceinline__ uint32_t hashv(const uint8_t* p, uint32_t& v32) const { uint64_t v = ld40u(p); v32 = (uint32_t)v; return h5(v); }
    __device__ __forceinline__ uint32_t get(uint32_t h) const { return t[h]; }
    __device__ __forceinline__ void put(uint32_t h, uint32_t pos) const { t[h] = pos; }
};

/* sum_{t<m} (accel + (t >> 6)): distance covered by m search steps after the first one,
 * from step = (searchMatchNb++ >> LZ4_skipTrigger) with searchMatchNb = accel << 6 (lz4.c:1044-1053) */
__device__ __forceinline__ uint32_t skip_distance(uint32_t m, uint32_t accel)
{
    uint32_t q = m >> kSkipTrigger, r = m & 63u;
    return accel * m + 32u * q * (q - 1u) + q * r;
}

/* run-length extension bytes: `len` -> 255,255,...,rem  (lz4.c:1123-1128) written by the warp */
__device__ __forceinline__ uint32_t emit_runlength(uint8_t* dst, uint32_t op, uint32_t len, int lane)
{
    uint32_t full = len / 255u;
    for (uint32_t k = lane; k < full; k += 32) dst[op + k] = 255;
    if (lane == 0) dst[op + full] = (uint8_t)(len - full * 255u);
    return op + full + 1;
}

template <bool SMALL>
__device__ int encode_block(const uint8_t* __restrict__ src, const int n, uint8_t* __restrict__ dst,
                            const int dstCap, const uint32_t accel, void* tableMem, const int lane)
{
    Table<SMALL> T;
    T.t = reinterpret_cast<decltype(T.t)>(tableMem);
    const int bound = n + n / 255 + 16;
    const bool limited = !(dstCap >= bound);
    const int64_t olimit = dstCap;
    const uint32_t mflimit1 = (uint32_t)n - kMfLimit + 1;     // lz4.c:963
    const uint32_t matchlimit = (uint32_t)n - kLastLiterals;  // lz4.c:964
    uint32_t ip = 0, anchor = 0, op = 0, cand = 0;

    {   // zeroed table, lz4.c:1558
        uint4* z = reinterpret_cast<uint4*>(tableMem);
        for (int k = lane; k < 1024; k += 32) z[k] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();

    if (n < kMinLength) goto tail;                            // lz4.c:1002

    if (lane == 0) T.put(T.hash(src), 0);                     // lz4.c:1005-1010
    __syncwarp();
    ip = 1;

    for (;;) {
        /* ---- search: lanes probe the next 32 positions of the reference's visiting order ---- */
        {
            uint32_t baseJ = 0;
            bool found = false;
            for (;;) {
                const uint32_t j = baseJ + lane;
                const uint32_t p = ip + (j ? 1u + skip_distance(j - 1, accel) : 0u);
                const uint32_t pnext = ip + 1u + skip_distance(j, accel);
                const bool term = pnext > mflimit1;           // lz4.c:1055 fires at this visit
                uint32_t v32 = 0, h = 0x80000000u | lane, old = 0;
                if (!term) { h = T.hashv(src + p, v32); old = T.get(h); }
                __syncwarp();                                 // every lane has read the table before any lane writes it (below)
                const unsigned peers = __match_any_sync(kFull, h);
                const unsigned lower = peers & ((1u << lane) - 1u);
                const int fromLane = lower ? (31 - __clz(lower)) : lane;
                const uint32_t pPrev = __shfl_sync(kFull, p, fromLane);
                const uint32_t c = lower ? pPrev : old;       // table value this visit would read
                bool hit = false;
                if (!term) {
                    if (SMALL || c + kMaxDistance >= p) hit = (ld32u<true>(src + c) == v32);   // lz4.c:1090-1096
                }
                const unsigned termMask = __ballot_sync(kFull, term);
                const unsigned hitMask = __ballot_sync(kFull, hit);
                const int firstTerm = termMask ? (__ffs(termMask) - 1) : 32;
                const int firstHit = hitMask ? (__ffs(hitMask) - 1) : 32;
                if (firstHit < firstTerm) {
                    /* visits 0..firstHit happened: each stored its position (lz4.c:1085); for equal
                     * hashes the latest visit wins */
                    const unsigned upto = (firstHit == 31) ? kFull : ((2u << firstHit) - 1u);
                    const unsigned mine = peers & upto;
                    if (lane <= firstHit && (mine >> lane) == 1u) T.put(h, p);
                    ip = __shfl_sync(kFull, p, firstHit);
                    cand = __shfl_sync(kFull, c, firstHit);
                    found = true;
                    __syncwarp();
                    break;
                }
                if (firstTerm < 32) break;                    // goto _last_literals
                if ((peers >> lane) == 1u) T.put(h, p);
                __syncwarp();
                baseJ += 32;
            }
            if (!found) goto tail;
        }

        /* ---- backward extension (lz4.c:1107-1109) ---- */
        for (;;) {
            const uint32_t room = min(ip - anchor, cand);
            const uint32_t k = lane + 1;
            const bool eq = (k <= room) && (ldb<true>(src + ip - k) == ldb<true>(src + cand - k));
            const unsigned m = __ballot_sync(kFull, eq);
            const uint32_t run = (m == kFull) ? 32u : (uint32_t)(__ffs(~m) - 1);
            ip -= run; cand -= run;
            if (run < 32) break;
        }

        {
            uint32_t lit = ip - anchor;
            bool haveLiterals = true;
            for (;;) {   /* the _next_match chain, lz4.c:1138-1294 */
                /* match length first (the token needs it): LZ4_count, lz4.c:1182 */
                uint32_t mcode = 0;
                {
                    const uint8_t* pa = src + ip + kMinMatch;
                    const uint8_t* pb = src + cand + kMinMatch;
                    const uint32_t lim = matchlimit - (ip + kMinMatch);    // bytes comparable
                    for (uint32_t base = 0;; base += 32) {
                        const uint32_t k = base + lane;
                        const bool eq = (k < lim) && (ldb<true>(pa + k) == ldb<true>(pb + k));
                        const unsigned m = __ballot_sync(kFull, eq);
                        if (m != kFull) { mcode = base + (uint32_t)(__ffs(~m) - 1); break; }
                    }
                }
                const uint32_t tokenPos = op;
                uint32_t o = op + 1;
                if (haveLiterals) {
                    /* lz4.c:1114-1117 */
                    if (limited && (int64_t)o + lit + (2 + 1 + kLastLiterals) + lit / 255 > olimit) return 0;
                    if (lit >= 15) o = emit_runlength(dst, o, lit - 15, lane);
                    for (uint32_t k = lane; k < lit; k += 32) dst[o + k] = (uint8_t)ldb<true>(src + anchor + k);
                    o += lit;
                }
                /* offset, lz4.c:1162 */
                const uint32_t off = ip - cand;
                if (lane == 0) { dst[o] = (uint8_t)off; dst[o + 1] = (uint8_t)(off >> 8); }
                o += 2;
                /* lz4.c:1187-1211 */
                if (limited && (int64_t)o + (1 + kLastLiterals) + (mcode + 240) / 255 > olimit) return 0;
                if (lane == 0) {
                    const uint32_t lt = haveLiterals ? min(lit, 15u) : 0u;
                    dst[tokenPos] = (uint8_t)((lt << 4) | min(mcode, 15u));
                }
                if (mcode >= 15) o = emit_runlength(dst, o, mcode - 15, lane);   // lz4.c:1213-1223
                op = o;

                ip += mcode + kMinMatch;
                anchor = ip;
                if (ip >= mflimit1) goto tail;                         // lz4.c:1233

                if (lane == 0) T.put(T.hash(src + ip - 2), ip - 2);    // lz4.c:1236-1242
                __syncwarp();
                {   /* immediate re-test at ip, lz4.c:1255-1294 */
                    uint32_t v32;
                    const uint32_t h = T.hashv(src + ip, v32);
                    cand = T.get(h);
                    __syncwarp();
                    if (lane == 0) T.put(h, ip);
                    __syncwarp();
                    if ((SMALL || cand + kMaxDistance >= ip) && ld32u<true>(src + cand) == v32) {
                        haveLiterals = false; lit = 0;
                        continue;
                    }
                }
                break;
            }
        }
        ip++;                                                          // lz4.c:1298
    }

tail:   /* lz4.c:1302-1329 */
    {
        const uint32_t last = (uint32_t)n - anchor;
        if (limited && (int64_t)op + last + 1 + (last + 255 - 15) / 255 > olimit) return 0;
        uint32_t o = op + 1;
        if (lane == 0) dst[op] = (uint8_t)(min(last, 15u) << 4);
        if (last >= 15) o = emit_runlength(dst, o, last - 15, lane);
        for (uint32_t k = lane; k < last; k += 32) dst[o + k] = (uint8_t)ldb<true>(src + anchor + k);
        return (int)(o + last);
    }
}

constexpr int kEncodeTableBytes = 16384;    // LZ4_HASHTABLESIZE, lz4.h:157-172 (LZ4_MEMORY_USAGE 14)

__global__ void __launch_bounds__(32) lz4_encode_kernel(lz4k_encode_args a)
{
    extern __shared__ uint4 tableMem[];
    const int lane = threadIdx.x;
    uint32_t accel = a.acceleration < 1 ? 1u : (a.acceleration > kAccelMax ? (uint32_t)kAccelMax : (uint32_t)a.acceleration);  // lz4.c:1386-1387
    for (int64_t b = blockIdx.x; b < a.nBlocks; b += gridDim.x) {
        const uint8_t* src = a.src + b * a.srcStride;
        uint8_t* dst = a.dst + b * a.dstStride;
        const int n = a.srcSizeArr ? a.srcSizeArr[b] : a.srcSize;
        int r;
        if ((uint32_t)n > kMaxInput) {                                  // lz4.c:1360
            r = 0;
        } else if (n == 0) {                                            // lz4.c:1361-1371
            const bool limited = !(a.dstCap >= 16);
            if (limited && a.dstCap <= 0) r = 0;
            else { if (lane == 0) dst[0] = 0; r = 1; }
        } else if (n < kSmallLimit) {                                   // lz4.c:1389
            r = encode_block<true>(src, n, dst, a.dstCap, accel, tableMem, lane);
        } else {
            r = encode_block<false>(src, n, dst, a.dstCap, accel, tableMem, lane);
        }
        if (lane == 0) a.outSize[b] = r;
        __syncwarp();
    }
}

#include "lz4_encode_par.cuh"

/* =============================================================================================
 * pack: exclusive scan of sizes (+ optional 4-byte headers) and gather into a contiguous stream
 * With a FrameRule the blocks become the body of an LZ4 frame (LZ4F_makeBlock, lz4frame.c:883-908): a block whose
 * compression did not gain (size 0 = did not fit size-1, or >= its source size) is stored raw -- payload = the
 * source bytes, bit 31 of the LE32 block header set.
 * ============================================================================================= */
struct FrameRule { const uint8_t* raw; int64_t rawStride; int32_t blockSize, lastSize; };     /* raw == nullptr: plain pack */

__device__ __forceinline__ int pack_payload(const FrameRule& fr, const int32_t* sizes, int64_t b, int64_t n, bool& isRaw)
{
    const int c = sizes[b];
    isRaw = false;
    if (!fr.raw) return c > 0 ? c : 0;
    const int sz = (b == n - 1 && fr.lastSize > 0) ? fr.lastSize : fr.blockSize;
    isRaw = (c <= 0 || c >= sz);                                /* lz4frame.c:896-899 */
    return isRaw ? sz : c;
}

__global__ void __launch_bounds__(1024) lz4_pack_scan_kernel(const int32_t* __restrict__ sizes, int64_t n,
                                                             int64_t* __restrict__ outOff, int headerBytes, FrameRule fr)
{
    __shared__ int64_t warpSums[32];
    __shared__ int64_t carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        int64_t i = base + threadIdx.x;
        int64_t v = 0;
        if (i < n) { bool isRaw; v = pack_payload(fr, sizes, i, n, isRaw) + headerBytes; }
        int64_t x = v;
        for (int d = 1; d < 32; d <<= 1) { int64_t y = __shfl_up_sync(kFull, x, d); if (lane >= d) x += y; }
        if (lane == 31) warpSums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int64_t w = warpSums[lane];
            for (int d = 1; d < 32; d <<= 1) { int64_t y = __shfl_up_sync(kFull, w, d); if (lane >= d) w += y; }
            warpSums[lane] = w;
        }
        __syncthreads();
        int64_t prefix = carry + (wid ? warpSums[wid - 1] : 0) + x - v;
        if (i < n) outOff[i] = prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) outOff[n] = carry;
}

__global__ void __launch_bounds__(256) lz4_pack_gather_kernel(const uint8_t* __restrict__ slots, int64_t slotStride,
                                                              const int32_t* __restrict__ sizes, int64_t n,
                                                              uint8_t* __restrict__ packed, const int64_t* __restrict__ outOff,
                                                              int headerBytes, FrameRule fr)
{
    for (int64_t b = blockIdx.x; b < n; b += gridDim.x) {
        bool isRaw;
        const int s = pack_payload(fr, sizes, b, n, isRaw);
        const uint8_t* from = isRaw ? fr.raw + b * fr.rawStride : slots + b * slotStride;
        uint8_t* to = packed + outOff[b];
        const uint32_t hw = (uint32_t)s | (isRaw ? 0x80000000u : 0u);
        if (headerBytes == 4 && threadIdx.x < 4) to[threadIdx.x] = (uint8_t)(hw >> (8 * threadIdx.x));   // lz4frame.c:896-907 LE32
        to += headerBytes;
        /* destination-aligned 16-byte stores, source read through aligned words */
        const uintptr_t ta = reinterpret_cast<uintptr_t>(to);
        int head = (int)((16 - (ta & 15)) & 15);
        if (head > s) head = s;
        for (int k = threadIdx.x; k < head; k += blockDim.x) to[k] = from[k];
        const int body = (s - head) >> 4;
        const uint8_t* fb = from + head;
        uint4* tb = reinterpret_cast<uint4*>(to + head);
        const uintptr_t fa = reinterpret_cast<uintptr_t>(fb);
        const uint32_t sh = (uint32_t)(fa & 3) * 8;
        const uint32_t* fw = reinterpret_cast<const uint32_t*>(fa & ~uintptr_t(3));
        for (int k = threadIdx.x; k < body; k += blockDim.x) {
            const uint32_t* w = fw + 4 * k;
            uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
            uint32_t w4 = sh ? __ldg(w + 4) : 0u;
            uint4 v;
            v.x = __funnelshift_r(w0, w1, sh); v.y = __funnelshift_r(w1, w2, sh);
            v.z = __funnelshift_r(w2, w3, sh); v.w = __funnelshift_r(w3, w4, sh);
            tb[k] = v;
        }
        for (int k = head + (body << 4) + threadIdx.x; k < s; k += blockDim.x) to[k] = from[k];
    }
}

}  // namespace

/* =============================================================================================
 * extern "C" launchers
 * ============================================================================================= */
extern "C" {

uint64_t lz4k_launch_count(void) { return g_launches; }

/* debug: read and reset the phase-cycle counters (all zero unless built with -DLZ4K_PHASE_TIMING) */
int lz4k_debug_phase_cycles(unsigned long long* out8)   /* out8: 12 values (8 phase cycles + 4 loop statistics) */
{
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaError_t e = cudaMemcpyFromSymbol(out8, g_phaseCycles, sizeof(z));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_phaseCycles, z, sizeof(z));
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out8 + 8, g_loopStats, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_loopStats, z, 4 * sizeof(unsigned long long));
    return (int)e;
}

static size_t ws_bytes_for(int64_t nBlocks, const int32_t* dstCapArr, int32_t dstCap)
{
    const bool wide = wide_batch(dstCapArr, dstCap);           /* the size that enables the tiles kernel */
    return ws_bytes(nBlocks, mark_stride(dstCapArr, dstCap, wide), wide ? tiles_of(dstCap) : 0u);
}

size_t lz4k_decode_workspace_bytes(int64_t nBlocks)               /* any capacities (worst case: 32 KB of marks per block) */
{
    return nBlocks < 0 ? 0 : ws_bytes(nBlocks, (uint32_t)kMaxSeqFast);
}

size_t lz4k_decode_workspace_bytes_for(int64_t nBlocks, int perBlockCaps, int32_t dstCap)
{
    static const int32_t one = 1;                                 /* any non-NULL pointer: "capacities are per block" */
    return nBlocks < 0 ? 0 : ws_bytes_for(nBlocks, perBlockCaps ? &one : nullptr, dstCap);
}

size_t lz4k_decode_workspace_bytes_min(int64_t nBlocks, int perBlockCaps, int32_t dstCap)   /* smallest workspace a launch accepts */
{
    static const int32_t one = 1;
    return nBlocks < 0 ? 0 : ws_bytes(nBlocks, mark_stride(perBlockCaps ? &one : nullptr, dstCap));
}

int lz4k_launch_decode(const lz4k_decode_args* a, int phases, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (a->nBlocks == 0) return 0;
    if (a->workspaceBytes < ws_bytes(a->nBlocks, mark_stride(a->dstCapArr, a->dstCap))) return (int)cudaErrorInvalidValue;
    const bool wide = use_wide(a->dstCapArr, a->dstCap, a->nBlocks, a->workspaceBytes);
    static int sms = 0, scanImpl = -1;
    if (sms == 0) {
        int dev = 0, v = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        /* opt in to the large dynamic shared memory (once; every device of a process runs the same kernels) */
        cudaError_t e = cudaFuncSetAttribute(lz4_expand_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RowsSmem));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(lz4_scan_par_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanParSmem));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(lz4_expand_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RowsSmem));
        if (e != cudaSuccess) return (int)e;
        const char* env = getenv("LZ4K_SCAN_IMPL");               /* developer A/B switch: "thread" | "par" | "split" */
        scanImpl = env ? (env[0] == 'p' ? 1 : env[0] == 's' ? 2 : 0) : -1;
        /* LZ4B200_RESERVE_SMS = k: the persistent expand kernels leave k SMs alone, for a communication kernel that runs
         * concurrently (N > 1: NCCL's send/recv of the previous chunk; bench.py --reserve-sms) */
        const char* rs = getenv("LZ4B200_RESERVE_SMS");
        const int reserve = rs ? atoi(rs) : 0;
        sms = (reserve > 0 && reserve < v) ? v - reserve : v;
    }
    if (phases & 1) {
        cudaError_t e = cudaMemsetAsync(a->workspace, 0, 256, s);     // WsHeader: list counter
        if (e != cudaSuccess) return (int)e;
        /* measured (profiles/): the lanes of the parallel scan re-walk their segments several times, ~10x the instructions of the
         * one-thread scan, so it only pays for blocks far beyond 64 KB (lz4frame's 4 MB blocks: 150 000 dependent steps for one thread) */
        const bool par = scanImpl >= 0 ? scanImpl == 1 : (a->dstCapArr == nullptr && a->dstCap > 65536);
        /* blocks of 16..64 KB: kSsLanes merging lanes per block (smaller blocks have too few sequences to split).  Measured
         * (profiles/README.md): faster than one thread per block while the batch is small (1.18 against 2.30 ms for
         * 8192 blocks, 1.56 against 2.31 ms for 16384), slower beyond (3.07 against 2.37 ms for 32768, 7.4 against 2.7 ms
         * for 65536: the one-thread scan's time hardly grows with the batch, the split scan's does) */
        const bool split = scanImpl >= 0 ? scanImpl == 2
                                         : (a->dstCapArr == nullptr && a->dstCap >= 16384 && a->dstCap <= 65536 && a->nBlocks <= kSplitMaxBlocks);
        if (split && !par) {
            const int64_t grid = (a->nBlocks + kSplitBlocksPerCta - 1) / kSplitBlocksPerCta;
            lz4_scan_split_kernel<<<(unsigned)grid, kSplitThreads, 0, s>>>(*a);
        } else if (par) {
            const int64_t want = (int64_t)sms * 12;                   // 3 resident CTAs per SM, 4 rounds for balance
            const int64_t grid = a->nBlocks < want ? a->nBlocks : want;
            lz4_scan_par_kernel<<<(unsigned)grid, kScanLanes, sizeof(ScanParSmem), s>>>(*a);
            if (wide) {                                                // tiles kernel: first sequence of every tile, flags cleared
                lz4_tile_index_kernel<<<(unsigned)(a->nBlocks < (int64_t)sms * 8 ? a->nBlocks : (int64_t)sms * 8), 256, 0, s>>>(*a);
                g_launches++;
            }
        } else {
            const int threads = 128;
            const int64_t grid = (a->nBlocks + threads - 1) / threads;
            lz4_scan_kernel<<<(unsigned)grid, threads, 0, s>>>(*a);
        }
        g_launches++;
    }
    if (phases & 2) {
        if (wide) {                                                  // blocks above 64 KB: one CTA per 60 KB output tile
            const int64_t units = a->nBlocks * (int64_t)tiles_of(a->dstCap);
            const int64_t grid = units < sms ? units : sms;          // all CTAs resident: a tile may wait for its predecessor
            {   /* the tiles' done flags start at 0 for THIS expand (the index kernel clears them too; an expand repeated after one scan must not see the last one's) */
                const WsView wv = ws_view(a->workspace, a->nBlocks, mark_stride(a->dstCapArr, a->dstCap, true), tiles_of(a->dstCap));
                cudaError_t e = cudaMemsetAsync(wv.tileFlag, 0, (size_t)a->nBlocks * (wv.tilesMax + 1) * sizeof(uint32_t), s);
                if (e != cudaSuccess) return (int)e;
            }
            lz4_expand_tiles_kernel<<<(unsigned)grid, kRowsThreads, sizeof(RowsSmem), s>>>(*a);
        } else {
            int64_t grid = a->nBlocks < sms ? a->nBlocks : sms;      // persistent: one CTA per SM
            lz4_expand_rows_kernel<<<(unsigned)grid, kRowsThreads, sizeof(RowsSmem), s>>>(*a);
        }
        g_launches++;
        const int threads = 128;                                     // 4 warps = 4 blocks per CTA, grid-stride over the slow list
        int64_t grid2 = (a->nBlocks * 32 + threads - 1) / threads;
        if (grid2 > (int64_t)sms * 16) grid2 = (int64_t)sms * 16;
        lz4_expand_generic_kernel<<<(unsigned)grid2, threads, 0, s>>>(*a);
        g_launches++;
    }
    return (int)cudaGetLastError();
}

int lz4k_launch_encode(const lz4k_encode_args* a, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (a->nBlocks == 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int perSm = 13;    // 16 KB table per warp: 13 x 16 KB + reserved fits the 228 KB SM
    int64_t grid = (int64_t)sms * perSm;
    if (grid > a->nBlocks) grid = a->nBlocks;
    lz4_encode_kernel<<<(unsigned)grid, 32, kEncodeTableBytes, s>>>(*a);
    g_launches++;
    return (int)cudaGetLastError();
}

/* developer tool: fill bytes [lo, hi) of every CTA's dynamic shared memory (parallel compressor layout) with a pattern, so
 * that a later launch that read shared memory it never wrote would show it (tests/perf/enc_determinism.py) */
__global__ void __launch_bounds__(kEpThreads, 2) lz4_poison_smem_kernel(uint32_t pattern, int lo, int hi)
{
    extern __shared__ __align__(16) uint8_t smemRaw[];
    for (int i = lo + (int)threadIdx.x; i < hi; i += kEpThreads) smemRaw[i] = (uint8_t)(pattern >> ((i & 3) * 8));
    __syncthreads();
    if (smemRaw[lo] == 1 && pattern == 0x12345678u && hi == -1) smemRaw[0] = 0;   /* keep the stores */
}
int lz4k_debug_poison_smem(uint32_t pattern, int lo, int hi, void* stream)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(lz4_poison_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncParSmem));
    if (e != cudaSuccess) return (int)e;
    if (hi > (int)sizeof(EncParSmem)) hi = (int)sizeof(EncParSmem);
    lz4_poison_smem_kernel<<<2 * sms, kEpThreads, sizeof(EncParSmem), (cudaStream_t)stream>>>(pattern, lo, hi);
    return (int)cudaGetLastError();
}

int lz4k_launch_ceiling(const lz4k_decode_args* a, int mode, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (a->nBlocks == 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(lz4_ceiling_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RowsSmem));
    if (e != cudaSuccess) return (int)e;
    const int64_t grid = a->nBlocks < sms ? a->nBlocks : sms;
    lz4_ceiling_kernel<<<(unsigned)grid, kRowsThreads, sizeof(RowsSmem), s>>>(*a, mode);
    return (int)cudaGetLastError();
}

int lz4k_launch_encode_par(const lz4k_encode_args* a, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (a->nBlocks == 0) return 0;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0, v = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = cudaFuncSetAttribute(lz4_encode_par_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncParSmem));
        if (e != cudaSuccess) return (int)e;
        sms = v;
    }
    const int64_t grid = a->nBlocks < 2 * sms ? a->nBlocks : 2 * sms;        // persistent: two CTAs per SM, blocks c, c + grid, ...
    lz4_encode_par_kernel<<<(unsigned)grid, kEpThreads, sizeof(EncParSmem), s>>>(*a);
    g_launches++;
    return (int)cudaGetLastError();
}

int lz4k_launch_pack(const uint8_t* slots, int64_t slotStride, const int32_t* sizes, int64_t nBlocks,
                     uint8_t* packed, int64_t* outOff, int headerBytes, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    FrameRule fr; fr.raw = nullptr; fr.rawStride = 0; fr.blockSize = 0; fr.lastSize = 0;
    lz4_pack_scan_kernel<<<1, 1024, 0, s>>>(sizes, nBlocks, outOff, headerBytes, fr);
    g_launches++;
    if (nBlocks > 0) {
        int64_t grid = nBlocks < 148 * 16 ? nBlocks : 148 * 16;
        lz4_pack_gather_kernel<<<(unsigned)grid, 256, 0, s>>>(slots, slotStride, sizes, nBlocks, packed, outOff, headerBytes, fr);
        g_launches++;
    }
    return (int)cudaGetLastError();
}

/* frame body: [LE32 header][payload] per block, stored-raw rule applied on the device (src = the blocks' source bytes) */
int lz4k_launch_pack_frame(const uint8_t* slots, int64_t slotStride, const int32_t* sizes, const uint8_t* src, int64_t srcStride,
                           int32_t blockSize, int32_t lastSize, int64_t nBlocks, uint8_t* packed, int64_t* outOff, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    FrameRule fr; fr.raw = src; fr.rawStride = srcStride; fr.blockSize = blockSize; fr.lastSize = lastSize;
    lz4_pack_scan_kernel<<<1, 1024, 0, s>>>(sizes, nBlocks, outOff, 4, fr);
    g_launches++;
    if (nBlocks > 0) {
        int64_t grid = nBlocks < 148 * 16 ? nBlocks : 148 * 16;
        lz4_pack_gather_kernel<<<(unsigned)grid, 256, 0, s>>>(slots, slotStride, sizes, nBlocks, packed, outOff, 4, fr);
        g_launches++;
    }
    return (int)cudaGetLastError();
}

}  // extern "C"

/*
 * lz4_api.c -- host side of liblz4_b200.so, in C (as the reference's lib/lz4.c is).
 *
 * Implements include/lz4_b200.h: the drop-in one-shot block API of lib/lz4.h over host pointers,
 * the batched device-pointer API, and the host-buffer batch calls.  Every codec call runs the
 * CUDA kernels of lz4_kernels.cu through the extern "C" launchers of lz4_kernels.h; there is no
 * CPU implementation of the codec in this library.
 */
#include "../../include/lz4_b200.h"
#include "lz4_kernels.h"

#include <cuda_runtime_api.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* trivial host-only entry points (lz4.c:749-752)                                              */
/* ------------------------------------------------------------------------------------------ */
int LZ4_versionNumber(void) { return LZ4B200_VERSION_NUMBER; }
const char* LZ4_versionString(void) { return "1.10.0"; }
int LZ4_sizeofState(void) { return LZ4B200_STATE_BYTES; }
int LZ4_compressBound(int inputSize)
{
    if ((unsigned)inputSize > (unsigned)LZ4B200_MAX_INPUT_SIZE) return 0;
    return inputSize + inputSize / 255 + 16;
}

/* ------------------------------------------------------------------------------------------ */
/* CUDA context shared by the host-pointer calls                                               */
/* ------------------------------------------------------------------------------------------ */
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static char g_cuda_error[256] = "";

static int cuda_fail(cudaError_t e, const char* where)
{
    snprintf(g_cuda_error, sizeof(g_cuda_error), "%s: %s", where, cudaGetErrorString(e));
    return LZ4B200_ERR_CUDA;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = cuda_fail(e_, #call); goto done; } } while (0)

const char* LZ4B200_last_cuda_error(void) { return g_cuda_error; }
uint64_t LZ4B200_launch_count(void) { return lz4k_launch_count(); }
/* not part of the public header: developer hook used by tests/perf/phase_timing.py */
__attribute__((visibility("default"))) int LZ4B200_debug_phase_cycles(unsigned long long* out8) { return lz4k_debug_phase_cycles(out8); }
int LZ4B200_peer_copy_async(void* d_dstPeer, int peerDevice, const void* d_src, size_t bytes, void* stream)
{
    static unsigned char enabled[64][64];
    int cur = 0;
    cudaError_t e;
    if (peerDevice < 0 || peerDevice >= 64 || (bytes && (!d_dstPeer || !d_src))) return LZ4B200_ERR_ARG;
    e = cudaGetDevice(&cur);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (cur >= 64) return LZ4B200_ERR_ARG;
    if (cur != peerDevice && !enabled[cur][peerDevice]) {
        int can = 0;
        e = cudaDeviceCanAccessPeer(&can, cur, peerDevice);
        if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceCanAccessPeer");
        if (!can) return LZ4B200_ERR_ARG;
        e = cudaDeviceEnablePeerAccess(peerDevice, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
        enabled[cur][peerDevice] = 1;
    }
    if (bytes == 0) return LZ4B200_OK;                   /* (a zero-byte call is the probe: peer access is enabled or refused) */
    e = cudaMemcpyAsync(d_dstPeer, d_src, bytes, cudaMemcpyDefault, (cudaStream_t)stream);
    return e == cudaSuccess ? LZ4B200_OK : cuda_fail(e, "cudaMemcpyAsync (peer)");
}

/* not part of the public header: developer hook used by tests/perf/enc_determinism.py */
__attribute__((visibility("default"))) int LZ4B200_debug_poison_smem(uint32_t pattern, int lo, int hi, void* stream)
{
    return lz4k_debug_poison_smem(pattern, lo, hi, stream);
}
/* not part of the public header: developer hook used by bench.py --ceiling (the rows kernel's skeleton without the decode) */
__attribute__((visibility("default"))) int LZ4B200_debug_ceiling(const void* d_src, const int64_t* d_srcOff, const int32_t* d_srcSize,
                                                                  void* d_dst, int64_t dstStride, int64_t nBlocks, int mode, void* stream)
{
    lz4k_decode_args a;
    memset(&a, 0, sizeof a);
    a.src = (const uint8_t*)d_src; a.srcOff = d_srcOff; a.srcSize = d_srcSize;
    a.dst = (uint8_t*)d_dst; a.dstStride = dstStride; a.nBlocks = nBlocks;
    return lz4k_launch_ceiling(&a, mode, stream);
}
int LZ4B200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

#define N_PIPE 4   /* chunks in flight for the host-buffer batch calls */

typedef struct {
    int ready;
    cudaStream_t stream[N_PIPE];
    /* growable device buffers, one set per pipeline slot */
    void* d_in[N_PIPE];   size_t in_cap[N_PIPE];
    void* d_out[N_PIPE];  size_t out_cap[N_PIPE];
    void* d_meta[N_PIPE]; size_t meta_cap[N_PIPE];
    void* d_ws[N_PIPE];   size_t ws_cap[N_PIPE];
    void* d_pack[N_PIPE]; size_t pack_cap[N_PIPE];   /* frame bodies assembled on the device */
    /* pinned host staging for the per-block tables, so that every copy of a chunk is truly
     * asynchronous (a pageable cudaMemcpyAsync blocks the host and serialises the pipeline) */
    void* h_meta[N_PIPE]; size_t hmeta_cap[N_PIPE];
    int32_t* pend_dst[N_PIPE]; int64_t pend_cnt[N_PIPE];   /* return values still to be handed to the caller */
} host_ctx;

static host_ctx g_ctx;

static int ctx_init(void)
{
    int rc = LZ4B200_OK, i;
    if (g_ctx.ready) return LZ4B200_OK;
    if (LZ4B200_device_count() <= 0) {
        snprintf(g_cuda_error, sizeof(g_cuda_error), "no CUDA device: liblz4_b200 has no CPU fallback");
        return LZ4B200_ERR_CUDA;
    }
    for (i = 0; i < N_PIPE; i++) CU(cudaStreamCreateWithFlags(&g_ctx.stream[i], cudaStreamNonBlocking));
    g_ctx.ready = 1;
done:
    return rc;
}

static int grow_pinned(void** p, size_t* cap, size_t need)
{
    int rc = LZ4B200_OK;
    if (need <= *cap) return rc;
    if (*p) { CU(cudaFreeHost(*p)); *p = NULL; *cap = 0; }
    need = (need + 65535) & ~(size_t)65535;
    CU(cudaHostAlloc(p, need, cudaHostAllocDefault));
    *cap = need;
done:
    return rc;
}

/* hand the return values of the chunk that last used `slot` to the caller (its stream is idle) */
static void flush_slot(int slot)
{
    if (g_ctx.pend_dst[slot]) {
        const int64_t cnt = g_ctx.pend_cnt[slot];
        const int32_t* h_ret = (const int32_t*)((char*)g_ctx.h_meta[slot] + (size_t)cnt * (sizeof(int64_t) + sizeof(int32_t)));
        memcpy(g_ctx.pend_dst[slot], h_ret, (size_t)cnt * sizeof(int32_t));
        g_ctx.pend_dst[slot] = NULL;
    }
}

static int grow(void** p, size_t* cap, size_t need)
{
    int rc = LZ4B200_OK;
    if (need <= *cap) return rc;
    if (*p) { CU(cudaFree(*p)); *p = NULL; *cap = 0; }
    need = (need + (size_t)(1 << 20)) & ~(size_t)((1 << 20) - 1);
    CU(cudaMalloc(p, need));
    *cap = need;
done:
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* batch layer: device pointers                                                                */
/* ------------------------------------------------------------------------------------------ */
size_t LZ4B200_decompress_workspace_bytes(int64_t nBlocks) { return lz4k_decode_workspace_bytes(nBlocks); }
size_t LZ4B200_decompress_workspace_bytes_for(int64_t nBlocks, int perBlockCaps, int32_t dstCap)
{
    return lz4k_decode_workspace_bytes_for(nBlocks, perBlockCaps, dstCap);
}

int LZ4B200_decompress_blocks_phased(const void* d_src, const int64_t* d_srcOff, const int32_t* d_srcSize,
                              void* d_dst, const int64_t* d_dstOff, int64_t dstStride,
                              const int32_t* d_dstCap, int32_t dstCap,
                              int32_t* d_outSize, int64_t nBlocks,
                              void* d_workspace, size_t workspaceBytes, int phases, void* stream)
{
    lz4k_decode_args a;
    cudaError_t e;
    if (nBlocks < 0) return LZ4B200_ERR_ARG;
    if (nBlocks == 0) return LZ4B200_OK;
    if (!d_src || !d_srcOff || !d_srcSize || !d_dst || !d_outSize || !d_workspace) return LZ4B200_ERR_ARG;
    if (workspaceBytes < lz4k_decode_workspace_bytes_min(nBlocks, d_dstCap != NULL, dstCap)) return LZ4B200_ERR_ARG;
    a.src = (const uint8_t*)d_src; a.srcOff = d_srcOff; a.srcSize = d_srcSize;
    a.dst = (uint8_t*)d_dst; a.dstOff = d_dstOff; a.dstStride = dstStride;
    a.dstCapArr = d_dstCap; a.dstCap = dstCap; a.outSize = d_outSize; a.nBlocks = nBlocks;
    a.workspace = d_workspace; a.workspaceBytes = workspaceBytes;
    if (phases < 1 || phases > 3) return LZ4B200_ERR_ARG;
    e = (cudaError_t)lz4k_launch_decode(&a, phases, stream);
    return e == cudaSuccess ? LZ4B200_OK : cuda_fail(e, "lz4k_launch_decode");
}

int LZ4B200_decompress_blocks(const void* d_src, const int64_t* d_srcOff, const int32_t* d_srcSize,
                              void* d_dst, const int64_t* d_dstOff, int64_t dstStride,
                              const int32_t* d_dstCap, int32_t dstCap,
                              int32_t* d_outSize, int64_t nBlocks,
                              void* d_workspace, size_t workspaceBytes, void* stream)
{
    return LZ4B200_decompress_blocks_phased(d_src, d_srcOff, d_srcSize, d_dst, d_dstOff, dstStride, d_dstCap, dstCap,
                                            d_outSize, nBlocks, d_workspace, workspaceBytes, 3, stream);
}

int LZ4B200_compress_blocks(const void* d_src, int64_t srcStride, const int32_t* d_srcSize, int32_t srcSize,
                            void* d_dst, int64_t dstStride, int32_t dstCap, int acceleration,
                            int32_t* d_outSize, int64_t nBlocks, void* stream)
{
    lz4k_encode_args a;
    cudaError_t e;
    if (nBlocks < 0) return LZ4B200_ERR_ARG;
    if (nBlocks == 0) return LZ4B200_OK;
    if (!d_dst || !d_outSize) return LZ4B200_ERR_ARG;
    a.src = (const uint8_t*)d_src; a.srcStride = srcStride; a.srcSizeArr = d_srcSize; a.srcSize = srcSize;
    a.dst = (uint8_t*)d_dst; a.dstStride = dstStride; a.dstCap = dstCap; a.acceleration = acceleration;
    a.outSize = d_outSize; a.nBlocks = nBlocks;
    e = (cudaError_t)lz4k_launch_encode(&a, stream);
    return e == cudaSuccess ? LZ4B200_OK : cuda_fail(e, "lz4k_launch_encode");
}

int LZ4B200_compress_blocks_parallel(const void* d_src, int64_t srcStride, const int32_t* d_srcSize, int32_t srcSize,
                                     void* d_dst, int64_t dstStride, int32_t dstCap, int acceleration,
                                     int32_t* d_outSize, int64_t nBlocks, void* stream)
{
    lz4k_encode_args a;
    cudaError_t e;
    if (nBlocks < 0) return LZ4B200_ERR_ARG;
    if (nBlocks == 0) return LZ4B200_OK;
    if (!d_dst || !d_outSize) return LZ4B200_ERR_ARG;
    a.src = (const uint8_t*)d_src; a.srcStride = srcStride; a.srcSizeArr = d_srcSize; a.srcSize = srcSize;
    a.dst = (uint8_t*)d_dst; a.dstStride = dstStride; a.dstCap = dstCap; a.acceleration = acceleration;
    a.outSize = d_outSize; a.nBlocks = nBlocks;
    /* the parallel parse stages a whole block in shared memory: blocks above 64 KB take the byte-identical encoder */
    e = (cudaError_t)(srcSize <= 65536 ? lz4k_launch_encode_par(&a, stream) : lz4k_launch_encode(&a, stream));
    return e == cudaSuccess ? LZ4B200_OK : cuda_fail(e, "lz4k_launch_encode_par");
}

int LZ4B200_pack_blocks(const void* d_slots, int64_t slotStride, const int32_t* d_sizes, int64_t nBlocks,
                        void* d_packed, int64_t* d_outOff, int headerBytes, void* stream)
{
    cudaError_t e;
    if (nBlocks < 0 || !d_outOff || (headerBytes != 0 && headerBytes != 4)) return LZ4B200_ERR_ARG;
    if (nBlocks > 0 && (!d_slots || !d_sizes || !d_packed)) return LZ4B200_ERR_ARG;
    e = (cudaError_t)lz4k_launch_pack((const uint8_t*)d_slots, slotStride, d_sizes, nBlocks, (uint8_t*)d_packed,
                                      d_outOff, headerBytes, stream);
    return e == cudaSuccess ? LZ4B200_OK : cuda_fail(e, "lz4k_launch_pack");
}

int LZ4B200_pack_frame_blocks(const void* d_slots, int64_t slotStride, const int32_t* d_sizes, const void* d_src, int64_t srcStride,
                              int32_t blockSize, int32_t lastSize, int64_t nBlocks, void* d_body, int64_t* d_outOff, void* stream)
{
    cudaError_t e;
    if (nBlocks < 0 || !d_outOff || blockSize <= 0 || lastSize < 0 || lastSize > blockSize) return LZ4B200_ERR_ARG;
    if (nBlocks > 0 && (!d_slots || !d_sizes || !d_src || !d_body)) return LZ4B200_ERR_ARG;
    e = (cudaError_t)lz4k_launch_pack_frame((const uint8_t*)d_slots, slotStride, d_sizes, (const uint8_t*)d_src, srcStride,
                                            blockSize, lastSize, nBlocks, (uint8_t*)d_body, d_outOff, stream);
    return e == cudaSuccess ? LZ4B200_OK : cuda_fail(e, "lz4k_launch_pack_frame");
}

/* ------------------------------------------------------------------------------------------ */
/* host-buffer batch calls: H2D -> kernels -> D2H, N_PIPE chunks in flight                     */
/* ------------------------------------------------------------------------------------------ */
/* target uncompressed bytes per pipeline chunk: large enough that a chunk's kernels run at batch speed, small enough that
 * the first H2D and the last D2H (which nothing overlaps) are a small part of the call.  LZ4B200_HOST_CHUNK_MB overrides
 * it (developer knob for tests/perf). */
static int64_t chunk_out_bytes(void)
{
    static int64_t v = 0;
    if (v == 0) {
        const char* e = getenv("LZ4B200_HOST_CHUNK_MB");
        long mb = e ? atol(e) : 0;
        v = (int64_t)((mb >= 1 && mb <= 4096) ? mb : 64) << 20;
    }
    return v;
}
#define CHUNK_OUT_BYTES chunk_out_bytes()

int LZ4B200_decompress_blocks_host(const void* h_src, const int64_t* h_srcOff, const int32_t* h_srcSize,
                                   void* h_dst, int64_t dstStride, int32_t dstCap,
                                   int32_t* h_outSize, int64_t nBlocks)
{
    int rc = LZ4B200_OK;
    int64_t perChunk, first;
    int slot = 0, i;
    if (nBlocks < 0 || dstCap < 0 || dstStride < dstCap) return LZ4B200_ERR_ARG;
    if (nBlocks == 0) return LZ4B200_OK;
    if (!h_src || !h_srcOff || !h_srcSize || !h_dst || !h_outSize) return LZ4B200_ERR_ARG;
    pthread_mutex_lock(&g_lock);
    if ((rc = ctx_init()) != LZ4B200_OK) goto done;

    perChunk = CHUNK_OUT_BYTES / (dstStride > 0 ? dstStride : 1);
    if (perChunk < 1) perChunk = 1;
    for (first = 0; first < nBlocks; first += perChunk, slot = (slot + 1) % N_PIPE) {
        const int64_t cnt = (nBlocks - first < perChunk) ? nBlocks - first : perChunk;
        cudaStream_t st = g_ctx.stream[slot];
        int64_t lo = h_srcOff[first], hi = lo, k;
        size_t inBytes, outBytes = (size_t)(cnt * dstStride), metaBytes, wsBytes;
        /* the chunk's compressed bytes span [lo, hi) of h_src (blocks may be in any order) */
        for (k = first; k < first + cnt; k++) {
            if (h_srcSize[k] < 0) { rc = LZ4B200_ERR_ARG; goto done; }
            if (h_srcOff[k] < lo) lo = h_srcOff[k];
            if (h_srcOff[k] + h_srcSize[k] > hi) hi = h_srcOff[k] + h_srcSize[k];
        }
        inBytes = (size_t)(hi - lo);
        metaBytes = (size_t)cnt * (sizeof(int64_t) + 2 * sizeof(int32_t));
        wsBytes = lz4k_decode_workspace_bytes_for(cnt, 0, dstCap);
        CU(cudaStreamSynchronize(st));                      /* slot reuse: previous chunk on it is finished */
        flush_slot(slot);
        if ((rc = grow(&g_ctx.d_in[slot], &g_ctx.in_cap[slot], inBytes + 16)) != LZ4B200_OK) goto done;
        if ((rc = grow(&g_ctx.d_out[slot], &g_ctx.out_cap[slot], outBytes + 16)) != LZ4B200_OK) goto done;
        if ((rc = grow(&g_ctx.d_meta[slot], &g_ctx.meta_cap[slot], metaBytes + 16)) != LZ4B200_OK) goto done;
        if ((rc = grow(&g_ctx.d_ws[slot], &g_ctx.ws_cap[slot], wsBytes)) != LZ4B200_OK) goto done;
        if ((rc = grow_pinned(&g_ctx.h_meta[slot], &g_ctx.hmeta_cap[slot], metaBytes + 16)) != LZ4B200_OK) goto done;
        {
            int64_t* d_off = (int64_t*)g_ctx.d_meta[slot];
            int32_t* d_size = (int32_t*)(d_off + cnt);
            int32_t* d_ret = d_size + cnt;
            int64_t* h_off = (int64_t*)g_ctx.h_meta[slot];          /* pinned: [offsets | sizes | return values] */
            int32_t* h_size = (int32_t*)(h_off + cnt);
            int32_t* h_ret = h_size + cnt;
            for (k = 0; k < cnt; k++) h_off[k] = h_srcOff[first + k] - lo;   /* relative to the chunk's first byte */
            memcpy(h_size, h_srcSize + first, (size_t)cnt * sizeof(int32_t));
            CU(cudaMemcpyAsync(d_off, h_off, (size_t)cnt * (sizeof(int64_t) + sizeof(int32_t)), cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(g_ctx.d_in[slot], (const char*)h_src + lo, inBytes, cudaMemcpyHostToDevice, st));
            rc = LZ4B200_decompress_blocks(g_ctx.d_in[slot], d_off, d_size, g_ctx.d_out[slot], NULL, dstStride,
                                           NULL, dstCap, d_ret, cnt, g_ctx.d_ws[slot], g_ctx.ws_cap[slot], st);
            if (rc != LZ4B200_OK) goto done;
            CU(cudaMemcpyAsync((char*)h_dst + first * dstStride, g_ctx.d_out[slot], outBytes, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(h_ret, d_ret, (size_t)cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            g_ctx.pend_dst[slot] = h_outSize + first;
            g_ctx.pend_cnt[slot] = cnt;
        }
    }
    for (i = 0; i < N_PIPE; i++) { CU(cudaStreamSynchronize(g_ctx.stream[i])); flush_slot(i); }
    goto unlock;
done:
    for (i = 0; i < N_PIPE; i++) { cudaStreamSynchronize(g_ctx.stream[i]); g_ctx.pend_dst[i] = NULL; }
unlock:
    pthread_mutex_unlock(&g_lock);
    return rc;
}

int LZ4B200_compress_blocks_host(const void* h_src, int64_t srcStride, int32_t srcSize, int64_t lastSize,
                                 void* h_dst, int64_t dstStride, int32_t dstCap, int acceleration,
                                 int32_t* h_outSize, int64_t nBlocks)
{
    int rc = LZ4B200_OK;
    int64_t perChunk, first;
    int slot = 0, i;
    if (nBlocks < 0 || srcSize < 0 || lastSize < 0 || lastSize > srcSize || srcStride < srcSize) return LZ4B200_ERR_ARG;
    if (nBlocks == 0) return LZ4B200_OK;
    if (!h_dst || !h_outSize || (!h_src && srcSize > 0)) return LZ4B200_ERR_ARG;
    pthread_mutex_lock(&g_lock);
    if ((rc = ctx_init()) != LZ4B200_OK) goto done;

    perChunk = CHUNK_OUT_BYTES / (srcStride > 0 ? srcStride : 1);
    if (perChunk < 1) perChunk = 1;
    for (first = 0; first < nBlocks; first += perChunk, slot = (slot + 1) % N_PIPE) {
        const int64_t cnt = (nBlocks - first < perChunk) ? nBlocks - first : perChunk;
        const int isLast = (first + cnt == nBlocks);
        cudaStream_t st = g_ctx.stream[slot];
        size_t inBytes = (size_t)((cnt - 1) * srcStride + (isLast ? lastSize : srcSize));
        size_t slotBytes = (size_t)(cnt * dstStride);
        /* same table layout as the decode call: [unused i64 | source sizes i32 | return values i32] */
        size_t metaBytes = (size_t)cnt * (sizeof(int64_t) + 2 * sizeof(int32_t));
        int32_t *d_sizes, *d_ret, *h_sizes, *h_ret;
        int64_t k;
        CU(cudaStreamSynchronize(st));
        flush_slot(slot);
        if ((rc = grow(&g_ctx.d_in[slot], &g_ctx.in_cap[slot], inBytes + 16)) != LZ4B200_OK) goto done;
        if ((rc = grow(&g_ctx.d_out[slot], &g_ctx.out_cap[slot], slotBytes + 16)) != LZ4B200_OK) goto done;
        if ((rc = grow(&g_ctx.d_meta[slot], &g_ctx.meta_cap[slot], metaBytes + 16)) != LZ4B200_OK) goto done;
        if ((rc = grow_pinned(&g_ctx.h_meta[slot], &g_ctx.hmeta_cap[slot], metaBytes + 16)) != LZ4B200_OK) goto done;
        d_sizes = (int32_t*)((int64_t*)g_ctx.d_meta[slot] + cnt);
        d_ret = d_sizes + cnt;
        h_sizes = (int32_t*)((int64_t*)g_ctx.h_meta[slot] + cnt);
        h_ret = h_sizes + cnt;
        if (inBytes) CU(cudaMemcpyAsync(g_ctx.d_in[slot], (const char*)h_src + first * srcStride, inBytes, cudaMemcpyHostToDevice, st));
        if (isLast && lastSize != srcSize) {
            /* per-block size table only needed for the ragged final block */
            for (k = 0; k < cnt; k++) h_sizes[k] = srcSize;
            h_sizes[cnt - 1] = (int32_t)lastSize;
            CU(cudaMemcpyAsync(d_sizes, h_sizes, (size_t)cnt * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            rc = LZ4B200_compress_blocks(g_ctx.d_in[slot], srcStride, d_sizes, srcSize, g_ctx.d_out[slot], dstStride,
                                         dstCap, acceleration, d_ret, cnt, st);
        } else {
            rc = LZ4B200_compress_blocks(g_ctx.d_in[slot], srcStride, NULL, srcSize, g_ctx.d_out[slot], dstStride,
                                         dstCap, acceleration, d_ret, cnt, st);
        }
        if (rc != LZ4B200_OK) goto done;
        CU(cudaMemcpyAsync((char*)h_dst + first * dstStride, g_ctx.d_out[slot], slotBytes, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h_ret, d_ret, (size_t)cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        g_ctx.pend_dst[slot] = h_outSize + first;
        g_ctx.pend_cnt[slot] = cnt;
    }
    for (i = 0; i < N_PIPE; i++) { CU(cudaStreamSynchronize(g_ctx.stream[i])); flush_slot(i); }
    goto unlock;
done:
    for (i = 0; i < N_PIPE; i++) { cudaStreamSynchronize(g_ctx.stream[i]); g_ctx.pend_dst[i] = NULL; }
unlock:
    pthread_mutex_unlock(&g_lock);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* drop-in layer: one block, host pointers, synchronous                                        */
/* ------------------------------------------------------------------------------------------ */
static int one_block_compress(const char* src, char* dst, int srcSize, int dstCapacity, int acceleration)
{
    int rc = LZ4B200_OK, ret = 0;
    int32_t* d_ret;
    cudaStream_t st;
    size_t capBytes;
    /* argument rules of LZ4_compress_generic, lz4.c:1360-1372, decided on the host so that a bad
     * size never reaches the device */
    if ((unsigned)srcSize > (unsigned)LZ4B200_MAX_INPUT_SIZE) return 0;
    if (dstCapacity <= 0) return 0;          /* nothing can be written (lz4.c:1362 and the limited checks) */
    if (dst == NULL) return 0;
    if (srcSize > 0 && src == NULL) return 0;
    pthread_mutex_lock(&g_lock);
    if ((rc = ctx_init()) != LZ4B200_OK) goto done;
    st = g_ctx.stream[0];
    {
        int bound = LZ4_compressBound(srcSize);
        capBytes = (size_t)(dstCapacity < bound ? dstCapacity : bound);
    }
    if ((rc = grow(&g_ctx.d_in[0], &g_ctx.in_cap[0], (size_t)srcSize + 16)) != LZ4B200_OK) goto done;
    if ((rc = grow(&g_ctx.d_out[0], &g_ctx.out_cap[0], capBytes + 16)) != LZ4B200_OK) goto done;
    if ((rc = grow(&g_ctx.d_meta[0], &g_ctx.meta_cap[0], 64)) != LZ4B200_OK) goto done;
    d_ret = (int32_t*)g_ctx.d_meta[0];
    /* copy-in completes before any output is produced: in-place calls (lz4.h:619-678) are safe */
    if (srcSize) CU(cudaMemcpyAsync(g_ctx.d_in[0], src, (size_t)srcSize, cudaMemcpyHostToDevice, st));
    rc = LZ4B200_compress_blocks(g_ctx.d_in[0], 0, NULL, srcSize, g_ctx.d_out[0], 0, dstCapacity, acceleration, d_ret, 1, st);
    if (rc != LZ4B200_OK) goto done;
    CU(cudaMemcpyAsync(&ret, d_ret, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (ret > 0) {
        CU(cudaMemcpyAsync(dst, g_ctx.d_out[0], (size_t)ret, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
done:
    pthread_mutex_unlock(&g_lock);
    return rc == LZ4B200_OK ? ret : 0;
}

int LZ4_compress_fast(const char* src, char* dst, int srcSize, int dstCapacity, int acceleration)
{
    return one_block_compress(src, dst, srcSize, dstCapacity, acceleration);
}

int LZ4_compress_default(const char* src, char* dst, int srcSize, int dstCapacity)
{
    return one_block_compress(src, dst, srcSize, dstCapacity, 1);
}

static int state_is_valid(const void* state)
{
    /* LZ4_initStream, lz4.c:1552-1560: non-NULL and aligned like LZ4_stream_t (8 bytes) */
    return state != NULL && (((uintptr_t)state) & 7u) == 0;
}

int LZ4_compress_fast_extState(void* state, const char* src, char* dst, int srcSize, int dstCapacity, int acceleration)
{
    if (!state_is_valid(state)) return 0;
    return one_block_compress(src, dst, srcSize, dstCapacity, acceleration);
}

int LZ4_compress_fast_extState_fastReset(void* state, const char* src, char* dst, int srcSize, int dstCapacity, int acceleration)
{
    if (!state_is_valid(state)) return 0;
    return one_block_compress(src, dst, srcSize, dstCapacity, acceleration);
}

int LZ4_decompress_safe(const char* src, char* dst, int compressedSize, int dstCapacity)
{
    int rc = LZ4B200_OK, ret = -1;
    cudaStream_t st;
    int64_t* d_off; int32_t* d_size; int32_t* d_ret;
    int64_t zero = 0;
    /* lz4.c:2036, :2064-2069 decided on the host */
    if (src == NULL || dstCapacity < 0) return -1;
    if (dstCapacity == 0) return (compressedSize == 1 && src[0] == 0) ? 0 : -1;
    if (compressedSize <= 0) return -1;
    if (dst == NULL) return -1;
    pthread_mutex_lock(&g_lock);
    if ((rc = ctx_init()) != LZ4B200_OK) goto done;
    st = g_ctx.stream[0];
    if ((rc = grow(&g_ctx.d_in[0], &g_ctx.in_cap[0], (size_t)compressedSize + 16)) != LZ4B200_OK) goto done;
    if ((rc = grow(&g_ctx.d_out[0], &g_ctx.out_cap[0], (size_t)dstCapacity + 16)) != LZ4B200_OK) goto done;
    if ((rc = grow(&g_ctx.d_meta[0], &g_ctx.meta_cap[0], 64)) != LZ4B200_OK) goto done;
    if ((rc = grow(&g_ctx.d_ws[0], &g_ctx.ws_cap[0], lz4k_decode_workspace_bytes_for(1, 0, dstCapacity))) != LZ4B200_OK) goto done;
    d_off = (int64_t*)g_ctx.d_meta[0];
    d_size = (int32_t*)(d_off + 1);
    d_ret = d_size + 1;
    CU(cudaMemcpyAsync(d_off, &zero, sizeof(zero), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_size, &compressedSize, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    /* the whole input is on the device before the first output byte is written back: in-place
     * decompression (lz4.h:619-678, tests/fuzzer.c:1177-1187) needs no special handling */
    CU(cudaMemcpyAsync(g_ctx.d_in[0], src, (size_t)compressedSize, cudaMemcpyHostToDevice, st));
    rc = LZ4B200_decompress_blocks(g_ctx.d_in[0], d_off, d_size, g_ctx.d_out[0], NULL, 0, NULL, dstCapacity,
                                   d_ret, 1, g_ctx.d_ws[0], g_ctx.ws_cap[0], st);
    if (rc != LZ4B200_OK) goto done;
    CU(cudaMemcpyAsync(&ret, d_ret, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (ret > 0) {
        CU(cudaMemcpyAsync(dst, g_ctx.d_out[0], (size_t)ret, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
done:
    pthread_mutex_unlock(&g_lock);
    return rc == LZ4B200_OK ? ret : -1;
}

int LZ4_decompress_safe_usingDict(const char* src, char* dst, int compressedSize, int dstCapacity,
                                  const char* dictStart, int dictSize)
{
    (void)dictStart;
    if (dictSize == 0) return LZ4_decompress_safe(src, dst, compressedSize, dstCapacity);   /* lz4.c:2721-2722 */
    return -1;   /* prefix / external dictionaries: SURVEY.md section 8 (f-4), not on the GPU path */
}

/* ------------------------------------------------------------------------------------------ */
/* frame layer: independent-block LZ4 frames over host buffers          */
/* ------------------------------------------------------------------------------------------ */
#define FRAME_MAGIC 0x184D2204u            /* lz4frame.c: LZ4F_MAGICNUMBER */
#define FRAME_MAGIC_SKIPPABLE 0x184D2A50u  /* .. LZ4F_MAGIC_SKIPPABLE_START */

static uint32_t rd_le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static void wr_le32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

/* XXH32 of a short message (len < 16), as published in the xxHash specification; the frame header
 * checksum is its second byte (lz4frame.c:781-809: HC = (XXH32(descriptor, 0) >> 8) & 0xFF). */
static uint32_t xxh32_short(const uint8_t* p, size_t len, uint32_t seed)
{
    const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    const uint8_t* end = p + len;
    uint32_t h = seed + P5 + (uint32_t)len;
    while (p + 4 <= end) { h += rd_le32(p) * P3; h = rotl32(h, 17) * P4; p += 4; }
    while (p < end) { h += (uint32_t)(*p) * P5; h = rotl32(h, 11) * P1; p++; }
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    return h;
}

static int64_t frame_block_bytes(int bsid) { return (int64_t)1 << (8 + 2 * bsid); }   /* 4..7 -> 64 KB..4 MB (lz4frame.c:333-342) */

/* LZ4F_optimalBSID, lz4frame.c:362-373 */
static int frame_optimal_bsid(int requested, int64_t srcSize)
{
    int proposed = 4;
    int64_t maxBlock = 64 * 1024;
    while (requested > proposed) {
        if (srcSize <= maxBlock) return proposed;
        proposed++;
        maxBlock <<= 2;
    }
    return requested;
}

int64_t LZ4B200_compressFrameBound(int64_t srcSize, int blockSizeID)
{
    int64_t bs, nBlocks;
    if (srcSize < 0) return LZ4B200_ERR_ARG;
    if (blockSizeID == 0) blockSizeID = 4;
    if (blockSizeID < 4 || blockSizeID > 7) return LZ4B200_ERR_ARG;
    bs = frame_block_bytes(blockSizeID);
    nBlocks = srcSize / bs + 1;
    return 19 + nBlocks * 4 + srcSize + 4 + 4;     /* max header + block headers + data + EndMark (+ slack) */
}

int64_t LZ4B200_compressFrame_host(const void* h_src, int64_t srcSize, void* h_dst, int64_t dstCapacity,
                                   int blockSizeID, int compressionLevel, int contentSizeFlag)
{
    const uint8_t* src = (const uint8_t*)h_src;
    uint8_t* dst = (uint8_t*)h_dst;
    int64_t bs, nFull, lastSize, nBlocks, op = 0;
    int accel, bsid;
    if (srcSize < 0 || (!h_src && srcSize > 0) || !h_dst) return LZ4B200_ERR_ARG;
    if (blockSizeID == 0) blockSizeID = 4;
    if (blockSizeID < 4 || blockSizeID > 7) return LZ4B200_ERR_ARG;
    if (compressionLevel >= 2) return LZ4B200_ERR_UNSUPPORTED;          /* LZ4HC levels (lz4frame.c:952-962) */
    if (dstCapacity < LZ4B200_compressFrameBound(srcSize, blockSizeID)) return LZ4B200_ERR_DSTSIZE;
    accel = compressionLevel < 0 ? -compressionLevel + 1 : 1;           /* lz4frame.c:913 */
    bsid = frame_optimal_bsid(blockSizeID, srcSize);                     /* lz4frame.c:447 */
    bs = frame_block_bytes(bsid);
    if (srcSize == 0) contentSizeFlag = 0;                               /* lz4frame.c:445: contentSize = srcSize = 0 -> no field */

    /* frame header, lz4frame.c:781-809 */
    wr_le32(dst, FRAME_MAGIC); op = 4;
    dst[op++] = (uint8_t)((1u << 6) | (1u << 5) | ((contentSizeFlag ? 1u : 0u) << 3));   /* version 01, B.Indep */
    dst[op++] = (uint8_t)((uint32_t)bsid << 4);
    if (contentSizeFlag) { int i; for (i = 0; i < 8; i++) dst[op++] = (uint8_t)((uint64_t)srcSize >> (8 * i)); }
    dst[op] = (uint8_t)(xxh32_short(dst + 4, (size_t)(op - 4), 0) >> 8);
    op++;

    nFull = srcSize / bs;
    lastSize = srcSize - nFull * bs;
    nBlocks = nFull + (lastSize > 0);
    if (nBlocks > 0) {
        /* Groups of ~256 MiB of source go through the device: H2D, one compress launch (dstCapacity = blockSize - 1; a
         * result of 0 or >= the block's size means "store raw", LZ4F_makeBlock lz4frame.c:883-908), the frame body
         * ([LE32 header][payload] per block) is assembled ON THE DEVICE, and one D2H per group lands it in h_dst.  Group
         * g+1 is issued before group g is retired, so its copies and kernels overlap g's D2H. */
        int64_t group = ((int64_t)256 << 20) / bs, nGroups, g;
        const int64_t slotStride = (bs + 15) & ~(int64_t)15;
        int rc = LZ4B200_OK;
        if (group < 1) group = 1;
        nGroups = (nBlocks + group - 1) / group;
        pthread_mutex_lock(&g_lock);
        if ((rc = ctx_init()) != LZ4B200_OK) { pthread_mutex_unlock(&g_lock); return rc; }
        for (g = 0; g <= nGroups && rc == LZ4B200_OK; g++) {
            if (g < nGroups) {                                         /* ---- issue group g ---- */
                const int slot = (int)(g % N_PIPE);
                const int64_t firstB = g * group;
                const int64_t cnt = (nBlocks - firstB < group) ? nBlocks - firstB : group;
                const int hasLast = (firstB + cnt == nBlocks) && lastSize > 0;
                const int64_t inBytes = (cnt - 1) * bs + (hasLast ? lastSize : bs);
                cudaStream_t st = g_ctx.stream[slot];
                int64_t* d_off; int32_t* d_sz; int32_t* d_ssz;
                if (cudaStreamSynchronize(st) != cudaSuccess) { rc = LZ4B200_ERR_CUDA; break; }
                if ((rc = grow(&g_ctx.d_in[slot], &g_ctx.in_cap[slot], (size_t)inBytes + 16)) != LZ4B200_OK) break;
                if ((rc = grow(&g_ctx.d_out[slot], &g_ctx.out_cap[slot], (size_t)(cnt * slotStride) + 16)) != LZ4B200_OK) break;
                if ((rc = grow(&g_ctx.d_pack[slot], &g_ctx.pack_cap[slot], (size_t)(cnt * (bs + 4)) + 16)) != LZ4B200_OK) break;
                if ((rc = grow(&g_ctx.d_meta[slot], &g_ctx.meta_cap[slot], (size_t)(cnt + 1) * 16 + 64)) != LZ4B200_OK) break;
                if ((rc = grow_pinned(&g_ctx.h_meta[slot], &g_ctx.hmeta_cap[slot], (size_t)cnt * 4 + 64)) != LZ4B200_OK) break;
                d_off = (int64_t*)g_ctx.d_meta[slot];
                d_sz = (int32_t*)(d_off + cnt + 1);
                d_ssz = d_sz + cnt;
                if (cudaMemcpyAsync(g_ctx.d_in[slot], src + firstB * bs, (size_t)inBytes, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = LZ4B200_ERR_CUDA; break; }
                if (hasLast) {                                           /* per-block sizes only for the group with the ragged block */
                    int32_t* h_ssz = (int32_t*)((char*)g_ctx.h_meta[slot] + 64);
                    int64_t k;
                    for (k = 0; k < cnt; k++) h_ssz[k] = (int32_t)bs;
                    h_ssz[cnt - 1] = (int32_t)lastSize;
                    if (cudaMemcpyAsync(d_ssz, h_ssz, (size_t)cnt * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = LZ4B200_ERR_CUDA; break; }
                }
                rc = LZ4B200_compress_blocks(g_ctx.d_in[slot], bs, hasLast ? d_ssz : NULL, (int32_t)bs, g_ctx.d_out[slot], slotStride,
                                             (int32_t)(bs - 1), accel, d_sz, cnt, st);
                if (rc != LZ4B200_OK) break;
                rc = LZ4B200_pack_frame_blocks(g_ctx.d_out[slot], slotStride, d_sz, g_ctx.d_in[slot], bs, (int32_t)bs,
                                               hasLast ? (int32_t)lastSize : 0, cnt, g_ctx.d_pack[slot], d_off, st);
                if (rc != LZ4B200_OK) break;
                if (cudaMemcpyAsync(g_ctx.h_meta[slot], d_off + cnt, sizeof(int64_t), cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = LZ4B200_ERR_CUDA; break; }
            }
            if (g > 0) {                                               /* ---- retire group g-1: its size is known once its stream is idle ---- */
                const int slot = (int)((g - 1) % N_PIPE);
                cudaStream_t st = g_ctx.stream[slot];
                int64_t total;
                if (cudaStreamSynchronize(st) != cudaSuccess) { rc = LZ4B200_ERR_CUDA; break; }
                total = *(const int64_t*)g_ctx.h_meta[slot];
                if (op + total + 4 > dstCapacity) { rc = LZ4B200_ERR_DSTSIZE; break; }
                if (cudaMemcpyAsync(dst + op, g_ctx.d_pack[slot], (size_t)total, cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = LZ4B200_ERR_CUDA; break; }
                op += total;
            }
        }
        {   int i; for (i = 0; i < N_PIPE; i++) if (cudaStreamSynchronize(g_ctx.stream[i]) != cudaSuccess && rc == LZ4B200_OK) rc = LZ4B200_ERR_CUDA; }
        pthread_mutex_unlock(&g_lock);
        if (rc == LZ4B200_ERR_CUDA) cuda_fail(cudaGetLastError(), "LZ4B200_compressFrame_host");
        if (rc != LZ4B200_OK) return rc;
    }
    wr_le32(dst + op, 0); op += 4;                                        /* EndMark, lz4frame.c:1222 */
    return op;
}

int64_t LZ4B200_decompressFrame_host(const void* h_src, int64_t srcSize, void* h_dst, int64_t dstCapacity,
                                     int64_t* consumed)
{
    const uint8_t* src = (const uint8_t*)h_src;
    uint8_t* dst = (uint8_t*)h_dst;
    int64_t ip, bs, nBlocks = 0, capBlocks = 0, k, total = 0, contentSize = -1;
    int64_t* off = NULL; int32_t* csz = NULL; int32_t* rets = NULL; uint8_t* raw = NULL;
    int64_t result = LZ4B200_ERR_FRAME;
    uint32_t flg, bd, hdrLen;
    int bsid;
    if (!h_src || srcSize < 0 || dstCapacity < 0 || (!h_dst && dstCapacity > 0)) return LZ4B200_ERR_ARG;
    if (srcSize < 7) return LZ4B200_ERR_FRAME;                            /* minFHSize, lz4frame.h:280 */
    if ((rd_le32(src) & 0xFFFFFFF0u) == FRAME_MAGIC_SKIPPABLE) return LZ4B200_ERR_UNSUPPORTED;
    if (rd_le32(src) != FRAME_MAGIC) return LZ4B200_ERR_FRAME;
    flg = src[4];
    if (((flg >> 6) & 3) != 1 || ((flg >> 1) & 1)) return LZ4B200_ERR_FRAME;   /* version, reserved bit */
    hdrLen = 7 + ((flg >> 3) & 1 ? 8 : 0) + ((flg & 1) ? 4 : 0);
    if (srcSize < hdrLen) return LZ4B200_ERR_FRAME;
    bd = src[5];
    bsid = (int)((bd >> 4) & 7);
    if ((bd >> 7) || (bd & 15) || bsid < 4) return LZ4B200_ERR_FRAME;     /* lz4frame.c:1406-1412 */
    if ((uint8_t)(xxh32_short(src + 4, hdrLen - 5, 0) >> 8) != src[hdrLen - 1]) return LZ4B200_ERR_FRAME;
    if (!((flg >> 5) & 1)) return LZ4B200_ERR_UNSUPPORTED;                 /* linked blocks (f-4) */
    if (((flg >> 4) & 1) || ((flg >> 2) & 1) || (flg & 1)) return LZ4B200_ERR_UNSUPPORTED;   /* checksums (f-3), dictID */
    if ((flg >> 3) & 1) { int i; uint64_t v = 0; for (i = 0; i < 8; i++) v |= (uint64_t)src[6 + i] << (8 * i); contentSize = (int64_t)v; }
    bs = frame_block_bytes(bsid);

    /* walk the chain of block headers (serial, O(#blocks)), lz4frame.c:1729-1758 */
    ip = hdrLen;
    for (;;) {
        uint32_t h; int64_t sz;
        if (ip + 4 > srcSize) goto done;
        h = rd_le32(src + ip); ip += 4;
        if (h == 0) break;                                                /* EndMark */
        sz = (int64_t)(h & 0x7FFFFFFFu);
        if (sz > bs || ip + sz > srcSize) goto done;                       /* lz4frame.c:1745: maxBlockSize_invalid */
        if (nBlocks == capBlocks) {
            int64_t* o2; int32_t* c2; uint8_t* r2;
            capBlocks = capBlocks ? capBlocks * 2 : 1024;
            o2 = (int64_t*)realloc(off, (size_t)capBlocks * sizeof(int64_t));
            if (o2) off = o2;
            c2 = (int32_t*)realloc(csz, (size_t)capBlocks * sizeof(int32_t));
            if (c2) csz = c2;
            r2 = (uint8_t*)realloc(raw, (size_t)capBlocks);
            if (r2) raw = r2;
            if (!o2 || !c2 || !r2) { result = LZ4B200_ERR_ARG; goto done; }     /* (the old blocks are freed at `done`) */
        }
        off[nBlocks] = ip; csz[nBlocks] = (int32_t)sz; raw[nBlocks] = (uint8_t)(h >> 31);
        nBlocks++;
        ip += sz;
    }
    if (consumed) *consumed = ip;
    if (nBlocks == 0) { result = (contentSize > 0) ? LZ4B200_ERR_FRAME : 0; goto done; }

    /* Every block is decoded with dstCapacity = maxBlockSize like lz4frame.c:1901-1904 does.  The frame is untrusted
     * input, so memory is bounded whatever it claims: blocks are decoded in GROUPS of at most ~256 MiB of block slots,
     * straight into h_dst while a group's slots (count x maxBlockSize) fit behind the bytes written so far -- short
     * blocks are then closed up in place -- and through one reusable bounce buffer otherwise (the tail of the caller's
     * buffer, or a flushed stream of many short blocks); the running total is checked after every group. */
    rets = (int32_t*)malloc((size_t)nBlocks * sizeof(int32_t));
    if (!rets) { result = LZ4B200_ERR_ARG; goto done; }
    {
        int64_t G = ((int64_t)256 << 20) / bs, g0;
        uint8_t* bounce = NULL;
        int32_t* tmpSz = NULL;
        if (G < 1) G = 1;
        if (G > nBlocks) G = nBlocks;
        tmpSz = (int32_t*)malloc((size_t)G * sizeof(int32_t));
        if (!tmpSz) { result = LZ4B200_ERR_ARG; goto done; }
        for (g0 = 0; g0 < nBlocks; g0 += G) {
            const int64_t cnt = (nBlocks - g0 < G) ? nBlocks - g0 : G;
            const int direct = (total + cnt * bs <= dstCapacity);
            uint8_t* base;
            int64_t pos = total;
            int anyGpu = 0, rc = LZ4B200_OK;
            if (!direct && !bounce) {
                bounce = (uint8_t*)malloc((size_t)(G * bs));
                if (!bounce) { free(tmpSz); result = LZ4B200_ERR_ARG; goto done; }
            }
            base = direct ? dst + total : bounce;
            /* stored-raw blocks need no GPU: they enter the batch with size 0 (rejected there) and are copied here */
            for (k = 0; k < cnt; k++) { tmpSz[k] = raw[g0 + k] ? 0 : csz[g0 + k]; anyGpu |= !raw[g0 + k]; }
            if (anyGpu) rc = LZ4B200_decompress_blocks_host(src, off + g0, tmpSz, base, bs, (int32_t)bs, rets + g0, cnt);
            if (rc != LZ4B200_OK) { free(bounce); free(tmpSz); result = rc; goto done; }
            for (k = 0; k < cnt; k++) {
                int64_t r;
                if (raw[g0 + k]) { memcpy(base + k * bs, src + off[g0 + k], (size_t)csz[g0 + k]); rets[g0 + k] = csz[g0 + k]; }
                r = rets[g0 + k];
                if (r < 0) { free(bounce); free(tmpSz); goto done; }      /* a block failed to decode */
                if (pos + r > dstCapacity || (contentSize >= 0 && pos + r > contentSize)) {
                    free(bounce); free(tmpSz);
                    result = (pos + r > dstCapacity) ? LZ4B200_ERR_DSTSIZE : LZ4B200_ERR_FRAME;
                    goto done;
                }
                if (dst + pos != base + k * bs) memmove(dst + pos, base + k * bs, (size_t)r);      /* close the gap (moves left) */
                pos += r;
            }
            total = pos;
        }
        free(bounce); free(tmpSz);
    }
    if (contentSize >= 0 && contentSize != total) goto done;              /* lz4frame.c: frameSize_wrong */
    result = total;
done:
    free(off); free(csz); free(raw); free(rets);
    return result;
}

/*
 * oracle/cpu_harness.c -- TEST INFRASTRUCTURE ONLY (see lz4_oracle.h).
 *
 * Times a CPU LZ4 block codec (the compiled reference from oracle/_ref, or the oracle port)
 * over a stream of independent blocks, the way programs/bench.c:464-555 loops the reference
 * (one call per block), with the blocks statically partitioned over T pthreads
 * (BASELINE.md section 3).  The codec is passed as a function pointer so this file links
 * against neither implementation.
 */
#define _GNU_SOURCE
#include "lz4_oracle.h"
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    int kind;   /* 0 decompress, 1 compress, 2 datagen, 3 first-touch of dst, 4 pack slots into a contiguous stream */
    int cpu;    /* CPU this worker pins itself to, or -1 */
    const int64_t* packOff;   /* kind 4: destination offset of block i */
    oracle_decomp_fn dfn;
    oracle_comp_fn cfn;
    const uint8_t* src;
    const int64_t* srcOff;
    const int32_t* srcSize;
    int64_t srcStride;
    int32_t blockSize;
    int64_t lastSize;
    uint8_t* dst;
    int64_t dstStride;
    int32_t dstCap;
    int accel;
    int32_t* outSizes;
    int64_t begin, end, nBlocks;
    int failed;
    /* datagen */
    size_t segBytes, total;
    double proba;
    unsigned seed0;
} job_t;

static void* worker(void* arg)
{
    job_t* j = (job_t*)arg;
    int64_t i;
    if (j->cpu >= 0) {           /* one worker per allowed CPU, the same CPU in every pass: memory it touched first stays local */
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(j->cpu, &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    for (i = j->begin; i < j->end; i++) {
        if (j->kind == 3) {
            memset(j->dst + i * j->dstStride, 0, (size_t)j->dstStride);
        } else if (j->kind == 4) {
            memcpy(j->dst + j->packOff[i], j->src + i * j->srcStride, (size_t)j->srcSize[i]);
        } else if (j->kind == 0) {
            int r = j->dfn((const char*)j->src + j->srcOff[i], (char*)j->dst + i * j->dstStride, j->srcSize[i], j->dstCap);
            j->outSizes[i] = r;
            if (r < 0) j->failed = 1;
        } else if (j->kind == 1) {
            int sz = (i == j->nBlocks - 1) ? (int)j->lastSize : j->blockSize;
            int r = j->cfn((const char*)j->src + i * j->srcStride, (char*)j->dst + i * j->dstStride, sz, j->dstCap, j->accel);
            j->outSizes[i] = r;
            if (r <= 0) j->failed = 1;
        } else {
            size_t off = (size_t)i * j->segBytes;
            size_t len = j->total - off < j->segBytes ? j->total - off : j->segBytes;
            oracle_datagen(j->dst + off, len, j->proba, 0.0, j->seed0 + (unsigned)i);
        }
    }
    return NULL;
}

static double run_jobs(job_t* proto, int64_t nUnits, int threads)
{
    pthread_t* th;
    job_t* jobs;
    int t, failed = 0;
    double t0, t1;
    if (threads < 1) threads = 1;
    if ((int64_t)threads > nUnits) threads = (int)(nUnits > 0 ? nUnits : 1);
    th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    jobs = (job_t*)malloc(sizeof(job_t) * (size_t)threads);
    {   /* worker t runs on the t-th CPU of the calling thread's affinity mask (static partition, programs/bench.c has one thread) */
        cpu_set_t allowed;
        int ncpu = 0, c, k;
        int* list = (int*)malloc(sizeof(int) * CPU_SETSIZE);
        if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
            for (c = 0; c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &allowed)) list[ncpu++] = c;
        /* pin only when the pass uses every allowed CPU (the all-threads reference arm); partial passes -- one rank of a
         * multi-process run generating its input -- stay unpinned so that processes do not pile onto the same CPUs */
        for (k = 0; k < threads; k++) jobs[k].cpu = (ncpu > 1 && threads >= ncpu) ? list[k % ncpu] : -1;
        free(list);
    }
    for (t = 0; t < threads; t++) {
        const int cpu = jobs[t].cpu;
        jobs[t] = *proto;
        jobs[t].cpu = cpu;
        jobs[t].begin = nUnits * t / threads;
        jobs[t].end = nUnits * (t + 1) / threads;
        jobs[t].failed = 0;
    }
    t0 = now_s();
    /* every worker is a created thread: the caller's own affinity mask is never narrowed */
    for (t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &jobs[t]);
    for (t = 0; t < threads; t++) pthread_join(th[t], NULL);
    t1 = now_s();
    for (t = 0; t < threads; t++) failed |= jobs[t].failed;
    free(th); free(jobs);
    return failed ? -1.0 : (t1 - t0);
}

double oracle_time_decompress(oracle_decomp_fn fn, const uint8_t* src, const int64_t* srcOff,
                              const int32_t* srcSize, uint8_t* dst, int64_t dstStride, int32_t dstCap,
                              int32_t* outSizes, int64_t nBlocks, int threads)
{
    job_t j = {0};
    j.kind = 0; j.dfn = fn; j.src = src; j.srcOff = srcOff; j.srcSize = srcSize;
    j.dst = dst; j.dstStride = dstStride; j.dstCap = dstCap; j.outSizes = outSizes; j.nBlocks = nBlocks;
    return run_jobs(&j, nBlocks, threads);
}

double oracle_time_compress(oracle_comp_fn fn, const uint8_t* src, int64_t srcStride, int32_t srcSize,
                            int64_t lastSize, uint8_t* dst, int64_t dstStride, int32_t dstCap, int accel,
                            int32_t* outSizes, int64_t nBlocks, int threads)
{
    job_t j = {0};
    j.kind = 1; j.cfn = fn; j.src = src; j.srcStride = srcStride; j.blockSize = srcSize; j.lastSize = lastSize;
    j.dst = dst; j.dstStride = dstStride; j.dstCap = dstCap; j.accel = accel; j.outSizes = outSizes; j.nBlocks = nBlocks;
    return run_jobs(&j, nBlocks, threads);
}

/* first-touch `dst` (nBlocks x dstStride bytes) with the partition / pinning of the timed passes */
void oracle_first_touch(uint8_t* dst, int64_t dstStride, int64_t nBlocks, int threads)
{
    job_t j = {0};
    j.kind = 3; j.dst = dst; j.dstStride = dstStride;
    run_jobs(&j, nBlocks, threads);
}

/* copy block i (srcSize[i] bytes at slots + i*slotStride) to packed + packOff[i], worker-partitioned like the timed passes */
void oracle_pack(const uint8_t* slots, int64_t slotStride, const int32_t* sizes, const int64_t* packOff,
                 uint8_t* packed, int64_t nBlocks, int threads)
{
    job_t j = {0};
    j.kind = 4; j.src = slots; j.srcStride = slotStride; j.srcSize = sizes; j.packOff = packOff; j.dst = packed;
    run_jobs(&j, nBlocks, threads);
}

void oracle_datagen_mt(uint8_t* buffer, size_t size, size_t segBytes, double matchProba, unsigned seed0, int threads)
{
    job_t j = {0};
    int64_t nSeg = (int64_t)((size + segBytes - 1) / segBytes);
    j.kind = 2; j.dst = buffer; j.segBytes = segBytes; j.total = size; j.proba = matchProba; j.seed0 = seed0;
    run_jobs(&j, nSeg, threads);
}

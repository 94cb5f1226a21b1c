/*
 * lz4_kernels.h -- the thin extern "C" FFI between the C host layer (lz4_api.c) and the CUDA
 * kernels (lz4_kernels.cu).  Launchers only enqueue work on `stream`; they return a cudaError_t
 * value as int.
 */
#ifndef LZ4_KERNELS_H
#define LZ4_KERNELS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    const uint8_t* src;        /* base of compressed bytes */
    const int64_t* srcOff;     /* [n] */
    const int32_t* srcSize;    /* [n] */
    uint8_t* dst;
    const int64_t* dstOff;     /* [n] or NULL -> i*dstStride */
    int64_t dstStride;
    const int32_t* dstCapArr;  /* [n] or NULL -> dstCap */
    int32_t dstCap;
    int32_t* outSize;          /* [n] */
    int64_t nBlocks;
    void* workspace;
    size_t workspaceBytes;
} lz4k_decode_args;

typedef struct {
    const uint8_t* src;
    int64_t srcStride;
    const int32_t* srcSizeArr; /* [n] or NULL -> srcSize */
    int32_t srcSize;
    uint8_t* dst;
    int64_t dstStride;
    int32_t dstCap;
    int32_t acceleration;
    int32_t* outSize;
    int64_t nBlocks;
} lz4k_encode_args;

size_t lz4k_decode_workspace_bytes(int64_t nBlocks);   /* any capacities */
/* tighter: per-block capacity array (perBlockCaps != 0) or one capacity for every block */
size_t lz4k_decode_workspace_bytes_for(int64_t nBlocks, int perBlockCaps, int32_t dstCap);
size_t lz4k_decode_workspace_bytes_min(int64_t nBlocks, int perBlockCaps, int32_t dstCap);   /* without the wide marks of the tiles kernel */
/* phases: bit 0 = scan (validate, sizes), bit 1 = expand (move bytes; needs a prior scan's outSize) */
int lz4k_launch_decode(const lz4k_decode_args* a, int phases, void* stream);
int lz4k_launch_encode(const lz4k_encode_args* a, void* stream);       /* byte-identical to LZ4_compress_fast */
int lz4k_launch_encode_par(const lz4k_encode_args* a, void* stream);   /* parallel parse; blocks of <= 65 536 bytes */
int lz4k_launch_pack(const uint8_t* slots, int64_t slotStride, const int32_t* sizes, int64_t nBlocks,
                     uint8_t* packed, int64_t* outOff, int headerBytes, void* stream);
/* the body of an LZ4 frame: [LE32 block header][payload] per block; blocks whose compression did not gain are stored raw */
int lz4k_launch_pack_frame(const uint8_t* slots, int64_t slotStride, const int32_t* sizes, const uint8_t* src, int64_t srcStride,
                           int32_t blockSize, int32_t lastSize, int64_t nBlocks, uint8_t* packed, int64_t* outOff, void* stream);
uint64_t lz4k_launch_count(void);
int lz4k_debug_poison_smem(uint32_t pattern, int lo, int hi, void* stream);   /* developer tool */
int lz4k_launch_ceiling(const lz4k_decode_args* a, int mode, void* stream);   /* developer tool: skeleton of the rows kernel without the decode */
int lz4k_debug_phase_cycles(unsigned long long* out8);

#ifdef __cplusplus
}
#endif
#endif

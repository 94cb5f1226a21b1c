/*
 * lz4_b200.h -- C ABI of liblz4_b200.so: a B200 (sm_100a) LZ4 block codec.
 *
 * Two layers, both plain C (pointers and sizes only, no torch / C++ types):
 *
 *  1. DROP-IN layer: the one-shot block API of the reference's lib/lz4.h, same names, argument
 *     meaning, return values and error behaviour, host pointers, synchronous.  A program built
 *     against the reference's own lz4.h links against this library unchanged for these symbols
 *     (each declaration cites the lz4.h / lz4.c line it replaces; v1.10.0).
 *
 *  2. BATCH layer (additive, prefix LZ4B200_): many independent blocks per call, device pointers,
 *     asynchronous on a CUDA stream -- the measured path.  `stream` is a
 *     cudaStream_t passed as void* so that this header needs no CUDA include.
 *
 * There is no CPU fallback: without a usable CUDA device every codec entry point fails
 * (compress -> 0, decompress -> negative, batch -> LZ4B200_ERR_CUDA).
 *
 * Threads and devices: every entry point is re-entrant (the host-pointer calls share one context
 * behind a mutex).  The host-pointer calls (layers 1, 3, 4) run on the CUDA device that is current
 * in the calling thread at their FIRST use in the process; the device-pointer calls (layer 2) run
 * on the current device / the given stream, and calls that overlap in time must use different
 * workspaces.
 */
#ifndef LZ4_B200_H
#define LZ4_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#  define LZ4B200_API __attribute__((visibility("default")))
#else
#  define LZ4B200_API
#endif

/* ---------------------------------------------------------------------------------------------
 * 1. Drop-in layer  (replaces lib/lz4.c behind lib/lz4.h)
 * ------------------------------------------------------------------------------------------- */

#define LZ4B200_VERSION_NUMBER 11000          /* == LZ4_VERSION_NUMBER, lz4.h:131-135 */
#define LZ4B200_MAX_INPUT_SIZE 0x7E000000     /* == LZ4_MAX_INPUT_SIZE, lz4.h:214 */
#define LZ4B200_STATE_BYTES    16416          /* == sizeof(LZ4_stream_t), lz4.h:729-733 */

/* lz4.h:142 / lz4.c:749 */
LZ4B200_API int LZ4_versionNumber(void);
/* lz4.h:143 / lz4.c:750 */
LZ4B200_API const char* LZ4_versionString(void);
/* lz4.h:226 / lz4.c:751: srcSize + srcSize/255 + 16, or 0 when srcSize > LZ4_MAX_INPUT_SIZE */
LZ4B200_API int LZ4_compressBound(int inputSize);
/* lz4.h:245 / lz4.c:752 */
LZ4B200_API int LZ4_sizeofState(void);

/* lz4.h:191 / lz4.c:1472.  Returns the number of bytes written into dst, 0 on failure (dst too
 * small or srcSize > LZ4_MAX_INPUT_SIZE).  Output is byte-identical to the reference's. */
LZ4B200_API int LZ4_compress_default(const char* src, char* dst, int srcSize, int dstCapacity);
/* lz4.h:236 / lz4.c:1453.  acceleration <= 0 -> 1; > 65537 -> 65537 (lz4.c:1386-1387). */
LZ4B200_API int LZ4_compress_fast(const char* src, char* dst, int srcSize, int dstCapacity, int acceleration);
/* lz4.h:246 / lz4.c:1382.  `state` is validated the way LZ4_initStream does (lz4.c:1552-1560:
 * non-NULL, 8-byte aligned) but not otherwise used: the match table lives in GPU shared memory.
 * Returns 0 for an invalid state (the reference dereferences it). */
LZ4B200_API int LZ4_compress_fast_extState(void* state, const char* src, char* dst, int srcSize, int dstCapacity, int acceleration);
/* lz4.h:611 / lz4.c:1414 -- the symbol lz4frame dispatches to (lz4frame.c:919).  Always starts
 * from a clean table, i.e. behaves like LZ4_compress_fast_extState. */
LZ4B200_API int LZ4_compress_fast_extState_fastReset(void* state, const char* src, char* dst, int srcSize, int dstCapacity, int acceleration);

/* lz4.h:208 / lz4.c:2451.  Returns the number of bytes decoded into dst (<= dstCapacity), or a
 * negative value for malformed input / too small dst -- the same value -(position)-1 as the
 * reference's x86-64 build (lz4.c:2443).  Never reads outside src[0,compressedSize) nor writes
 * outside dst[0,dstCapacity). */
LZ4B200_API int LZ4_decompress_safe(const char* src, char* dst, int compressedSize, int dstCapacity);
/* lz4.h:546 / lz4.c:2717.  The call lz4frame makes per block (lz4frame.c:1901,1946).  dictSize == 0
 * routes to LZ4_decompress_safe exactly as lz4.c:2721-2722 does; a non-empty dictionary is out
 * of scope for the GPU path and returns -1. */
LZ4B200_API int LZ4_decompress_safe_usingDict(const char* src, char* dst, int compressedSize, int dstCapacity,
                                              const char* dictStart, int dictSize);

/* ---------------------------------------------------------------------------------------------
 * 2. Batch layer  (device pointers, asynchronous)
 * ------------------------------------------------------------------------------------------- */

#define LZ4B200_OK            0
#define LZ4B200_ERR_ARG      (-1)   /* bad argument (NULL pointer, negative count, workspace too small) */
#define LZ4B200_ERR_CUDA     (-2)   /* CUDA runtime error; see LZ4B200_last_cuda_error() */

/* Library / device introspection. */
LZ4B200_API int LZ4B200_device_count(void);
LZ4B200_API const char* LZ4B200_last_cuda_error(void);
/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
LZ4B200_API uint64_t LZ4B200_launch_count(void);

/* Bytes of device workspace LZ4B200_decompress_blocks needs for nBlocks blocks of ANY capacity
 * (worst case: 32 KB of sequence marks per block). */
LZ4B200_API size_t LZ4B200_decompress_workspace_bytes(int64_t nBlocks);
/* The workspace one call wants: perBlockCaps != 0 when d_dstCap is passed, else the uniform dstCap.
 * Marks are sized by the capacity (dstCap/4 + 2 slots, at most 8192), so batches of small blocks need little
 * workspace (<= LZ4B200_decompress_workspace_bytes(nBlocks)).  A batch of blocks ABOVE 64 KB with a uniform capacity
 * of at most 64 MB (lz4frame's 256 KB .. 4 MB blocks) is decoded in 60 KB output tiles and wants two words per possible
 * sequence, i.e. about 2 x dstCap bytes per block; with a smaller workspace (at least
 * LZ4B200_decompress_workspace_bytes(nBlocks)) such a batch is still decoded, by the slow one-warp-per-block kernel. */
LZ4B200_API size_t LZ4B200_decompress_workspace_bytes_for(int64_t nBlocks, int perBlockCaps, int32_t dstCap);

/*
 * Decompress nBlocks independent LZ4 blocks (the batched LZ4_decompress_safe; replaces the serial
 * per-block loops of bench.c:514-555 and lz4frame.c:1901).
 *   block i input : d_src + d_srcOff[i], d_srcSize[i] bytes
 *   block i output: d_dst + (d_dstOff ? d_dstOff[i] : i*dstStride), capacity (d_dstCap ? d_dstCap[i] : dstCap)
 *   d_outSize[i]  : LZ4_decompress_safe's return value for block i (negative = that block is
 *                   malformed; other blocks are unaffected)
 * Output regions of different blocks must not overlap.  Enqueued on `stream`; returns LZ4B200_OK
 * once enqueued.
 * Readable padding: blocks of up to 65 535 bytes are staged by 16-byte-granular bulk loads, so the 16-byte
 * granules that hold the first and the last byte of every block are read in full -- up to 15 bytes before
 * d_src + d_srcOff[i] and after the block's end must be readable device memory (e.g. 16 bytes of slack at both
 * ends of d_src; bytes of neighbouring blocks are fine).  Nothing outside a block influences its result.
 */
LZ4B200_API int LZ4B200_decompress_blocks(const void* d_src, const int64_t* d_srcOff, const int32_t* d_srcSize,
                                          void* d_dst, const int64_t* d_dstOff, int64_t dstStride,
                                          const int32_t* d_dstCap, int32_t dstCap,
                                          int32_t* d_outSize, int64_t nBlocks,
                                          void* d_workspace, size_t workspaceBytes, void* stream);

/*
 * The same call split in its two phases (phases: 1 = scan only -- validate every block and fill
 * d_outSize with the decoded sizes / error codes without moving data, the batched equivalent of
 * asking "what would LZ4_decompress_safe return"; 2 = expand only -- move the bytes of the
 * blocks a previous scan accepted, same arguments and workspace; 3 = both).  An expand consumes what
 * its scan left in the workspace (sizes, sequence marks, the list of blocks for the generic kernel): one
 * expand per scan.
 */
LZ4B200_API int LZ4B200_decompress_blocks_phased(const void* d_src, const int64_t* d_srcOff, const int32_t* d_srcSize,
                                                 void* d_dst, const int64_t* d_dstOff, int64_t dstStride,
                                                 const int32_t* d_dstCap, int32_t dstCap,
                                                 int32_t* d_outSize, int64_t nBlocks,
                                                 void* d_workspace, size_t workspaceBytes, int phases, void* stream);

/*
 * Compress nBlocks independent blocks (the batched LZ4_compress_fast; replaces bench.c:464-493 and
 * lz4frame.c:1046-1055).
 *   block i input : d_src + i*srcStride, size (d_srcSize ? d_srcSize[i] : srcSize)
 *   block i output: d_dst + i*dstStride, capacity dstCap  (use LZ4_compressBound(srcSize) for the
 *                   reference's notLimited behaviour)
 *   d_outSize[i]  : LZ4_compress_fast's return value for block i (0 = did not fit)
 * Each block's bytes are identical to the reference's LZ4_compress_fast output.
 */
LZ4B200_API int LZ4B200_compress_blocks(const void* d_src, int64_t srcStride, const int32_t* d_srcSize, int32_t srcSize,
                                        void* d_dst, int64_t dstStride, int32_t dstCap, int acceleration,
                                        int32_t* d_outSize, int64_t nBlocks, void* stream);

/*
 * The same call with the PARALLEL PARSE (throughput mode): every block is a valid LZ4 block that
 * LZ4_decompress_safe expands to the input, the result is deterministic, but it is NOT byte-identical to
 * LZ4_compress_fast: every position is a match candidate (no skipping), matches are chosen by the same greedy
 * rule, and the compression ratio is within 2 % of the reference's at acceleration 1 (usually above it).
 * acceleration 1..4 probe every position, 5..12 every 2nd, 13..20 every 3rd, ... (faster, lower ratio -- monotone
 * like the reference's parameter, not the same curve).  Blocks above 65 536 bytes are compressed by the
 * byte-identical encoder.  d_outSize[i] == 0: block i did not fit dstCap (as LZ4_compress_fast).
 */
LZ4B200_API int LZ4B200_compress_blocks_parallel(const void* d_src, int64_t srcStride, const int32_t* d_srcSize, int32_t srcSize,
                                        void* d_dst, int64_t dstStride, int32_t dstCap, int acceleration,
                                        int32_t* d_outSize, int64_t nBlocks, void* stream);

/*
 * Pack per-block slots into one contiguous stream (what the CPU does implicitly by writing blocks
 * back to back, lz4frame.c:1046-1055): d_outOff[i] = sum_{k<i} max(d_sizes[k],0) + i*headerBytes,
 * d_outOff[nBlocks] = total; block i's bytes are copied to d_packed + d_outOff[i] + headerBytes.
 * With headerBytes == 4 a little-endian block header (the size) is written before each block, as
 * LZ4F_makeBlock does (lz4frame.c:896-907); use 0 for a bare stream.
 */
LZ4B200_API int LZ4B200_pack_blocks(const void* d_slots, int64_t slotStride, const int32_t* d_sizes, int64_t nBlocks,
                                    void* d_packed, int64_t* d_outOff, int headerBytes, void* stream);

/*
 * The body of an LZ4 frame of independent blocks, assembled on the device (LZ4F_makeBlock, lz4frame.c:883-908):
 * for every block [LE32 header][payload]; a block whose compressed size d_sizes[i] is 0 (it did not fit size - 1) or
 * not smaller than its source is stored raw (payload = its source bytes from d_src + i*srcStride, bit 31 of the
 * header set).  Block i has blockSize source bytes, the last one lastSize if lastSize > 0.  d_outOff[nBlocks] = body bytes.
 */
LZ4B200_API int LZ4B200_pack_frame_blocks(const void* d_slots, int64_t slotStride, const int32_t* d_sizes,
                                          const void* d_src, int64_t srcStride, int32_t blockSize, int32_t lastSize,
                                          int64_t nBlocks, void* d_body, int64_t* d_outOff, void* stream);

/*
 * Multi-GPU reassembly (the ordered-writer role of programs/lz4io.c:594-635 when every GPU decodes a shard of a frame):
 * enqueue on `stream` the copy of `bytes` device bytes from d_src (current device) to d_dstPeer, memory of device
 * `peerDevice` that is mapped into this process (cudaIpcOpenMemHandle / a same-process allocation).  Peer access is
 * enabled on first use (a call with bytes == 0 only does that: a probe); the copy runs on a copy engine over NVLink and
 * takes no SM from the codec kernels.
 * Returns LZ4B200_OK, LZ4B200_ERR_ARG (the devices cannot reach each other directly) or LZ4B200_ERR_CUDA.
 */
LZ4B200_API int LZ4B200_peer_copy_async(void* d_dstPeer, int peerDevice, const void* d_src, size_t bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * 3. Host-buffer batch calls (synchronous): the batch layer with the host<->device copies inside,
 *    pipelined over chunks.  h_* are host pointers (pinned memory gives full PCIe speed).
 * ------------------------------------------------------------------------------------------- */
LZ4B200_API int LZ4B200_decompress_blocks_host(const void* h_src, const int64_t* h_srcOff, const int32_t* h_srcSize,
                                               void* h_dst, int64_t dstStride, int32_t dstCap,
                                               int32_t* h_outSize, int64_t nBlocks);
LZ4B200_API int LZ4B200_compress_blocks_host(const void* h_src, int64_t srcStride, int32_t srcSize, int64_t lastSize,
                                             void* h_dst, int64_t dstStride, int32_t dstCap, int acceleration,
                                             int32_t* h_outSize, int64_t nBlocks);

/* ---------------------------------------------------------------------------------------------
 * 4. Frame layer: one-shot LZ4 frames of INDEPENDENT blocks, host buffers.
 *    The container logic (header, block headers, EndMark) runs on the host, every block goes through
 *    the batch layer in one pipeline instead of lz4frame.c's serial per-block loop
 *    (lz4frame.c:1046-1055 / :1839-1919).
 * ------------------------------------------------------------------------------------------- */
#define LZ4B200_ERR_FRAME        (-3)   /* malformed frame (bad magic / header checksum / block size / truncated) */
#define LZ4B200_ERR_UNSUPPORTED  (-4)   /* valid LZ4 frame feature outside this layer: linked blocks, block or
                                           content checksums, dictID, skippable frames, HC levels */
#define LZ4B200_ERR_DSTSIZE      (-5)   /* destination buffer too small */

/* Upper bound of the frame LZ4B200_compressFrame_host can produce (>= LZ4F_compressFrameBound, lz4frame.c:406). */
LZ4B200_API int64_t LZ4B200_compressFrameBound(int64_t srcSize, int blockSizeID);

/*
 * LZ4F_compressFrame (lz4frame.h:224 / lz4frame.c:484) for preferences
 *   { frameInfo = { blockSizeID, LZ4F_blockIndependent, noContentChecksum, LZ4F_frame,
 *                   contentSize = contentSizeFlag ? srcSize : 0, dictID 0, noBlockChecksum },
 *     compressionLevel <= 1 (0/1: acceleration 1; negative: acceleration -level+1, lz4frame.c:913) }.
 * blockSizeID: 0 (default = 4) or 4..7 = 64 KB / 256 KB / 1 MB / 4 MB; shrunk for small inputs exactly like
 * LZ4F_optimalBSID (lz4frame.c:362-373).  The frame is byte-identical to the reference's.
 * Returns the frame size, or a negative LZ4B200_ERR_* value.
 */
LZ4B200_API int64_t LZ4B200_compressFrame_host(const void* h_src, int64_t srcSize, void* h_dst, int64_t dstCapacity,
                                               int blockSizeID, int compressionLevel, int contentSizeFlag);

/*
 * Decode ONE whole frame (the one-shot use of LZ4F_decompress, lz4frame.c:1613): header checks as
 * LZ4F_decodeHeader (lz4frame.c:1340-1425), then all blocks in one batch (stored-raw blocks are copied).
 * Returns the number of decoded bytes, or a negative LZ4B200_ERR_* value (LZ4B200_ERR_FRAME also when a
 * block fails to decode or the content size field disagrees).  *consumed (optional) = frame length.
 */
LZ4B200_API int64_t LZ4B200_decompressFrame_host(const void* h_src, int64_t srcSize, void* h_dst, int64_t dstCapacity,
                                                 int64_t* consumed);

#ifdef __cplusplus
}
#endif
#endif /* LZ4_B200_H */

// Host build of the decoder's scan (lz4_b200/csrc/lz4_scan_core.h): TEST INFRASTRUCTURE.
// The same text the scan kernel compiles for the device, compiled by g++ so that its logic -- the
// exact return value of LZ4_decompress_safe and the per-sequence marks -- can be checked against the
// oracle and the golden vectors without a GPU (tests/test_scan_core_host.py).
#include <stdint.h>
#include "../../lz4_b200/csrc/lz4_scan_core.h"

// src must be readable from the 4-byte aligned address at or below src up to the aligned word that
// holds src[n-1] (the device reads whole aligned words too).
extern "C" int scan_host(const uint8_t* src, int n, int cap, uint32_t* nSeqOut, uint32_t* marks)
{
    *nSeqOut = 0;
    MemPtr<true> mem{src};
    return scan_block(mem, n, cap, nSeqOut, marks, (uint32_t)kMaxSeqFast);
}

// the same scan through the per-thread shared-memory ring (MemRing): the host build of the ring queues its
// asynchronous copies and poisons their destination until the matching wait, so a read that the device could
// see before the data arrived changes the result here.  waits[0] receives the number of explicit waits.
extern "C" int scan_host_ring(const uint8_t* src, int n, int cap, uint32_t* nSeqOut, uint32_t* marks)
{
    *nSeqOut = 0;
    if (n <= 0) { MemPtr<true> m0{src}; return scan_block(m0, n, cap, nSeqOut, marks, (uint32_t)kMaxSeqFast); }
    alignas(16) static uint8_t ring[kRingBytes];
    for (int k = 0; k < kRingBytes; k++) ring[k] = 0x5C;
    MemRing mem;
    mem.init(src, n, ring);
    const int r = scan_block(mem, n, cap, nSeqOut, marks, (uint32_t)kMaxSeqFast);
    mem.cp.wait(0);
    return r;
}
extern "C" int scan_host_max_seq(void) { return kMaxSeqFast; }

// CPU emulator of the warp-per-block scan (lz4_b200/csrc/lz4_scan_v2.h): TEST INFRASTRUCTURE.
// Runs the 32 lanes of a warp phase by phase (the device puts a __syncwarp between phases) on the
// host build of the same header and returns what the kernel would store for the block.
#include <stdint.h>
#include <string.h>
#include "../../lz4_b200/csrc/lz4_scan_v2.h"

// stats[0] = fix-up rounds, stats[1] = lane walks in the fix-up rounds, stats[2] = lane that finished the block
extern "C" int scan_v2_host(const uint8_t* src, int n, int cap, uint32_t* nSeqOut, uint32_t* marks, int* stats)
{
    *nSeqOut = 0;
    stats[0] = stats[1] = 0; stats[2] = -1;
    if (cap < 64 || n < kSv2MinBytes) return scan_block(src, n, cap, nSeqOut, marks);   // lane 0, one-thread scan
    SV2Shared S;
    SV2Lane L[32];
    memset(&S, 0, sizeof(S));
    for (int l = 0; l < 32; l++) sv2_phase0(l, L[l], S, src, n, cap);
    for (;;) {
        for (int l = 0; l < 32; l++) sv2_decide(l, L[l], S);
        S.changed = 0;
        for (int l = 0; l < 32; l++) { if (L[l].need && !L[l].newVoid) stats[1]++; sv2_redo(l, L[l], S, src, n, cap); }
        if (!S.changed) break;
        stats[0]++;
        if (stats[0] > 64) return -1000000;                    // must converge within 32 rounds
    }
    for (int l = 0; l < 32; l++) sv2_write(l, L[l], S, src, n, cap, marks);
    S.ret = -2000000;
    for (int l = 0; l < 32; l++) sv2_finish(l, S, src, n, cap, marks);
    for (int l = 0; l < 32; l++) if (S.end[l].kind != SV2_RAN) { stats[2] = l; break; }
    *nSeqOut = S.nseq;
    return S.ret;
}

// CPU emulator of phase B (uniform-body variant, lz4_b200/csrc/lz4_phaseb_v2.h): TEST INFRASTRUCTURE.
// Replays the kernel's loop -- 32 warps x 32 lanes, ballots, dynamic hand-out, done flags -- one
// iteration per warp per tick, with the warp order and the lane order inside a warp shuffled every
// tick (a lane never relies on another lane of the same iteration), and compares the assembled
// output window with the expected decoded bytes.  Built by tests/test_phaseb_v2_emul.py (g++).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>

#include "../../lz4_b200/csrc/lz4_phaseb_v2.h"

namespace {
struct Rng { uint64_t s; uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); } };
constexpr int kIn = 65536 + 64, kOutDelta = kIn + 16;     // the layout of FastSmem: in[], outPad[16], out[]
}

// stats[0] warp-iterations, [1] lane-iterations with a piece, [2] blocked, [3] first differing byte or -1
extern "C" int pb_emulate(const uint8_t* comp, int n, const uint8_t* expect, int total, int head, uint32_t seed, long long* stats)
{
    if (n > 65535 || total > 65536 || head < 0 || head > 15) return -2;
    std::vector<uint8_t> window(16 + kOutDelta + 65536 + 16 + 16, 0xEE);
    uint8_t* in = window.data() + 16;                        // 16 bytes of pad in front (inPad)
    while (((uintptr_t)in) & 15) in++;
    uint8_t* out = in + kOutDelta;
    memcpy(in + head, comp, (size_t)n);
    // ---- phase A on the host: records, start bits, ranks (same encoding as the kernel) ----
    std::vector<pb_rec> rec;
    std::vector<uint32_t> bits(2048, 0);
    std::vector<uint16_t> seqbase(2048, 0);
    {
        int p = 0, op = 0;
        std::vector<int> ops, lls, lss, offs;
        for (;;) {
            const int tok = comp[p++];
            int ll = tok >> 4;
            if (ll == 15) { int x; do { x = comp[p++]; ll += x; } while (x == 255); }
            ops.push_back(op); lls.push_back(ll); lss.push_back(p);
            p += ll; op += ll;
            if (p >= n) { offs.push_back(0); break; }
            const int off = comp[p] | (comp[p + 1] << 8);
            p += 2;
            int ml = (tok & 15) + 4;
            if ((tok & 15) == 15) { int x; do { x = comp[p++]; ml += x; } while (x == 255); }
            offs.push_back(off);
            op += ml;
        }
        if (op != total) return -3;
        const int nseq = (int)ops.size();
        if (nseq > 8192) return -4;
        rec.resize((size_t)nseq + 2);
        for (int k = 0; k < nseq; k++) {
            const int nxt = (k + 1 == nseq) ? total : ops[k + 1];
            rec[k].x = ((uint32_t)(ops[k] + lls[k]) & 0xFFFFu) | ((uint32_t)nxt << 16);
            rec[k].y = ((uint32_t)(lss[k] - ops[k]) & 0xFFFFu) | ((uint32_t)offs[k] << 16);
            if (ops[k] < total) bits[ops[k] >> 5] |= 1u << (ops[k] & 31);
        }
        uint32_t run = 0;
        for (int i = 0; i < 2048; i++) { seqbase[i] = (uint16_t)run; run += (uint32_t)__builtin_popcount(bits[i]); }
    }
    std::vector<uint8_t> done8(8192 + 16, 0);
    done8[kPbSentinel] = 1;

    PBView V;
    V.window = in; V.out = out; V.outDelta = kOutDelta;
    V.rec = rec.data(); V.bits = bits.data(); V.seqbase = seqbase.data(); V.done8 = done8.data();
    V.head = head; V.total = total;

    const int nwarps = 32;
    std::vector<PBLane> lanes(nwarps * 32);
    for (auto& L : lanes) pb_init(L);
    std::vector<int> warpNext(nwarps, 0), warpChunks(nwarps, 0), finished(nwarps, 0);
    const int nstrips = (total + 255) >> 8;
    for (int w = 0; w < nwarps; w++) warpChunks[w] = (w < nstrips) ? (((nstrips - 1 - w) >> 5) + 1) << 5 : 0;
    Rng rng{seed * 2654435761ull + 12345};
    long long iters = 0, laneIters = 0, blocked = 0, guard = 0;
    int live = nwarps;
    std::vector<int> worder(nwarps), lorder(32);
    for (int i = 0; i < nwarps; i++) worder[i] = i;
    for (int i = 0; i < 32; i++) lorder[i] = i;
    while (live > 0) {
        if (++guard > 4000000) return -5;                    // livelock guard
        for (int i = nwarps - 1; i > 0; i--) std::swap(worder[i], worder[rng.next() % (uint32_t)(i + 1)]);
        for (int wi = 0; wi < nwarps; wi++) {
            const int w = worder[wi];
            if (finished[w]) continue;
            if (seed && (rng.next() & 3) == 0) continue;     // this warp is not scheduled in this tick
            PBLane* L = &lanes[w * 32];
            unsigned want = 0;
            for (int l = 0; l < 32; l++) if (L[l].needNew && !L[l].exhausted) want |= 1u << l;
            if (want) {
                for (int l = 0; l < 32; l++) {
                    if (!(want >> l & 1)) continue;
                    const int c = warpNext[w] + __builtin_popcount(want & ((1u << l) - 1u));
                    pb_take(L[l], V, w, c, warpChunks[w]);
                }
                warpNext[w] += __builtin_popcount(want);
            }
            unsigned act = 0, notEx = 0;
            for (int l = 0; l < 32; l++) { if (!L[l].needNew) act |= 1u << l; if (!L[l].exhausted) notEx |= 1u << l; }
            if (!act) { if (!notEx) { finished[w] = 1; live--; } continue; }
            iters++;
            laneIters += __builtin_popcount(act);
            for (int i = 31; i > 0; i--) std::swap(lorder[i], lorder[rng.next() % (uint32_t)(i + 1)]);
            for (int li = 0; li < 32; li++) {
                const int l = lorder[li];
                if (L[l].needNew) continue;
                if (!pb_body(L[l], V)) blocked++;
            }
        }
    }
    stats[0] = iters; stats[1] = laneIters; stats[2] = blocked; stats[3] = -1;
    for (int i = 0; i < total; i++) if (out[i] != expect[i]) { stats[3] = i; return 1; }
    // every chunk must have been flagged, nothing outside the window touched
    for (int c = 0; c < (total + 7) / 8; c++) if (!done8[c]) { stats[3] = c * 8; return 2; }
    return 0;
}

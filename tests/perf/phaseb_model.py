"""CPU model of phase B of lz4_expand_fast_kernel.

Replays the kernel's work distribution on real reference-compressed blocks, lane by lane, and counts
what the clock64/loop-statistics build measured on the GPU (warp-iterations, lanes holding a piece,
blocked lane-iterations) plus an estimate of the warp-instructions issued, using the per-path
instruction counts read off the SASS of the shipped kernel (cuobjdump -sass, loop 0x1b00-0x2a50):

    header 10 | hand-out 45 (any lane takes a chunk) | literal 18 | match 28 | tail 28 |
    store+flag 12 (any lane completes a chunk) | next-sequence 18 | loop 2

Scheduling model: the 32 warps of the CTA advance in lock step, one loop iteration per tick; a done
flag written in tick t is visible to other warps from tick t+1 (and to higher lanes of the same warp
never earlier than the next iteration -- as in the kernel, where the flag is read before the store).

Variants (--variant):
  base        the shipped schedule: warp w owns 256-byte strips w, w+32, ...; dynamic hand-out per warp
  strip128    same with 128-byte strips (4 KB in flight instead of 8 KB)
  inorder     CTA-wide in-order hand-out (chunks handed out in output order across all warps)
  park        base + a lane whose match piece is blocked parks the chunk and takes another one
              (one parked chunk per lane), resuming the parked chunk when it has no fresh one
  v2          the uniform-body loop of lz4_phaseb_v2.h (same schedule as base, its own instruction counts)
  v2x2        v2 with the body unrolled twice: a lane handles up to two pieces of its chunk per iteration
  twopass     chunks that lie inside ONE literal or match run ("simple": one unaligned 8-byte copy)
              are done first by a short uniform loop (est. 40 instr / iteration), the rest by the
              shipped piece loop
Usage: python tests/perf/phaseb_model.py [--blocks N] [--proba P] [--variant V ...]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle, Reference, have_reference  # noqa: E402

COST = dict(header=10, handout=45, literal=18, match=28, tail=28, store=12, nextseq=18, loop=2,
            simple_iter=40)
# the uniform-body variant (lz4_phaseb_v2.h), per-path counts from its SASS
COST_V2 = dict(header=18, handout=40, record=17, body=33, tail=15, store=13, step=6, loop=2)


def parse_sequences(b, total):
    """-> arrays op (sequence output start), m (match start), e (end), off (offset; 0 for the last sequence)."""
    n, p, o = len(b), 0, 0
    ops, ms, es, offs = [], [], [], []
    while True:
        tok = b[p]
        p += 1
        ll = tok >> 4
        if ll == 15:
            while True:
                s = b[p]
                p += 1
                ll += s
                if s != 255:
                    break
        ops.append(o)
        o += ll
        p += ll
        ms.append(o)
        if p >= n:
            es.append(o)
            offs.append(0)
            break
        off = b[p] | (b[p + 1] << 8)
        p += 2
        ml = (tok & 15) + 4
        if (tok & 15) == 15:
            while True:
                s = b[p]
                p += 1
                ml += s
                if s != 255:
                    break
        o += ml
        es.append(o)
        offs.append(off)
    assert o == total, (o, total)
    return np.array(ops), np.array(ms), np.array(es), np.array(offs)


class Block:
    def __init__(self, comp, total):
        self.total = total
        self.op, self.m, self.e, self.off = parse_sequences(comp, total)
        self.nseq = len(self.op)
        # sequence index of every output byte
        self.seq_of = np.searchsorted(self.op, np.arange(total), side="right") - 1
        # guard against empty sequences sharing a start (last empty sequence)
        self.nchunks = (total + 7) >> 3

    def chunk_pieces(self, c):
        """pieces of chunk c: list of (pos, end, kind, srcpos, off)  kind 0 literal, 1 match"""
        p, pe = c * 8, min(c * 8 + 8, self.total)
        out, pos = [], p
        k = int(self.seq_of[p])
        while pos < pe:
            while self.e[k] <= pos and k + 1 < self.nseq:
                k += 1
            if pos < self.m[k]:
                end = min(int(self.m[k]), pe)
                out.append((pos, end, 0, -1, 0))
            else:
                end = min(int(self.e[k]), pe)
                out.append((pos, end, 1, pos - int(self.off[k]), int(self.off[k])))
            pos = end
        return out


def chunk_order(variant, total, nwarps=32):
    """per-warp list of chunk ids in hand-out order (None for a CTA-wide list)."""
    nchunks_pad = ((total + 255) >> 8) << 5
    if variant == "inorder":
        return None, list(range((total + 7) >> 3))
    strip_chunks = 16 if variant == "strip128" else 32
    nstrips = (nchunks_pad + strip_chunks - 1) // strip_chunks
    lists = [[] for _ in range(nwarps)]
    for s in range(nstrips):
        w = s % nwarps
        for j in range(strip_chunks):
            c = s * strip_chunks + j
            if c * 8 < total:
                lists[w].append(c)
    return lists, None


def simulate(blk, variant="base", nwarps=32, only_chunks=None):
    per_tick = 2 if variant == "v2x2" else 1
    v2 = variant in ("v2", "v2x2")
    """returns dict(iters, lane_iters, blocked, instr, pieces)"""
    total = blk.total
    lists, global_list = chunk_order("base" if variant in ("park", "twopass", "v2", "v2x2") else variant, total, nwarps)
    if only_chunks is not None:
        if lists is not None:
            lists = [[c for c in l if c in only_chunks] for l in lists]
        else:
            global_list = [c for c in global_list if c in only_chunks]
    done = np.zeros(blk.nchunks + 1, dtype=bool)
    if only_chunks is not None:
        # chunks outside this pass were completed by an earlier pass
        mask = np.ones(blk.nchunks + 1, dtype=bool)
        mask[list(only_chunks)] = False
        done |= mask
    nxt = [0] * nwarps
    gnext = 0
    # lane state: current chunk pieces + index, optional parked chunk
    cur = [[None] * 32 for _ in range(nwarps)]
    parked = [[None] * 32 for _ in range(nwarps)]
    iters = lane_iters = blocked = instr = pieces_done = 0
    live = True
    while live:
        live = False
        newly_done = []
        for w in range(nwarps):
            lanes = cur[w]
            took = False
            # hand-out
            for l in range(32):
                if lanes[l] is None:
                    c = None
                    if lists is not None:
                        if nxt[w] < len(lists[w]):
                            c = lists[w][nxt[w]]
                            nxt[w] += 1
                    else:
                        if gnext < len(global_list):
                            c = global_list[gnext]
                            gnext += 1
                    if c is not None:
                        lanes[l] = [c, blk.chunk_pieces(c), 0]
                        took = True
                    elif variant == "park" and parked[w][l] is not None:
                        lanes[l], parked[w][l] = parked[w][l], None
            act = [l for l in range(32) if lanes[l] is not None]
            if not act:
                continue
            live = True
            iters += 1
            lane_iters += len(act)
            any_lit = any_match = any_ok = any_store = any_next = False
            sub_live = [False] * per_tick
            for l in act:
                for sub in range(per_tick):
                    if lanes[l] is None:
                        break
                    c, pcs, i = lanes[l]
                    pos, end, kind, src, off = pcs[i]
                    ok = True
                    sub_live[sub] = True
                    if kind == 0:
                        any_lit = True
                    else:
                        any_match = True
                        if off >= 8:
                            ok = done[src >> 3] and done[(end - 1 - off) >> 3]
                        elif off > 0:
                            lo = max(src, 0)
                            # bytes before this chunk must be final (same-chunk bytes come from the accumulator)
                            if lo < c * 8:
                                ok = bool(done[lo >> 3])
                    if not ok:
                        blocked += 1
                        if variant == "park":                  # park it; with a parked chunk already, swap the two
                            parked[w][l], lanes[l] = lanes[l], parked[w][l]
                        break
                    any_ok = True
                    pieces_done += 1
                    i += 1
                    if i == len(pcs):
                        newly_done.append(c)
                        lanes[l] = None
                        any_store = True
                    else:
                        lanes[l][2] = i
                        any_next = any_next or (pcs[i][2] == 0)      # a new sequence starts with its literal run
            if v2:
                body = COST_V2["record"] + COST_V2["body"] + COST_V2["tail"] + COST_V2["step"]
                instr += (COST_V2["header"] + COST_V2["loop"] + (COST_V2["handout"] if took else 0)
                          + body * sum(1 for x in sub_live if x) + (COST_V2["store"] if any_store else 0))
            else:
                instr += (COST["header"] + COST["loop"] + (COST["handout"] if took else 0) + (COST["literal"] if any_lit else 0)
                          + (COST["match"] if any_match else 0) + (COST["tail"] if any_ok else 0)
                          + (COST["store"] if any_store else 0) + (COST["nextseq"] if any_next else 0))
        for c in newly_done:
            done[c] = True
        if variant == "park":
            live = live or any(p is not None for w in range(nwarps) for p in parked[w])
    return dict(iters=iters, lane_iters=lane_iters, blocked=blocked, instr=instr, pieces=pieces_done)


def simulate_twopass(blk):
    simple = set()
    for c in range(blk.nchunks):
        pcs = blk.chunk_pieces(c)
        if len(pcs) == 1 and (pcs[0][2] == 0 or pcs[0][4] >= 8):
            simple.add(c)
    # pass 1: simple chunks, dynamic hand-out in output order per warp strips; a blocked lane retries
    r1 = simulate(blk, "base", only_chunks=simple)
    instr1 = r1["iters"] * COST["simple_iter"]
    rest = set(range(blk.nchunks)) - simple
    # pass 2 runs after pass 1: the simple chunks are final
    r2 = simulate(blk, "base", only_chunks=rest)
    return dict(iters=r1["iters"] + r2["iters"], lane_iters=r1["lane_iters"] + r2["lane_iters"],
                blocked=r1["blocked"] + r2["blocked"], instr=instr1 + r2["instr"], pieces=r1["pieces"] + r2["pieces"],
                simple_frac=len(simple) / blk.nchunks, pass1_iters=r1["iters"], pass2_iters=r2["iters"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=4)
    ap.add_argument("--proba", type=float, default=0.5)
    ap.add_argument("--variant", nargs="*", default=["base", "v2", "v2x2", "strip128", "inorder", "park", "twopass"])
    a = ap.parse_args()
    codec = Reference() if have_reference() else Oracle()
    # block 0 of a generated buffer is an outlier (dependency chains ~190 deep against ~25 for every
    # other block): skip it, the 4 GiB workload has one such block per 64 MiB segment
    data = codec.datagen((a.blocks + 1) * 65536, a.proba, 0)
    blks = []
    for k in range(1, a.blocks + 1):
        _, c = codec.compress(bytes(data[k * 65536:(k + 1) * 65536]), 1)
        blks.append(Block(bytes(c), 65536))
    out = {"proba": a.proba, "blocks": a.blocks, "mean_seq": float(np.mean([b.nseq for b in blks])), "variants": {}}
    for v in a.variant:
        acc = {}
        for b in blks:
            r = simulate_twopass(b) if v == "twopass" else simulate(b, v)
            for key, val in r.items():
                acc[key] = acc.get(key, 0) + val / a.blocks
        acc["lanes_active_of_32"] = acc["lane_iters"] / acc["iters"]
        acc["blocked_frac"] = acc["blocked"] / acc["lane_iters"]
        out["variants"][v] = {k: round(val, 3) for k, val in acc.items()}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

"""CPU study for the round-2 intra-block parallel scan.

Question: if a 64 KB compressed block is cut into fixed segments and every segment is parsed
speculatively from its first byte (treating that byte as a token), after how many bytes / sequences
does the speculative token chain land on a TRUE token position (from where on it is identical to the
serial parse)?  Also: how many of the 32 chains started at 32 consecutive byte positions are still
distinct after k bytes (chains merge and never split again)?

Pure Python on reference-compressed data (oracle/_ref when present, else the oracle port); no GPU.
Usage: python tests/perf/sync_study.py [nBlocks] [proba]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle, Reference, have_reference  # noqa: E402


def next_token(b, p):
    """Position of the token that follows the sequence whose token is at p (or None past the end)."""
    n = len(b)
    tok = b[p]
    p += 1
    ll = tok >> 4
    if ll == 15:
        while True:
            if p >= n:
                return None
            s = b[p]
            p += 1
            ll += s
            if s != 255:
                break
    p += ll + 2
    if p > n:
        return None
    if (tok & 15) == 15:
        while True:
            if p >= n:
                return None
            s = b[p]
            p += 1
            if s != 255:
                break
    return p if p < n else None


def main():
    n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    proba = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    codec = Reference() if have_reference() else Oracle()
    data = codec.datagen(n_blocks * 65536, proba, 0)
    rows = {}
    for seg in (256, 512, 1024, 2048):
        rows[seg] = {"bytes": [], "seqs": [], "never": 0, "segments": 0}
    merge_hist = []
    nseq_total = 0
    csize_total = 0
    for k in range(n_blocks):
        _, c = codec.compress(bytes(data[k * 65536:(k + 1) * 65536]), 1)
        b = bytes(c)
        n = len(b)
        csize_total += n
        true_pos = np.zeros(n + 1, dtype=bool)
        p = 0
        while p is not None:
            true_pos[p] = True
            nseq_total += 1
            p = next_token(b, p)
        for seg, r in rows.items():
            for s in range(seg, n, seg):
                r["segments"] += 1
                p, steps = s, 0
                while p is not None and not true_pos[p]:
                    p = next_token(b, p)
                    steps += 1
                if p is None:
                    r["never"] += 1
                else:
                    r["bytes"].append(p - s)
                    r["seqs"].append(steps)
        # 32 chains from 32 consecutive starts: number of distinct chains after 64/128/256/512 bytes
        for s in range(1024, n - 2048, 4096):
            heads = list(range(s, s + 32))
            row = []
            for horizon in (64, 128, 256, 512, 1024):
                cur = []
                for h in heads:
                    p = h
                    while p is not None and p < s + horizon:
                        p = next_token(b, p)
                    cur.append(p)
                heads = cur
                row.append(len(set(x for x in cur if x is not None)))
            merge_hist.append(row)
    out = {"proba": proba, "blocks": n_blocks, "mean_csize": csize_total / n_blocks,
           "mean_seq_per_block": nseq_total / n_blocks, "segments": {}}
    for seg, r in rows.items():
        by = np.array(r["bytes"]) if r["bytes"] else np.zeros(1)
        sq = np.array(r["seqs"]) if r["seqs"] else np.zeros(1)
        out["segments"][seg] = {
            "n": r["segments"], "never_synced": r["never"],
            "bytes_to_sync": {"mean": float(by.mean()), "p50": float(np.percentile(by, 50)),
                              "p90": float(np.percentile(by, 90)), "p99": float(np.percentile(by, 99)),
                              "max": float(by.max())},
            "seqs_to_sync": {"mean": float(sq.mean()), "p90": float(np.percentile(sq, 90)), "max": float(sq.max())},
            "frac_synced_within_segment": float((by <= seg).mean()),
        }
    mh = np.array(merge_hist)
    out["distinct_chains_of_32_after_bytes"] = {str(h): float(mh[:, i].mean())
                                                for i, h in enumerate((64, 128, 256, 512, 1024))}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

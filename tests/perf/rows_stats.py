"""CPU study behind DESIGN.md sections 5 and 8.1: what the rows expand kernel's waves have to do on the BASELINE data.
For reference-compressed 64 KB blocks of datagen P20 / P50 / P90: runs per block, how many 32-byte rows and 4-byte words lie
inside ONE run (candidates for word-wide copies), and how many bytes / rows have their source inside the same wave (the
"hop" loop) for waves of 1 / 2 / 4 KB.  Usage: python tests/perf/rows_stats.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle, Reference, have_reference  # noqa: E402

BS = 65536


def runs_of(blk, n_out):
    """(start, length, source) per run: source = output position the run copies from, or -1 for literals"""
    out = []
    ip = op = 0
    n = len(blk)
    while ip < n:
        t = blk[ip]; ip += 1
        ll = t >> 4
        if ll == 15:
            while True:
                x = blk[ip]; ip += 1; ll += x
                if x != 255:
                    break
        if ll:
            out.append((op, ll, -1))
        ip += ll; op += ll
        if ip >= n:
            break
        off = blk[ip] | (blk[ip + 1] << 8); ip += 2
        ml = t & 15
        if ml == 15:
            while True:
                x = blk[ip]; ip += 1; ml += x
                if x != 255:
                    break
        ml += 4
        out.append((op, ml, op - off))
        op += ml
    assert op == n_out
    return out


def main():
    codec = Reference() if have_reference() else Oracle()
    print("%-5s %7s %8s %9s %9s   %s" % ("data", "runs", "run len", "rows in", "words in", "bytes (rows) whose source lies in the same wave: 1 KB | 2 KB | 4 KB"))
    print("%-5s %7s %8s %9s %9s" % ("", "/block", "(bytes)", "one run", "one run"))
    for proba in (0.2, 0.5, 0.9):
        raw = codec.datagen(16 * BS, proba, 7)
        acc = {"runs": 0, "rows1": 0, "words1": 0, "hopb": {1024: 0, 2048: 0, 4096: 0}, "hopr": {1024: 0, 2048: 0, 4096: 0}}
        for b in range(16):
            d = bytes(raw[b * BS:(b + 1) * BS])
            _, comp = codec.compress(np.frombuffer(d, dtype=np.uint8), 1)
            rs = runs_of(comp, BS)
            acc["runs"] += len(rs)
            run_id = np.zeros(BS, dtype=np.int32)
            src = np.full(BS, -1, dtype=np.int64)
            for k, (st, ln, so) in enumerate(rs):
                run_id[st:st + ln] = k
                if so >= 0:
                    src[st:st + ln] = np.arange(so, so + ln)
            r = run_id.reshape(-1, 32)
            acc["rows1"] += int((r.min(axis=1) == r.max(axis=1)).sum())
            w = run_id.reshape(-1, 4)
            acc["words1"] += int((w.min(axis=1) == w.max(axis=1)).sum())
            pos = np.arange(BS)
            for wave in (1024, 2048, 4096):
                inwave = (src >= 0) & (src >= (pos // wave) * wave)
                acc["hopb"][wave] += int(inwave.sum())
                acc["hopr"][wave] += int(inwave.reshape(-1, 32).any(axis=1).sum())
        nb, nrows, nwords = 16, 16 * BS // 32, 16 * BS // 4
        print("P%-4d %7d %8.1f %8.1f%% %8.1f%%   %s" % (
            int(proba * 100), acc["runs"] // nb, 16 * BS / acc["runs"], 100 * acc["rows1"] / nrows, 100 * acc["words1"] / nwords,
            " | ".join("%4.1f%% (%4.1f%%)" % (100 * acc["hopb"][w] / (16 * BS), 100 * acc["hopr"][w] / nrows) for w in (1024, 2048, 4096))))


if __name__ == "__main__":
    main()

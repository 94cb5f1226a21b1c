"""Developer tool: per-phase cycle breakdown of the parallel-parse compressor (build with -DLZ4K_PHASE_TIMING)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    from lz4_b200 import _lib, batch
    from oracle.pyoracle import Oracle
    _lib.load()
    raw = C.CDLL(_lib.LIB_PATH)
    orc = Oracle()
    n_blocks, bs = 8192, 65536
    data = orc.datagen_mt(n_blocks * bs, 64 << 20, float(os.environ.get("PROBA", "0.5")), 0)
    src = torch.from_numpy(data).cuda()
    slots, sizes, stride = batch.compress_blocks(src, bs, 1, mode="parallel")
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 12)()
    raw.LZ4B200_debug_phase_cycles(buf)
    reps = 3
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        batch.compress_blocks(src, bs, 1, slots=slots, out_sizes=sizes, mode="parallel")
    t1.record()
    torch.cuda.synchronize()
    raw.LZ4B200_debug_phase_cycles(buf)
    names = ["find 1 (T2: earliest of window)", "find 2 (candidate bits)", "last literals", "-", "select (chain rounds)", "emit + insert", "load + tables", "-"]
    tot = sum(buf[:8])
    print("compress ms per launch %.3f, blocks %d, ratio %.4f" % (t0.elapsed_time(t1) / reps, n_blocks, n_blocks * bs / float(sizes.sum())))
    for n, v in zip(names, buf[:8]):
        print("%-28s %8.0f cycles/block  %5.1f%%" % (n, v / (reps * n_blocks), 100.0 * v / max(tot, 1)))
    wins = max(buf[11], 1)
    print("select: %.1f rounds per window, %.0f lane walks per window (1024 lanes), %d windows per block" % (
        buf[8] / wins, buf[9] / wins, wins / ((reps + 1) * n_blocks)))


if __name__ == "__main__":
    main()

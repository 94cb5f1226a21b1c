"""`lz4 -b#` for the GPU batch codec.

Host mirror of the reference's in-memory benchmark `BMK_benchMem` (programs/bench.c:360-619), so
that the numbers it prints can be read next to the tool users know:

  * same block split: every file is cut into `blockSize` pieces, blocks never straddle files, the
    last block of a file is ragged (bench.c:393-417); each block gets `LZ4_compressBound` room;
  * same level rule: `acceleration = level < 0 ? -level + 1 : 1` (bench.c:229); levels >= 2
    (`LZ4HC_CLEVEL_MIN`, lz4hc.h:47) are the HC family, which is outside this library
    (bench.c:306-310) and refused;
  * same default block size: none -- without `-B#` every file is a single block (bench.c:123,365);
    `-B4` gives the 64 KB blocks of the headline configuration;
  * same timing rule: passes are repeated in loops sized to ~1 s until `nbSeconds` of work has been
    done, the speed reported is the FASTEST pass (bench.c:431-445,480-492,543-554), decode through
    the call with capacity == the block's original size (bench.c:525-533);
  * same verification: the regenerated buffer must equal the source (the reference compares XXH64
    digests, bench.c:567-594; here the two device buffers are compared byte for byte);
  * same result lines (bench.c:499-503,560-565,600-609).

What differs, by construction: one "pass" is ONE batched call over all blocks
(`LZ4B200_compress_blocks` / `LZ4B200_decompress_blocks`) instead of a serial loop over blocks, the
source sits in device memory for the whole run (the reference's source sits in host memory for the
whole run), and passes are timed on the device with CUDA events.

The harness itself (`bench_mem`) is codec-agnostic: it drives any object with the `Codec` methods
below.  `GpuCodec` is the product codec; there is no CPU codec in this package.

CLI:  python -m lz4_b200.lz4bench [-b#] [-e#] [--fast=#] [-i#] [-B#] [-q] FILE...
"""
import os
import sys
from dataclasses import dataclass, field

TIMELOOP_NS = 1_000_000_000           # bench.c:69
DEFAULT_SECONDS = 3                   # bench.c:67
HC_CLEVEL_MIN = 2                     # lz4hc.h:47 LZ4HC_CLEVEL_MIN: levels >= 2 select LZ4_compress_HC (bench.c:306-310)
DEFAULT_BLOCK = 0                     # bench.c:123,365: without -B# every file is ONE block
LZ4_MAX_INPUT_SIZE = 0x7E000000


def level_to_acceleration(level):
    """bench.c:229 -- LZ4_compressBlockNoStream."""
    if level >= HC_CLEVEL_MIN:
        raise ValueError("level %d selects LZ4_compress_HC (bench.c:306-310): not provided by the GPU codec" % level)
    return -level + 1 if level < 0 else 1


def block_size_from_flag(value):
    """lz4cli.c -B#: 4..7 are block size IDs (64 KB .. 4 MB), values >= 32 are bytes."""
    if 4 <= value <= 7:
        return 1 << (8 + 2 * value)
    if value < 32:
        raise ValueError("-B%d: block size must be an ID in 4..7 or >= 32 bytes" % value)
    return value


def split_blocks(file_sizes, block_size):
    """bench.c:393-417: list of (source offset, size); blocks never straddle files."""
    blocks, pos = [], 0
    for fsize in file_sizes:
        remaining = fsize
        while remaining > 0:
            this = min(remaining, block_size)
            blocks.append((pos, this))
            pos += this
            remaining -= this
    return blocks


@dataclass
class BenchResult:
    name: str
    level: int
    src_size: int
    c_size: int
    ratio: float
    c_ns: int                 # fastest compression pass
    d_ns: int                 # fastest decompression pass
    error: int
    passes: dict = field(default_factory=dict)

    @property
    def c_speed(self):        # MB/s as bench.c prints it: bytes / ns * 1000
        return self.src_size / self.c_ns * 1000.0 if self.c_ns else 0.0

    @property
    def d_speed(self):
        return self.src_size / self.d_ns * 1000.0 if self.d_ns else 0.0

    def line(self):
        """bench.c:560-565 followed by :598 (the final, persistent form of the progress line)."""
        return "%2i#%-17.17s :%10u ->%10u (%5.3f),%6.1f MB/s, %6.1f MB/s" % (
            self.level, self.name[-17:], self.src_size, self.c_size, self.ratio, self.c_speed, self.d_speed)

    def quiet_line(self):
        """bench.c:604 (-q)."""
        return "-%-3i%11i (%5.3f) %6.2f MB/s %6.1f MB/s  %s " % (
            self.level, self.c_size, self.ratio, self.c_speed, self.d_speed, self.name[-17:])


class Codec:
    """What bench_mem drives.  A pass = all blocks once."""

    def setup(self, src, blocks, acceleration):            # src: bytes-like of the concatenated files
        raise NotImplementedError

    def compress_passes(self, n):                          # -> elapsed ns for n passes
        raise NotImplementedError

    def compressed_sizes(self):                            # -> list of per-block return values of the last pass
        raise NotImplementedError

    def decompress_passes(self, n):                        # -> elapsed ns for n passes
        raise NotImplementedError

    def decoded_sizes(self):                               # -> list of per-block return values of the last pass
        raise NotImplementedError

    def verify(self):                                      # -> index of the first differing byte, or -1
        raise NotImplementedError


def _timed_loop(run, nb_seconds, first_loops):
    """The loop structure of bench.c:464-492 / :514-554 for one direction.
    Returns (fastest ns per pass, total passes, total ns)."""
    max_time = nb_seconds * TIMELOOP_NS + 100
    loops = 1 if nb_seconds == 0 else first_loops
    fastest, total_ns, total_passes = None, 0, 0
    while True:
        ns = int(run(loops))
        total_passes += loops
        if ns > 0:
            if fastest is None or ns < fastest * loops:
                fastest = max(ns // loops, 1)
            loops = TIMELOOP_NS // fastest + 1             # aim for ~1 s
        else:
            loops *= 100
        total_ns += ns
        if total_ns > max_time or nb_seconds == 0:
            break
    return fastest or 1, total_passes, total_ns


def bench_mem(codec, src, file_sizes, name, level, block_size=DEFAULT_BLOCK, nb_seconds=DEFAULT_SECONDS, out=None):
    """BMK_benchMem (bench.c:360) for one level.  Returns a BenchResult (error != 0 on failure)."""
    src_size = sum(file_sizes)
    if src_size > LZ4_MAX_INPUT_SIZE:
        raise ValueError("input larger than LZ4_MAX_INPUT_SIZE (bench.c:706-708 truncates; split the input instead)")
    bs = block_size if block_size >= 32 else max(src_size, 1)            # bench.c:365
    blocks = split_blocks(file_sizes, bs)
    accel = level_to_acceleration(level)
    codec.setup(src, blocks, accel)
    error = 0

    first_c = (5 << 20) // (src_size + 1) + 1                              # bench.c:433
    first_d = (200 << 20) // (src_size + 1) + 1                            # bench.c:434
    c_ns, c_passes, c_total = _timed_loop(codec.compress_passes, nb_seconds, first_c)
    csizes = codec.compressed_sizes()
    for i, r in enumerate(csizes):
        if r <= 0 and blocks[i][1] > 0:
            print("LZ4 compression failed on block %u " % i, file=out or sys.stderr)
            error = 1
    c_size = sum(int(r) for r in csizes) or 1
    ratio = src_size / c_size

    d_ns, d_passes, d_total = _timed_loop(codec.decompress_passes, nb_seconds, first_d)
    for i, r in enumerate(codec.decoded_sizes()):
        if r < 0:
            print("LZ4_decompress_safe_usingDict() failed on block %u of size %u " % (i, blocks[i][1]),
                  file=out or sys.stderr)
            error = 1
            break
    if not error:
        bad = codec.verify()
        if bad >= 0:
            acc, seg = 0, 0
            for seg, (_, sz) in enumerate(blocks):
                if acc + sz > bad:
                    break
                acc += sz
            print("\n!!! WARNING !!! %17s : Invalid Checksum" % name, file=out or sys.stderr)
            print("Decoding error at pos %u (block %u, sub %u, pos %u) " % (bad, seg, (bad - acc) // (128 << 10), bad - acc),
                  file=out or sys.stderr)
            error = 1
    return BenchResult(name, level, src_size, c_size, ratio, c_ns, d_ns, error,
                       {"blocks": len(blocks), "block_size": bs, "acceleration": accel,
                        "compress_passes": c_passes, "compress_ns": c_total,
                        "decompress_passes": d_passes, "decompress_ns": d_total})


class GpuCodec(Codec):
    """The batch layer of liblz4_b200.so on device-resident buffers (one launch sequence per pass)."""

    def __init__(self, device="cuda:0"):
        import torch
        from . import _lib
        self.torch = torch
        self.lib = _lib.load()
        self._check = _lib.check
        self.device = torch.device(device)
        if not torch.cuda.is_available():
            raise RuntimeError("lz4bench needs a CUDA device: the codec has no CPU path")

    def setup(self, src, blocks, acceleration):
        import numpy as np
        torch = self.torch
        self.accel = int(acceleration)
        n = len(blocks)
        self.n = n
        bs = max((sz for _, sz in blocks), default=1)
        self.bs = bs
        host = np.frombuffer(src, dtype=np.uint8) if not isinstance(src, np.ndarray) else src
        # device layout: block i at i*bs (files start on a block boundary, ragged blocks are short)
        staged = np.zeros(max(n, 1) * bs, dtype=np.uint8)
        for i, (off, sz) in enumerate(blocks):
            staged[i * bs:i * bs + sz] = host[off:off + sz]
        self.d_src = torch.from_numpy(staged).to(self.device)
        sizes = np.array([sz for _, sz in blocks] or [0], dtype=np.int32)
        self.d_src_sizes = torch.from_numpy(sizes).to(self.device)
        cap = int(self.lib.LZ4_compressBound(bs))                          # cRoom, bench.c:408
        self.cap = cap
        self.stride = (cap + 15) // 16 * 16
        self.d_slots = torch.empty(max(n, 1) * self.stride, dtype=torch.uint8, device=self.device)
        self.d_csize = torch.zeros(max(n, 1), dtype=torch.int32, device=self.device)
        self.d_slot_off = torch.arange(max(n, 1), dtype=torch.int64, device=self.device) * self.stride
        self.d_res = torch.empty(max(n, 1) * bs, dtype=torch.uint8, device=self.device)
        self.d_rsize = torch.zeros(max(n, 1), dtype=torch.int32, device=self.device)
        ws = int(self.lib.LZ4B200_decompress_workspace_bytes(max(n, 1)))
        self.d_ws = torch.empty(ws, dtype=torch.uint8, device=self.device)
        self.stream = torch.cuda.current_stream(self.device)

    def _time(self, n, fn):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(self.device)
        e0.record(self.stream)
        for _ in range(n):
            fn()
        e1.record(self.stream)
        e1.synchronize()
        return int(e0.elapsed_time(e1) * 1e6)

    def _compress_once(self):
        rc = self.lib.LZ4B200_compress_blocks(self.d_src.data_ptr(), self.bs, self.d_src_sizes.data_ptr(), self.bs,
                                              self.d_slots.data_ptr(), self.stride, self.cap, self.accel,
                                              self.d_csize.data_ptr(), self.n, self.stream.cuda_stream)
        self._check(rc, "LZ4B200_compress_blocks")

    def _decompress_once(self):
        # capacity of block i = its original size (bench.c:525-533)
        rc = self.lib.LZ4B200_decompress_blocks(self.d_slots.data_ptr(), self.d_slot_off.data_ptr(), self.d_csize.data_ptr(),
                                                self.d_res.data_ptr(), None, self.bs, self.d_src_sizes.data_ptr(), 0,
                                                self.d_rsize.data_ptr(), self.n, self.d_ws.data_ptr(), self.d_ws.numel(),
                                                self.stream.cuda_stream)
        self._check(rc, "LZ4B200_decompress_blocks")

    def compress_passes(self, n):
        if self.n == 0:
            return 1
        self.d_slots.fill_(0xE5)                                           # bench.c:457
        return self._time(n, self._compress_once)

    def compressed_sizes(self):
        return self.d_csize[:self.n].cpu().tolist()

    def decompress_passes(self, n):
        if self.n == 0:
            return 1
        self.d_res.fill_(0xD6)                                             # bench.c:510
        return self._time(n, self._decompress_once)

    def decoded_sizes(self):
        return self.d_rsize[:self.n].cpu().tolist()

    def verify(self):
        torch = self.torch
        if self.n == 0:
            return -1
        if not torch.equal(self.d_rsize[:self.n], self.d_src_sizes[:self.n]):
            i = int((self.d_rsize[:self.n] != self.d_src_sizes[:self.n]).nonzero()[0])
            return i * self.bs
        # bytes past a ragged block's end are not part of the file: mask them out of the comparison
        idx = torch.arange(self.bs, device=self.device, dtype=torch.int32)
        valid = idx.unsqueeze(0) < self.d_src_sizes[:self.n].unsqueeze(1)
        diff = ((self.d_res.view(-1, self.bs)[:self.n] != self.d_src.view(-1, self.bs)[:self.n]) & valid)
        if not bool(diff.any()):
            return -1
        flat = diff.view(-1).nonzero()[0]
        blk, pos = int(flat) // self.bs, int(flat) % self.bs
        return int(self.d_src_sizes[:blk].sum().item()) + pos


def load_files(paths, limit=LZ4_MAX_INPUT_SIZE):
    """BMK_loadFiles (bench.c:676-708): concatenate, stop at `limit` bytes."""
    chunks, sizes, total = [], [], 0
    for p in paths:
        if os.path.isdir(p):
            print("Ignoring %s directory...       " % p, file=sys.stderr)
            sizes.append(0)
            continue
        with open(p, "rb") as f:
            data = f.read(max(limit - total, 0))
        chunks.append(data)
        sizes.append(len(data))
        total += len(data)
        if total >= limit:
            break
    if total == 0:
        raise SystemExit("no data to bench")
    return b"".join(chunks), sizes


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    level, level_last, seconds, block, quiet, files = 1, None, DEFAULT_SECONDS, DEFAULT_BLOCK, False, []
    for a in argv:
        if a.startswith("-b") and a[2:].lstrip("-").isdigit():
            level = int(a[2:])
        elif a == "-b":                                                    # bench mode, level unchanged (lz4cli.c:652)
            pass
        elif a.startswith("-e") and a[2:].lstrip("-").isdigit():
            level_last = int(a[2:])
        elif a.startswith("--fast"):                                       # lz4cli.c:497-509
            level = -int(a[7:]) if a.startswith("--fast=") else -1
        elif a.startswith("-i") and a[2:].isdigit():
            seconds = int(a[2:])
        elif a.startswith("-B") and a[2:].isdigit():
            block = block_size_from_flag(int(a[2:]))
        elif a == "-q":
            quiet = True
        elif a.startswith("-") and a != "-":
            raise SystemExit("unknown option %s (supported: -b# -e# --fast=# -i# -B# -q)" % a)
        else:
            files.append(a)
    if not files:
        raise SystemExit("usage: python -m lz4_b200.lz4bench [-b#] [-e#] [--fast=#] [-i#] [-B#] [-q] FILE...\n"
                         "(the reference's no-file mode benches a generated Lorem ipsum text, bench.c:765-793; "
                         "this tool benches files only)")
    if level_last is None or level_last < level:
        level_last = level
    src, sizes = load_files(files)
    name = os.path.basename(files[0]) if len(files) == 1 else " %u files" % len(files)
    codec = GpuCodec()
    if quiet:
        print("bench lz4_b200: input %u bytes, %u seconds, %u KB blocks" % (len(src), seconds, block >> 10), file=sys.stderr)
    err = 0
    for lv in range(level, level_last + 1):
        res = bench_mem(codec, src, sizes, name, lv, block, seconds)
        print(res.quiet_line() if quiet else res.line())
        err |= res.error
    return err


if __name__ == "__main__":
    sys.exit(main())

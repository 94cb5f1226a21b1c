"""Developer check: the parallel compressor must give the same, valid bytes whatever the destination layout / capacity --
and whatever tool watches it.  Runs the same 8 blocks into differently laid out slot buffers and reports sizes, first
differing byte and the oracle's decode of every block; `--poison` also fills the CTAs' shared memory with patterns before
each run (a kernel that read shared memory it never wrote would give pattern-dependent output).
Usage (under gpurun):  python tests/perf/enc_determinism.py [--poison]
                       compute-sanitizer --tool memcheck python tests/perf/enc_determinism.py     (same sizes, every block decodes)
This is the check that exposed the warp-cooperative match finish of round 2."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lz4_b200 import batch  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

BS = 65536
orc = Oracle()
d = orc.datagen_mt(8 * BS, 1 << 20, 0.5, 21)
src = torch.from_numpy(d).cuda()


def run(cap=None, extra=0, fill=None):
    if cap is None:
        slots, sizes, stride = batch.compress_blocks(src, BS, 1, mode="parallel")
    else:
        stride = (cap + 15) // 16 * 16 + extra
        slots = torch.full((8 * stride,), 0xA5 if fill is None else fill, dtype=torch.uint8, device="cuda")
        sizes = torch.zeros(8, dtype=torch.int32, device="cuda")
        batch.compress_blocks(src, BS, 1, slots=slots, slot_stride=stride, slot_capacity=cap, out_sizes=sizes, mode="parallel")
    torch.cuda.synchronize()
    return slots.cpu().numpy(), sizes.cpu().numpy(), stride


ref_h, ref_s, ref_st = run()
print("run 0 sizes", ref_s.tolist())
for i in range(8):
    blk = ref_h[i * ref_st:i * ref_st + ref_s[i]].tobytes()
    ret, out = orc.decompress(blk, BS)
    want = d[i * BS:(i + 1) * BS].tobytes()
    if (ret, out) != (BS, want):
        bad = [k for k in range(min(len(out), BS)) if out[k] != want[k]]
        print("run 0 block %d DOES NOT DECODE to the input: ret %d, %d wrong bytes, first at %s, last at %s" %
              (i, ret, len(bad), bad[:8], bad[-3:]))
    else:
        print("run 0 block %d decodes" % i)
np.save(os.path.join(ROOT, "gpurun_out", "enc_det_%s.npy" % os.environ.get("TAG", "plain")), np.concatenate([ref_s.astype(np.uint8).view(np.uint8), ref_h]))
cfgs = [(None, 0), (int(ref_s.max()), 64), (int(ref_s.max()), 64), (int(ref_s.max()) + 5, 80), (70000, 16), (None, 0)]
for n, (cap, extra) in enumerate(cfgs, 1):
    h, s, st = run(cap, extra)
    bad = []
    for i in range(8):
        a = ref_h[i * ref_st:i * ref_st + ref_s[i]]
        b = h[i * st:i * st + s[i]]
        ok_dec = orc.decompress(b.tobytes(), BS) == (BS, d[i * BS:(i + 1) * BS].tobytes()) if s[i] > 0 else None
        if s[i] != ref_s[i] or not np.array_equal(a, b):
            m = min(len(a), len(b))
            diff = np.nonzero(a[:m] != b[:m])[0]
            bad.append((i, int(ref_s[i]), int(s[i]), int(diff[0]) if len(diff) else m, ok_dec))
    print("run %d cap %s stride %d: %s" % (n, cap, st, "identical" if not bad else "DIFFERS (block, size0, size, first diff, decodes) %s" % bad))

# ---- the capacity ladder of test_limited_output_and_never_past_capacity ----
for cap in (int(ref_s.max()), int(ref_s.max()) - 1, int(ref_s.min()), int(ref_s.min()) - 1, 1000, 1):
    h, s, st = run(cap, 64)
    rep = []
    for i in range(8):
        if ref_s[i] <= cap:
            a = ref_h[i * ref_st:i * ref_st + ref_s[i]]
            b = h[i * st:i * st + max(int(s[i]), 0)]
            if s[i] != ref_s[i] or not np.array_equal(a, b):
                m = min(len(a), len(b))
                diff = np.nonzero(a[:m] != b[:m])[0]
                dec = orc.decompress(b.tobytes(), BS) == (BS, d[i * BS:(i + 1) * BS].tobytes()) if s[i] > 0 else None
                rep.append((i, int(ref_s[i]), int(s[i]), int(diff[0]) if len(diff) else m, int(len(diff)), dec))
        elif s[i] != 0:
            rep.append((i, "should not fit", int(s[i])))
    print("cap %d: %s" % (cap, "ok" if not rep else "BAD (block, size0, size, first diff, #diff, decodes) %s" % rep))
if "--poison" not in sys.argv:
    sys.exit(0)
# ---- does the result depend on shared memory the kernel never wrote?  poison regions of it before a run ----
import ctypes as C  # noqa: E402
from lz4_b200 import _lib  # noqa: E402
lib = _lib.load()
lib.LZ4B200_debug_poison_smem.restype = C.c_int
lib.LZ4B200_debug_poison_smem.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_void_p]
regions = {"all": (0, 1 << 20), "pad": (0, 16), "src": (16, 65616), "src tail": (16 + 65536, 65616), "TT": (65616, 98384),
           "stage": (98384, 107088), "small arrays": (107088, 1 << 20)}
for name, (lo, hi) in regions.items():
    outs = []
    for pattern in (0x00000000, 0xFFFFFFFF, 0x5A17C3E9):
        assert lib.LZ4B200_debug_poison_smem(pattern, lo, hi, torch.cuda.current_stream().cuda_stream) == 0
        h, s, st = run()
        outs.append((s.tolist(), [int(np.bitwise_xor.reduce(h[i * st:i * st + s[i]].astype(np.uint8))) for i in range(8)]))
    same = all(o == outs[0] for o in outs)
    print("poison %-12s -> %s" % (name, "same result" if same else "RESULT DEPENDS ON IT: %s" % [o[0] for o in outs]))

"""tcgen05.mma issue-rate ceilings on this GPU (mipnerf_b200_selftest_umma_rate): cycles per M=128 x N x K=16 MMA for
the SS form (A from shared memory, two layouts) and the TS form (A from tensor memory), one CTA alone and one CTA on
every SM.  Ideal: N / 2 cycles (8192 dense 16-bit FLOP per clock per SM)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from mipnerf_pl_b200 import _cabi  # noqa: E402

lib = _cabi.lib()
dev = torch.device("cuda", 0)
sms = torch.cuda.get_device_properties(dev).multi_processor_count
names = {0: "SS, A in 128-byte-swizzle slabs", 1: "SS, A in dense 32-byte-swizzle K=16 blocks", 2: "TS, A in tensor memory"}
iters = 2000
for n in (256, 128):
    for mode in (0, 1, 2):
        for ctas in (1, sms):
            out = torch.zeros(ctas, dtype=torch.int64, device=dev)
            for _ in range(2):
                _cabi.check(lib.mipnerf_b200_selftest_umma_rate(mode, n, iters, _cabi.BF16, ctas, out.data_ptr(),
                                                                torch.cuda.current_stream().cuda_stream), "umma_rate")
            torch.cuda.synchronize()
            per = out.double() / (iters * 16)
            print(f"N={n:3d} {names[mode]:45s} {ctas:3d} CTA(s): {float(per.mean()):7.1f} cycles per MMA "
                  f"(min {float(per.min()):.1f}, max {float(per.max()):.1f}; ideal {n // 2})")

print("CTA pairs (cta_group::2, M = 256 over two SMs; ideal N / 2 cycles per MMA):")
for n in (256, 128):
    for mode in (0, 2):
        for pairs in (1, sms // 2):
            out = torch.zeros(pairs, dtype=torch.int64, device=dev)
            for _ in range(2):
                _cabi.check(lib.mipnerf_b200_selftest_umma_rate_pair(mode, n, iters, _cabi.BF16, pairs, out.data_ptr(),
                                                                     torch.cuda.current_stream().cuda_stream), "umma_rate_pair")
            torch.cuda.synchronize()
            per = out.double() / (iters * 16)
            print(f"N={n:3d} {names[mode]:45s} {pairs:3d} pair(s): {float(per.mean()):7.1f} cycles per MMA "
                  f"(min {float(per.min()):.1f}, max {float(per.max()):.1f}; ideal {n // 2})")

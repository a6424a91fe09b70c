"""Diagnostic sweep of the UMMA self-test (no asserts): prints one line per variant so a single
GPU call tells which operand-staging conventions are right."""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))
from test_gpu_tc import run_selftest  # noqa: E402

for precision in ("bf16", "fp16"):
    for variant in (4,):
        for n, k in ((128, 32), (256, 32)):
            try:
                d, ref = run_selftest(n, k, precision, variant)
                err = (d.double() - ref).abs()
                nan = int(torch.isnan(d).sum())
                print(f"{precision} variant={variant} n={n:3d} k={k:3d}: max_err={float(err.nan_to_num(1e9).max()):.3e} "
                      f"ref_max={float(ref.abs().max()):.2f} nan={nan} "
                      f"bad_rows={int((err.nan_to_num(1e9).max(1).values > 1e-2).sum())} "
                      f"bad_cols={int((err.nan_to_num(1e9).max(0).values > 1e-2).sum())}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"{precision} variant={variant} n={n} k={k}: EXC {e!r}", flush=True)
                traceback.print_exc()
                try:
                    torch.cuda.synchronize()
                except Exception as e2:  # noqa: BLE001
                    print("device unusable after failure:", e2, flush=True)
                    sys.exit(0)

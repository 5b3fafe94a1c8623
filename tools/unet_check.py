"""Print parity metrics of the CUDA U-Net against the CPU oracle (run on the GPU box)."""
import os
import sys
import time
import traceback

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "tests"))

import torch  # noqa: E402

import unet_checks as U  # noqa: E402

CHECKS = [
    ("tiny_all_ones", lambda: U.check_all_ones()),
    ("tiny_hard", lambda: U.check_hard()),
    ("tiny_hard_beta", lambda: U.check_hard(beta_std=0.1)),
    ("tiny_soft", lambda: U.check_soft()),
    ("tiny_cfg", lambda: U.check_cfg_doubling()),
    ("full_hard_b2_h32", lambda: U.check_hard(tiny=False, B=2, H=32, code_ids=(0, 3))),
]

if __name__ == "__main__":
    pats = sys.argv[1:]
    for name, fn in CHECKS:
        if pats and not any(p in name for p in pats):
            continue
        t0 = time.time()
        try:
            ma, cos = fn()
            ok = ma <= U.MAX_ABS_TOL and cos >= U.COS_TOL
            print(f"{'PASS' if ok else 'FAIL'} {name}: max_abs/scale={ma:.4g} cos={cos:.6f} ({time.time()-t0:.1f}s)", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"FAIL {name}: {type(e).__name__}: {str(e)[:400]}", flush=True)
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception as e2:  # noqa: BLE001
                print("CUDA context dead:", e2)
                break

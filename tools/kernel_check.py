"""Run every per-kernel numerics check on the GPU without stopping at the first failure.
Usage (on the GPU box): python tools/kernel_check.py [name-substring ...]"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import torch  # noqa: E402

import kernel_checks  # noqa: E402


def main():
    pats = sys.argv[1:]
    n_fail = 0
    import backward_checks
    for name, fn in kernel_checks.ALL + backward_checks.ALL:
        if pats and not any(p in name for p in pats):
            continue
        t0 = time.time()
        try:
            fn()
            torch.cuda.synchronize()
            print(f"PASS {name} ({time.time() - t0:.2f}s)", flush=True)
        except Exception as e:  # noqa: BLE001
            n_fail += 1
            msg = str(e).splitlines()[0] if str(e) else repr(e)
            print(f"FAIL {name}: {type(e).__name__}: {msg}", flush=True)
            if os.environ.get("APTP_TRACE"):
                traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception as e2:  # noqa: BLE001
                print(f"  CUDA context is dead: {e2}", flush=True)
                break
    print(f"{n_fail} failures")
    return 1 if n_fail else 0


if __name__ == "__main__":
    sys.exit(main())

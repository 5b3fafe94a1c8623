"""Run the training-path parity checks verbosely (GPU box)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import train_checks as T  # noqa: E402

if __name__ == "__main__":
    if "grads" in sys.argv or len(sys.argv) == 1:
        out = T.check_train_grads()
        for k, v in out.items():
            print(k, ["%.5g" % x for x in v], flush=True)
        T.assert_train(out)
        print("train parity ok", flush=True)
    if "step" in sys.argv or len(sys.argv) == 1:
        lg, lr, gm = T.check_pruning_step()
        for k in lr:
            print(f"{k:20s} got {lg[k]:.6g} ref {lr[k]:.6g}")
        for k, v in gm.items():
            print(f"grad {k:10s} rel {v[0]:.4g} cos {v[1]:.6f} |ref| {v[2]:.4g}")
        T.assert_step(lg, lr, gm)
        print("pruning step parity ok", flush=True)

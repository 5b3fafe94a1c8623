"""Run the training-path parity check verbosely (GPU box)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import train_checks as T  # noqa: E402

if __name__ == "__main__":
    out = T.check_train_grads()
    for k, v in out.items():
        print(k, ["%.5g" % x for x in v], flush=True)
    T.assert_train(out)
    print("train parity ok")

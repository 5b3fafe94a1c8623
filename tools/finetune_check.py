#!/usr/bin/env python
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import train_checks as T
out = T.check_finetune_grads()
rows = sorted(out.items(), key=lambda kv: -kv[1][0])
print("worst 40 of", len(out))
bad = [kv for kv in out.items() if not (kv[1][0] <= 8e-2 and kv[1][1] >= 0.995)]
print("outside tolerance:", len(bad))
for n, (rel, cos, rn) in bad:
    print(f"BAD {n:66s} rel {rel:9.3g} cos {cos:8.5f} |ref| {rn:9.3g}")
for n, (rel, cos, rn) in rows[:10]:
    print(f"{n:70s} rel {rel:9.3g} cos {cos:8.5f} |ref| {rn:9.3g}")
ok = sum(1 for v in out.values() if v[0] <= 8e-2 and v[1] >= 0.995)
print("within tolerance:", ok, "of", len(out))

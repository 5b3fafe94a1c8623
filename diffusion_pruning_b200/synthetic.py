"""Synthetic architecture codes / inputs for benchmarks, smoke tests and parity tests (there are no
checkpoints or datasets offline). Mirrors SURVEY.md section 8(d), config 1."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

DEPTH_ORDER = [-1, -2, 0, 1, -3, -4, 2, 3, -5, -6, 4, 5, -7, 6]  # configs/pruning/sd-2-1_cc3m.yaml:38


def synthetic_codes(structure: Dict[str, List[List[int]]], n_codes: int = 8, seed: int = 2,
                    keep_lo: float = 0.35, keep_hi: float = 0.95, drop_min: int = 2, drop_max: int = 4,
                    depth_order: Sequence[int] = DEPTH_ORDER) -> torch.Tensor:
    """[n_codes, dim] 0/1 codes: per (code, width gate) a keep probability ~ U(keep_lo, keep_hi), Bernoulli
    bits, unit 0 forced on if a gate came out empty (as estimation_utils.py:27-31 does); each code drops
    drop_min..drop_max depth gates, least important first (reverse of the quantizer's depth_order)."""
    rng = np.random.RandomState(seed)
    widths = [w for ws in structure["width"] for w in ws]
    n_depth = sum(1 for d in structure["depth"] if d == [1])
    order = [i % n_depth for i in depth_order] if n_depth else []
    codes = np.zeros((n_codes, sum(widths) + n_depth), dtype=np.float32)
    for c in range(n_codes):
        col = 0
        for w in widths:
            p = rng.uniform(keep_lo, keep_hi)
            bits = (rng.uniform(size=w) < p).astype(np.float32)
            if bits.sum() == 0:
                bits[0] = 1.0
            codes[c, col:col + w] = bits
            col += w
        d = np.ones(n_depth, dtype=np.float32)
        if n_depth:
            n_drop = rng.randint(drop_min, drop_max + 1)
            for j in range(n_drop):
                d[order[n_depth - 1 - j]] = 0.0  # importance rank n_depth-1 is the first to go
        codes[c, col:] = d
    return torch.from_numpy(codes)


def split_arch(arch: torch.Tensor, structure: Dict[str, List[List[int]]]) -> Dict[str, List[torch.Tensor]]:
    """HyperStructure.transform_structure_vector (pdm/models/hypernet/hypernet.py:86-101)."""
    widths = [w for ws in structure["width"] for w in ws]
    n_w = sum(widths)
    out_w, s = [], 0
    for w in widths:
        out_w.append(arch[:, s:s + w])
        s += w
    n_d = arch.shape[1] - n_w
    return {"width": out_w, "depth": [arch[:, n_w + i] for i in range(n_d)]}

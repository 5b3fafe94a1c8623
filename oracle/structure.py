"""Gate structure of the shipped SD-2.1 layout (configs/pruning/sd-2-1_cc3m.yaml:11-26) without
allocating weights -- TEST INFRASTRUCTURE. Order = UNet2DConditionModelGated.get_structure()
(unet_2d_conditional.py:1332-1363): per block, resnets first, then attentions (blocks.py:1814-1831)."""
from typing import Dict, List, Sequence


def gate_structure(num_heads: Sequence[int] = (5, 10, 20, 20), down_has_attn=(True, True, True, False),
                   layers_per_block: int = 2, groups: int = 32, ff_gate_width: int = 32) -> Dict[str, List[List[int]]]:
    width, depth = [], []

    def block(n_layers, heads, has_attn, half_depth=True):
        for i in range(n_layers):
            width.append([groups])
            depth.append([1] if (half_depth and i == n_layers - 1) else [0])
        if has_attn:
            for i in range(n_layers):
                width.append([heads, heads, ff_gate_width])
                depth.append([1] if (half_depth and i == n_layers - 1) else [0])

    for i, h in enumerate(num_heads):
        block(layers_per_block, h, down_has_attn[i])
    # mid block: resnet, resnet, attention -- width gates only (blocks.py:2554-2736)
    width.extend([[groups], [groups], [num_heads[-1], num_heads[-1], ff_gate_width]])
    depth.extend([[0], [0], [0]])
    for i, h in enumerate(reversed(num_heads)):
        block(layers_per_block + 1, h, list(reversed(down_has_attn))[i])
    return {"width": width, "depth": depth}


def sd21_gate_structure():
    return gate_structure()

"""Drop-in `HyperStructure` (reference: pdm/models/hypernet/hypernet.py:28) -- the architecture predictor.

Same constructor, parameter names (`mh_fc.{i}.weight/bias`, or `arch` for single_arch_param) and
methods. The 70 width Linears + 1 depth Linear of the reference (hypernet.py:72-79: 71 tiny GEMMs and a
cat) are evaluated as ONE [B,768] x [768,1620] product over the row-concatenated weights; per-Linear
orthogonal initialisation (hypernet.py:58-63) is preserved because the parameters stay separate tensors.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.parametrizations import weight_norm

from ._mixin import ConfigModelMixin
from .quantizer import hard_concrete


class _PackedLinear(torch.autograd.Function):
    """y = x W^T + b over the row-concatenated Linears on the fp32 K9 kernels (csrc/smallops.cu): no cuBLAS, fixed
    summation order; backward gives dW / db (split back to the 71 Linears by autograd's cat) and dx."""

    @staticmethod
    def forward(ctx, x, W, b):
        from . import kernels as K
        x32 = x.detach().to(torch.float32).contiguous()
        W32 = W.detach().to(torch.float32).contiguous()
        b32 = b.detach().to(torch.float32).contiguous() if b is not None else None
        B, Kd = x32.shape
        N = W32.shape[0]
        y = torch.empty(B, N, device=x32.device, dtype=torch.float32)
        K.linear_f32_fwd(x32, W32, b32, y, B, Kd, N)
        ctx.save_for_backward(x32, W32)
        ctx.has_bias, ctx.x_dtype, ctx.w_dtype = b is not None, x.dtype, W.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import kernels as K
        x32, W32 = ctx.saved_tensors
        B, Kd = x32.shape
        N = W32.shape[0]
        dy = dy.to(torch.float32).contiguous()
        dW = torch.empty_like(W32) if ctx.needs_input_grad[1] else None
        db = torch.empty(N, device=dy.device, dtype=torch.float32) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        dx = torch.empty_like(x32) if ctx.needs_input_grad[0] else None
        if B > 0 and (dW is not None or dx is not None):
            if dW is None and db is not None:
                dW = torch.empty_like(W32)
            K.linear_f32_bwd(x32, W32, dy, dx, dW, db, B, Kd, N)
        return (dx.to(ctx.x_dtype) if dx is not None else None,
                dW.to(ctx.w_dtype) if (dW is not None and ctx.needs_input_grad[1]) else None, db)


class HyperStructure(ConfigModelMixin, nn.Module):
    def __init__(self, structure, input_dim=768, wn_flag=True, linear_bias=False, single_arch_param=False):
        super().__init__()
        self.register_to_config(structure=structure, input_dim=input_dim, wn_flag=wn_flag, linear_bias=linear_bias,
                                single_arch_param=single_arch_param)
        self.structure, self.input_dim, self.linear_bias, self.wn_flag = structure, input_dim, linear_bias, wn_flag
        self.width_list = [w for sub in structure["width"] for w in sub]
        self.depth_list = [d for sub in structure["depth"] for d in sub]
        self.single_arch_param = single_arch_param
        n = sum(self.width_list) + sum(self.depth_list)
        if single_arch_param:
            self.arch = nn.Parameter(torch.randn(1, n))
            self.arch_gs = torch.zeros(1, n)
        else:
            linears = [nn.Linear(input_dim, w, bias=linear_bias) for w in self.width_list]
            linears.append(nn.Linear(input_dim, sum(self.depth_list), bias=linear_bias))
            if wn_flag:
                linears = [weight_norm(l) for l in linears]
            self.mh_fc = nn.ModuleList(linears)
            self.initialize_weights()

    def initialize_weights(self):
        for name, p in self.named_parameters():
            if "weight" in name:
                nn.init.orthogonal_(p)
            elif "bias" in name:
                nn.init.zeros_(p)

    def forward(self, x):
        if self.single_arch_param:
            return self.arch
        return self._forward(x)

    def _forward(self, x):
        w0 = self.mh_fc[0].weight
        if w0.is_cuda:
            x = x.to(w0.device)
        # one fused product instead of 71 (autograd splits the gradient back to the per-Linear parameters)
        W = torch.cat([l.weight for l in self.mh_fc], dim=0)
        b = torch.cat([l.bias for l in self.mh_fc], dim=0) if self.linear_bias else None
        if not W.is_cuda:
            raise RuntimeError("HyperStructure runs on the sm_100a CUDA path only: move the module to a CUDA device "
                               "(there is no CPU fallback)")
        return _PackedLinear.apply(x, W, b)

    def transform_structure_vector(self, inputs):
        """hypernet.py:86-101."""
        nw = sum(self.width_list)
        assert inputs.shape[1] == nw + sum(self.depth_list)
        width_list, s = [], 0
        for w in self.width_list:
            width_list.append(inputs[:, s:s + w])
            s += w
        depth_list = [inputs[:, nw + i] for i in range(sum(self.depth_list))]
        return {"width": width_list, "depth": depth_list}

    @classmethod
    def transform_arch_vector(cls, inputs, structure, force_width_non_zero=False):
        """hypernet.py:103-129."""
        width_list = [w for sub in structure["width"] for w in sub]
        depth_list = [d for sub in structure["depth"] for d in sub]
        nw = sum(width_list)
        assert inputs.shape[1] == nw + sum(depth_list)
        w_list, s = [], 0
        for w in width_list:
            sub = inputs[:, s:s + w]
            if force_width_non_zero:
                tot = hard_concrete(sub).sum(dim=1)
                if not tot.all():
                    ind = tot == 0
                    sub = sub.clone()
                    sub[ind, 0] = sub[ind, 0] + 0.5
            w_list.append(sub)
            s += w
        return {"width": w_list, "depth": [inputs[:, nw + i] for i in range(sum(depth_list))]}

    @classmethod
    def get_random_arch_vector(cls, target_ratio, structure):
        """hypernet.py:131-153."""
        width_list = [w for sub in structure["width"] for w in sub]
        depth_list = [d for sub in structure["depth"] for d in sub]
        vecs = []
        for w in width_list:
            v = torch.zeros(1, w)
            v[0, torch.randperm(w)[: int(target_ratio * w)]] = 0.9
            vecs.append(v)
        for _ in range(sum(depth_list)):
            vecs.append(torch.tensor([[0.9]]))
        return torch.cat(vecs, dim=1)

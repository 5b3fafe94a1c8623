"""Import shim used ONLY by tests/golden/make_goldens.py in the build container, to execute the
reference's own Python (read from /root/reference, never copied) without the `diffusers` package.

`install_router_shim()` registers a stub `diffusers` providing ModelMixin / ConfigMixin /
register_to_config -- enough for pdm/models/hypernet/hypernet.py, pdm/models/vq/quantizer.py,
pdm/models/unet/gates.py, pdm/utils/estimation_utils.py and pdm/losses to import unchanged.
TEST INFRASTRUCTURE: nothing under diffusion_pruning_b200/ imports this.
"""
import importlib.util
import sys
import types

import torch.nn as nn

REFERENCE_ROOT = "/root/reference"


def install_router_shim():
    if "diffusers" in sys.modules and getattr(sys.modules["diffusers"], "__aptp_shim__", False):
        return
    d = types.ModuleType("diffusers")
    d.__aptp_shim__ = True
    d.__path__ = []

    class ModelMixin(nn.Module):
        pass

    class ConfigMixin:
        pass

    def register_to_config(fn):
        return fn

    d.ModelMixin = ModelMixin
    d.ConfigMixin = ConfigMixin
    cu = types.ModuleType("diffusers.configuration_utils")
    cu.register_to_config = register_to_config
    cu.ConfigMixin = ConfigMixin
    d.configuration_utils = cu
    sys.modules["diffusers"] = d
    sys.modules["diffusers.configuration_utils"] = cu


def load_reference_module(name: str, relpath: str):
    """Load one reference source file by path under a private module name."""
    spec = importlib.util.spec_from_file_location(name, f"{REFERENCE_ROOT}/{relpath}")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_router_reference():
    """Returns (estimation_utils, hypernet, quantizer, contrastive_loss, resource_loss) reference modules."""
    install_router_shim()
    # minimal `pdm` package skeleton so the reference's absolute imports resolve to the real files
    for pkg in ("pdm", "pdm.utils", "pdm.models", "pdm.losses"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    est = load_reference_module("pdm.utils.estimation_utils", "pdm/utils/estimation_utils.py")
    hyp = load_reference_module("pdm.models.hypernet_ref", "pdm/models/hypernet/hypernet.py")
    vq = load_reference_module("pdm.models.vq_ref", "pdm/models/vq/quantizer.py")
    cl = load_reference_module("pdm.losses.contrastive_loss", "pdm/losses/contrastive_loss.py")
    rl = load_reference_module("pdm.losses.resource_loss", "pdm/losses/resource_loss.py")
    gates = load_reference_module("pdm.models.unet_gates_ref", "pdm/models/unet/gates.py")
    return {"estimation_utils": est, "hypernet": hyp, "quantizer": vq, "contrastive_loss": cl, "resource_loss": rl,
            "gates": gates}

"""Minimal stand-in for diffusers' ModelMixin/ConfigMixin surface that the reference's callers use on
the hypernet and quantizer: `.config`, `register_to_config`, `save_pretrained`, `from_pretrained`
(pdm/training/trainer.py:253-313 saves `hypernet/` and `quantizer/` this way). File names follow
diffusers (config.json + diffusion_pytorch_model.safetensors) so checkpoints stay interchangeable."""
from __future__ import annotations

import json
import os
from types import SimpleNamespace

import torch

CONFIG_NAME = "config.json"
WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"


class ConfigModelMixin:
    def register_to_config(self, **kwargs):
        self._config_dict = dict(kwargs)
        self._config_dict["_class_name"] = type(self).__name__

    @property
    def config(self):
        return SimpleNamespace(**self._config_dict)

    def save_pretrained(self, save_directory: str, **unused):
        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, CONFIG_NAME), "w") as f:
            json.dump(self._config_dict, f, indent=2, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o))
        from safetensors.torch import save_file
        sd = {k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()}
        save_file(sd, os.path.join(save_directory, WEIGHTS_NAME))

    @classmethod
    def from_pretrained(cls, directory: str, subfolder: str = None, **overrides):
        if subfolder:
            directory = os.path.join(directory, subfolder)
        with open(os.path.join(directory, CONFIG_NAME)) as f:
            cfg = json.load(f)
        cfg.pop("_class_name", None)
        cfg.update(overrides)
        model = cls(**cfg)
        from safetensors.torch import load_file
        model.load_state_dict(load_file(os.path.join(directory, WEIGHTS_NAME)))
        return model

"""Import the UNMODIFIED reference `models/*` from /root/reference in THIS container.

TEST INFRASTRUCTURE: used only by oracle/make_golden.py to generate tests/golden/*; /root/reference does not
exist on the GPU box and nothing at test/bench time imports this module.

The reference cannot be imported as-is here (transformers 5.5 rejects a removed TrainingArguments kwarg in
config.py:100; preprocessing/__init__.py pulls srtm/geopandas; ...).  The shim does not touch reference source:
  1. `config`      <- exec of /root/reference/config.py with a permissive TrainingArguments stand-in;
  2. `preprocessing` <- a package exposing only preprocessing/geo_utils.py and preprocessing/utils.py;
  3. sys.path gets /root/reference so that `models.*` resolves to the reference files;
  4. `device_redirect()` lets ProtoRefiner.forward's hard-coded 'cuda' (proto_refiner.py:172-230) run on the CPU.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

REF = os.environ.get("PIGEON_REFERENCE", "/root/reference")


class _Args:
    """Stand-in for transformers.TrainingArguments: keeps the kwargs as attributes (the reference only reads
    per_device_*_batch_size etc. from them, training/train_eval_loop.py:55,187)."""

    def __init__(self, *a, **kw):
        self.__dict__.update(kw)


def install():
    if not os.path.isdir(REF):
        raise RuntimeError(f"{REF} not present: golden generation only runs in the authoring container")
    if "config" not in sys.modules or getattr(sys.modules["config"], "__pigeon_shim__", False) is False:
        cfg = types.ModuleType("config")
        src = open(os.path.join(REF, "config.py")).read()
        g = cfg.__dict__
        import transformers
        real = transformers.TrainingArguments
        transformers.TrainingArguments = _Args
        try:
            exec(compile(src, os.path.join(REF, "config.py"), "exec"), g)
        finally:
            transformers.TrainingArguments = real
        cfg.__pigeon_shim__ = True
        sys.modules["config"] = cfg
    if "preprocessing" not in sys.modules or not getattr(sys.modules["preprocessing"], "__pigeon_shim__", False):
        pkg = types.ModuleType("preprocessing")
        pkg.__path__ = []
        pkg.__pigeon_shim__ = True
        sys.modules["preprocessing"] = pkg
        for name in ("geo_utils", "utils"):
            spec = importlib.util.spec_from_file_location(f"preprocessing.{name}", os.path.join(REF, "preprocessing", f"{name}.py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[f"preprocessing.{name}"] = mod
            spec.loader.exec_module(mod)
            for k, v in mod.__dict__.items():
                if not k.startswith("_"):
                    setattr(pkg, k, v)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import datasets
    datasets.config.TORCHVISION_AVAILABLE = False
    import models  # noqa: F401  (the reference package)
    return sys.modules["models"]


@contextlib.contextmanager
def device_redirect():
    """Run reference code that hard-codes device 'cuda' on the CPU, without editing it: 'cuda' device
    arguments of torch.tensor / Tensor.to become 'cpu' for the duration of the context."""
    import torch
    real_tensor, real_to = torch.tensor, torch.Tensor.to

    def fix(dev):
        if isinstance(dev, str) and dev.startswith("cuda"):
            return "cpu"
        if isinstance(dev, torch.device) and dev.type == "cuda":
            return torch.device("cpu")
        return dev

    def tensor(*a, **kw):
        if "device" in kw:
            kw["device"] = fix(kw["device"])
        return real_tensor(*a, **kw)

    def to(self, *a, **kw):
        a = tuple(fix(x) for x in a)
        if "device" in kw:
            kw["device"] = fix(kw["device"])
        return real_to(self, *a, **kw)

    torch.tensor, torch.Tensor.to = tensor, to
    try:
        yield
    finally:
        torch.tensor, torch.Tensor.to = real_tensor, real_to


@contextlib.contextmanager
def chdir(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)

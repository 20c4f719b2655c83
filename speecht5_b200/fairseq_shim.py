"""Minimal stand-in for the fairseq registries/base classes used by the SpeechT5 plugin surface, used ONLY when
`import fairseq` fails (it does in the build container: omegaconf<2.1 / numpy<1.24 pins). With a real fairseq on the
path the same decorators come from fairseq itself (fairseq/models/__init__.py:99,150; fairseq/tasks/__init__.py:49;
fairseq/criterions/__init__.py) and `--user-dir speecht5_b200` registers task/criterion/model under the reference's
names."""
import torch.nn as nn

import sys as _sys

try:  # pragma: no cover - exercised only where fairseq is installed
    if "fairseq" in _sys.modules and not hasattr(_sys.modules["fairseq"], "__file__"):
        raise ImportError("a test stub package named fairseq is loaded (oracle/ref_loader.py), not fairseq itself")
    from fairseq.models import (FairseqEncoderDecoderModel, register_model,  # noqa: F401
                                register_model_architecture)
    from fairseq.tasks import LegacyFairseqTask, register_task  # noqa: F401
    from fairseq.criterions import FairseqCriterion, register_criterion  # noqa: F401
    from fairseq import metrics  # noqa: F401
    HAVE_FAIRSEQ = True
except Exception:  # noqa: BLE001
    HAVE_FAIRSEQ = False

    class _Metrics:
        """fairseq.metrics stand-in: keeps the last value per key (log_scalar) and the derived-meter callables."""

        def __init__(self):
            self.scalars, self.derived = {}, {}

        def reset(self):
            self.scalars, self.derived = {}, {}

        def log_scalar(self, key, value, weight=1, priority=10, round=None):
            self.scalars[key] = value

        def log_derived(self, key, fn, priority=20):
            self.derived[key] = fn

    metrics = _Metrics()
    MODEL_REGISTRY, ARCH_MODEL_REGISTRY, ARCH_CONFIG_REGISTRY = {}, {}, {}
    TASK_REGISTRY, CRITERION_REGISTRY = {}, {}

    def register_model(name, dataclass=None):
        def wrap(cls):
            if name in MODEL_REGISTRY:
                raise ValueError(f"Cannot register duplicate model ({name})")
            MODEL_REGISTRY[name] = cls
            return cls
        return wrap

    def register_model_architecture(model_name, arch_name):
        def wrap(fn):
            if model_name not in MODEL_REGISTRY:
                raise ValueError(f"Cannot register model architecture for unknown model type ({model_name})")
            ARCH_MODEL_REGISTRY[arch_name] = MODEL_REGISTRY[model_name]
            ARCH_CONFIG_REGISTRY[arch_name] = fn
            return fn
        return wrap

    def register_task(name, dataclass=None):
        def wrap(cls):
            TASK_REGISTRY[name] = cls
            return cls
        return wrap

    def register_criterion(name, dataclass=None):
        def wrap(cls):
            CRITERION_REGISTRY[name] = cls
            cls.__dataclass = dataclass
            return cls
        return wrap

    class FairseqEncoderDecoderModel(nn.Module):
        def __init__(self, encoder, decoder):
            super().__init__()
            self.encoder, self.decoder = encoder, decoder

        def set_num_updates(self, num_updates):
            for m in self.modules():
                if hasattr(m, "set_num_updates") and m is not self:
                    m.set_num_updates(num_updates)

    class FairseqCriterion(nn.Module):
        def __init__(self, task):
            super().__init__()
            self.task = task

    class LegacyFairseqTask:
        def __init__(self, args):
            self.args = args

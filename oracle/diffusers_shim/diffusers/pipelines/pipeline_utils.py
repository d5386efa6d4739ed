"""diffusers.pipelines.pipeline_utils.DiffusionPipeline: the members pipeline_chronoedit.py touches ([diffusers-mem] 0.35.2):
register_modules (:175), _execution_device (:217), progress_bar (:694), maybe_free_model_hooks (:807), to()."""
import contextlib

import torch


class _Bar:
    def __init__(self, total):
        self.total, self.n = total, 0

    def update(self, k=1):
        self.n += k


class DiffusionPipeline:
    def register_modules(self, **kwargs):
        self._modules_registered = dict(getattr(self, "_modules_registered", {}))
        for k, v in kwargs.items():
            self._modules_registered[k] = v
            setattr(self, k, v)

    @property
    def components(self):
        return dict(self._modules_registered)

    @property
    def _execution_device(self):
        for m in self._modules_registered.values():
            if isinstance(m, torch.nn.Module):
                try:
                    return next(m.parameters()).device
                except StopIteration:
                    continue
        return torch.device("cpu")

    @property
    def device(self):
        return self._execution_device

    def to(self, *args, **kwargs):
        for m in self._modules_registered.values():
            if isinstance(m, torch.nn.Module):
                m.to(*args, **kwargs)
        return self

    @contextlib.contextmanager
    def progress_bar(self, iterable=None, total=None):
        yield _Bar(total)

    def maybe_free_model_hooks(self):
        return None

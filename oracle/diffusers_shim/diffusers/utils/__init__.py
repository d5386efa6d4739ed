import logging as _logging

USE_PEFT_BACKEND = False


class _Logging:
    @staticmethod
    def get_logger(name):
        return _logging.getLogger(name)


logging = _Logging()


def scale_lora_layers(model, weight):
    return None


def unscale_lora_layers(model, weight=None):
    return None


def deprecate(*args, **kwargs):  # diffusers.utils.deprecate: warns only
    return None


def is_scipy_available():
    return False


def is_ftfy_available():
    return False


def is_torch_xla_available():
    return False


def replace_example_docstring(example_docstring):
    def wrap(fn):
        return fn

    return wrap

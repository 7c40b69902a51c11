import functools
import inspect


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # deepcopy/pickle probe dunder names: must be AttributeError
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class ConfigMixin:
    def register_to_config(self, **kw):
        if "_internal_dict" not in self.__dict__:
            object.__setattr__(self, "_internal_dict", _AttrDict())
        self._internal_dict.update(kw)

    @property
    def config(self):
        return self._internal_dict


def register_to_config(init):
    sig = inspect.signature(init)

    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        # nn.Module.__setattr__ needs __init__ to have run for Module subclasses; store via object.__setattr__
        if "_internal_dict" not in self.__dict__:
            object.__setattr__(self, "_internal_dict", _AttrDict())
        self._internal_dict.update(cfg)
        init(self, *args, **kwargs)

    return wrapper

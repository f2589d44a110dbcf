import functools
import inspect


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        if not hasattr(self, "_internal_dict"):
            self._internal_dict = FrozenDict()
        d = dict(self._internal_dict)
        d.update(kwargs)
        self._internal_dict = FrozenDict(d)

    @property
    def config(self):
        return self._internal_dict


def register_to_config(init):
    @functools.wraps(init)
    def inner_init(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = {n: p.default for i, (n, p) in enumerate(sig.parameters.items()) if i > 0}
        new = {}
        for a, name in zip(args, params.keys()):
            new[name] = a
        new.update({k: kwargs.get(k, d) for k, d in params.items() if k not in new})
        init(self, *args, **kwargs)
        getattr(self, "register_to_config")(**new)
    return inner_init

import logging as _logging
from collections import OrderedDict


def deprecate(*args, **kwargs):
    return None


class BaseOutput(OrderedDict):
    def __post_init__(self):
        pass

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


class logging:  # noqa: N801 (mirrors diffusers.utils.logging)
    @staticmethod
    def get_logger(name=None):
        return _logging.getLogger(name)

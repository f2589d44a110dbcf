class FromOriginalControlNetMixin:
    pass


class LoraLoaderMixin:
    pass

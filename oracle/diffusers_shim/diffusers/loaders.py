class FromOriginalControlNetMixin:
    pass


class LoraLoaderMixin:
    pass


class UNet2DConditionLoadersMixin:
    pass

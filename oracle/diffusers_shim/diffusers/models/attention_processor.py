class AttnProcessor:
    pass


class AttnAddedKVProcessor:
    pass


AttentionProcessor = AttnProcessor
ADDED_KV_ATTENTION_PROCESSORS = (AttnAddedKVProcessor,)
CROSS_ATTENTION_PROCESSORS = (AttnProcessor,)

"""Import-path shim for /root/reference/model/ctrl_helper.py (inference.py:13).

`ControlNetHelper` wraps the condition extractors (depth / canny / pose estimators) and the CLIP prompt encoder that run
ONCE per generation, before the denoising loop -- SURVEY.md section 8 puts them outside the hot path this repository
implements, so the name resolves but construction fails loudly instead of silently doing something else.
"""


class ControlNetHelper:
    def __init__(self, *_args, **_kwargs):
        raise NotImplementedError("model.ctrl_helper.ControlNetHelper (condition extractors / prompt encoding) is outside the "
                                  "scope of ctrl_adapter_b200; use the reference's helper to produce the control images and "
                                  "ControlNet prompt embeddings, then feed them to the B200 modules / loops")

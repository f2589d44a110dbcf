"""Import-path shim for /root/reference/model/ctrl_helper.py (inference.py:13): the helper's pipeline-facing half
(prepare_images for any batch size, encode_controlnet_prompt on a supplied / locally loaded CLIP text encoder,
_get_add_time_ids) is implemented in ctrl_adapter_b200.helper; the condition extractors fail loudly (out of scope)."""
from ctrl_adapter_b200.helper import ControlNetHelper  # noqa: F401

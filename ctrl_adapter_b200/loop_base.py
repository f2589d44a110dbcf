"""What the three denoising loops share: the per-step control schedule (``controlnet_keep``), eager / CUDA-graph stepping.

Reference: the ``controlnet_keep`` list and ``cond_scale = controlnet_conditioning_scale * controlnet_keep[i]`` of
/root/reference/sdxl/pipelines/sdxl_controlnet_adapter_pipeline.py:1207-1211, :1296-1302 (identical code in
i2vgen_xl/...pipeline.py:844-850, :921-927 and svd/...pipeline.py:624-628, :652-658).  Every shipped inference script
passes ``--control_guidance_end`` < 1, so the steps past that fraction run with ``cond_scale == 0``: the reference then
drops the down-block residuals (:1346 ``if cond_scale == 0``).  A loop body is therefore parameterised by the step's
conditioning scale; one CUDA graph is captured per distinct scale value (normally two: the user's scale and 0) and the
per-step scalars (t, sigma, ...) live in device rows that a 16-byte copy refreshes before each replay.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import torch

Scale = Union[float, tuple]


def controlnet_keep(num_steps: int, start: Sequence[float], end: Sequence[float]) -> List[List[float]]:
    """keeps[i][k] for ControlNet k at step i -- the reference's list comprehension, verbatim semantics."""
    return [[1.0 - float(i / num_steps < s or (i + 1) / num_steps > e) for s, e in zip(start, end)]
            for i in range(num_steps)]


class DenoiseLoopBase:
    num_inference_steps: int

    def _init_control(self, conditioning_scale, control_guidance_start, control_guidance_end, num_nets: int):
        """Mirrors the reference's list alignment of start / end / scale (sdxl pipeline :1090-1105)."""
        s, e = control_guidance_start, control_guidance_end
        if not isinstance(s, (list, tuple)) and isinstance(e, (list, tuple)):
            s = len(e) * [s]
        elif not isinstance(e, (list, tuple)) and isinstance(s, (list, tuple)):
            e = len(s) * [e]
        elif not isinstance(s, (list, tuple)) and not isinstance(e, (list, tuple)):
            s, e = num_nets * [s], num_nets * [e]
        self._keep = controlnet_keep(self.num_inference_steps, list(s), list(e))
        self._multi = num_nets > 1
        if self._multi and not isinstance(conditioning_scale, (list, tuple)):
            conditioning_scale = [conditioning_scale] * num_nets
        self._scale = ([float(c) for c in conditioning_scale] if isinstance(conditioning_scale, (list, tuple))
                       else float(conditioning_scale))
        self._graphs: Dict[Scale, torch.cuda.CUDAGraph] = {}

    def control_scale(self, i: int) -> Scale:
        """cond_scale of step i: a float for one ControlNet, a tuple for Multi-ControlNet (reference :1296-1302)."""
        keeps = self._keep[i]
        if self._multi:
            return tuple(c * k for c, k in zip(self._scale, keeps))
        c = self._scale[0] if isinstance(self._scale, list) else self._scale
        return c * keeps[0]

    # ---- subclass contract --------------------------------------------------------------------------------
    def _state(self) -> List[torch.Tensor]:  # tensors a body run mutates (restored after capture warm-ups)
        raise NotImplementedError

    def _load_step(self, i: int) -> None:    # refresh the device rows of step i
        raise NotImplementedError

    def _body(self, scale: Scale) -> None:   # one denoising iteration at conditioning scale `scale`
        raise NotImplementedError

    # ---- stepping -----------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, i: Optional[int] = None):
        """One denoising iteration through the modules' public forward()s (eager launches)."""
        i = self.step_index if i is None else i
        self._load_step(i)
        self._body(self.control_scale(i))
        self.step_index = i + 1
        return self.latents

    @torch.no_grad()
    def capture(self, warmup: int = 2, scale: Optional[Scale] = None):
        """Record the step body at conditioning scale `scale` (default: step 0's) into a CUDA graph, after `warmup` eager
        runs that pack weights and set kernel attributes.  Loop state is restored afterwards."""
        scale = self.control_scale(0) if scale is None else scale
        saved = [t.clone() for t in self._state()]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body(scale)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._body(scale)
        self._graphs[scale] = g
        for t, v in zip(self._state(), saved):
            t.copy_(v)
        return g

    @torch.no_grad()
    def step_graph(self, i: Optional[int] = None):
        i = self.step_index if i is None else i
        scale = self.control_scale(i)
        if scale not in self._graphs:
            self.capture(scale=scale)
        self._load_step(i)
        self._graphs[scale].replay()
        self.step_index = i + 1
        return self.latents

    @torch.no_grad()
    def run(self, use_graph: bool = True):
        for i in range(self.num_inference_steps):
            (self.step_graph if use_graph else self.step)(i)
        return self.latents

"""Inference service shape of the reference's evaluator (TBIEvaluator.py:210-258, DisplayInference.py:158-206), SURVEY 8f-3.

The reference forks one process per test image, re-loads the SavedModel in each, runs ``prob, _ = SegNet(testX)`` on a batch
of ONE and post-processes with numpy.  Here one resident model serves batches of requests and the post-processing is the
epilogue kernel: ``probOut = prob[..., -1]`` and ``probO = 1 - p0 - 0.5 p1 + p2`` (:239-252), optionally behind the brain-mask
pre-pass (:225-231: a second model's rounded first class zeroes the input planes).

    from ultrasound_modeling_b200.evaluator import Evaluator
    ev = Evaluator(net)                                  # net: VisionTransformer (or anything with ._forward_logits / TBI ResNest)
    out = ev(testX)                                      # testX [N,256,80,10] -> dict(prob, probOut, probO) device tensors
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib


class Evaluator:
    def __init__(self, model, brain_mask_model=None, max_batch: int = 64):
        self.model, self.mask_model, self.max_batch = model, brain_mask_model, max_batch
        self.L = _lib.lib()

    def _logits(self, model, x):
        if hasattr(model, "forward_logits"):                    # VisionTransformer (CUDA-graph replay per batch shape)
            return model.forward_logits(x)[0]
        if hasattr(model, "_forward_logits"):
            return model._forward_logits(x)[0]
        if hasattr(model, "engine"):                            # TBI_ResNest.ResNest: probabilities -> log (softmax(log p) == p)
            return torch.log(model.predict(x, dropout_masks=None).clamp_min(1e-30))
        raise TypeError("Evaluator: unsupported model object")

    def _maps(self, z):
        n, h, w, nc = z.shape
        z = z.to(torch.float32).contiguous()
        prob = torch.empty_like(z)
        prob_out = torch.empty(n, h, w, dtype=torch.float32, device=z.device)
        prob_o = torch.empty(n, h, w, dtype=torch.float32, device=z.device)
        _lib.check(self.L.tbi_softmax_prob_maps(n * h * w, nc, z.data_ptr(), prob.data_ptr(), prob_out.data_ptr(), prob_o.data_ptr(),
                                                torch.cuda.current_stream(z.device).cuda_stream), "softmax_prob_maps")
        return prob, prob_out, prob_o

    def __call__(self, testX):
        """any number of samples; served in batches of max_batch -> dict(prob [N,H,W,C], probOut [N,H,W], probO [N,H,W])"""
        dev = getattr(self.model, "device", None) or self.model.engine.device
        x = torch.as_tensor(np.asarray(testX) if not torch.is_tensor(testX) else testX).to(device=dev, dtype=torch.float32).contiguous()
        outs = []
        for i in range(0, x.shape[0], self.max_batch):
            xb = x[i:i + self.max_batch].clone()
            if self.mask_model is not None:
                mp, _, _ = self._maps(self._logits(self.mask_model, xb))
                n, h, w, c = xb.shape
                _lib.check(self.L.tbi_apply_brain_mask(n * h * w, mp.shape[3], c, mp.data_ptr(), xb.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                           "apply_brain_mask")
            outs.append(self._maps(self._logits(self.model, xb)))
        cat = lambda k: torch.cat([o[k] for o in outs], 0)
        return dict(prob=cat(0), probOut=cat(1), probO=cat(2))

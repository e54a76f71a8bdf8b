"""Shared machinery of the drop-in model classes: an ``nn.Module`` that keeps the reference's layer
tree (so ``state_dict`` keys, ``.to()``, ``.eval()``, ``load_state_dict`` and checkpoints behave
identically) but whose eval-mode forward on CUDA tensors runs in libdfs_b200.so.

* eval mode + CUDA input  -> native scorer (handle built lazily from the current weights and rebuilt
  whenever a parameter/buffer is modified in place, re-loaded or moved).
* train mode              -> the inherited PyTorch layers (training is out of scope of the engine;
  the extension has no backward).
* eval mode + CPU input   -> RuntimeError: there is deliberately no CPU fallback on the scoring path.
"""
import os
import sys

import torch
import torch.nn as nn

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)


class NativeBackedModule(nn.Module):
    # "fp16" (default): tensor-core path, fp16 operands / fp32 accumulation.  "fp32": the full-fp32 CUDA-core kernels, for evaluations
    # where the rank order of near-equal scores matters.  "split": the 2D-CNN's accurate tensor-core mode (every operand as fp16 value +
    # residual, ~0.4x the fp16 rate, fp32-class scores); the 1D-CNN and the CAE have no such mode and take their fp32 kernels for it
    # (at least as accurate).  Set it on the instance (or class) before the first CUDA forward, or export DFS_B200_PRECISION.
    precision = None
    HAS_SPLIT = False            # CNN2D overrides

    def __init__(self):
        super().__init__()
        self._native = None
        self._native_key = None

    def _precision(self):
        p = self.precision or os.environ.get("DFS_B200_PRECISION", "fp16")
        return "fp32" if (p == "split" and not self.HAS_SPLIT) else p

    def _make_scorer(self, state_dict, device_index):  # pragma: no cover - overridden
        raise NotImplementedError

    def _weights_key(self, device):
        tensors = list(self.parameters()) + list(self.buffers())
        return (str(device),) + tuple((t.data_ptr(), t._version) for t in tensors)

    def native(self, device):
        """The native scorer for the current weights on `device` (cached)."""
        key = self._weights_key(device) + (self._precision(),)
        if self._native is None or self._native_key != key:
            if self._native is not None:
                self._native.close()
            sd = {k: v.detach().float().cpu() for k, v in self.state_dict().items()}
            self._native = self._make_scorer(sd, device.index if device.index is not None else torch.cuda.current_device())
            self._native_key = key
        return self._native

    def _use_native(self, x):
        if self.training:
            return False
        if not x.is_cuda:
            raise RuntimeError(f"{type(self).__name__}: eval-mode scoring runs on CUDA only (dfs_b200 has no CPU fallback); "
                               "move the model and the features to a CUDA device")
        return True

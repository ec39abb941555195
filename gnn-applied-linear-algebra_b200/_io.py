"""Input/output placement shared by the drop-in layers: inputs may live on the host (the
reference's users hold CPU tensors) or on the GPU; compute always happens on the GPU and
results are returned where the inputs came from."""
import torch

from . import _runtime as rt


class Placement:
    def __init__(self, *tensors):
        self.device = rt.compute_device(*tensors)
        first = next((t for t in tensors if isinstance(t, torch.Tensor)), None)
        self.host = first is not None and not first.is_cuda

    def up(self, t, dtype=None):
        return rt.to_device(t, self.device, dtype)

    def down(self, t):
        if t is None or not self.host:
            return t
        return t.cpu()


def float_dtype(*tensors):
    dt = None
    for t in tensors:
        if t is None or not t.dtype.is_floating_point:
            continue
        dt = t.dtype if dt is None else torch.promote_types(dt, t.dtype)
    dt = dt or torch.float32
    rt.suffix(dt)  # raises for anything but float32 / float64
    return dt

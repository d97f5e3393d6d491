"""TEST INFRASTRUCTURE: a torch-backed stand-in for exactly the ``tf`` calls ct_pvae_b200/tf_bridge.py makes, so that
the TensorFlow binding's own code (DLPack capsules, plan lookup, device syncs, custom_gradient wiring) can execute on
the GPU in an image where TensorFlow cannot be installed.  It is NOT TensorFlow: "tensors" are torch tensors,
``custom_gradient`` is a torch.autograd.Function, ``py_function`` calls the function eagerly."""
import contextlib
import types

import torch
from torch.utils import dlpack as _dlpack

float32, uint8, float64 = torch.float32, torch.uint8, torch.float64
_device = []


@contextlib.contextmanager
def device(name):
    _device.append(torch.device(str(name)))
    try:
        yield
    finally:
        _device.pop()


def _dev():
    return _device[-1] if _device else torch.device("cuda", torch.cuda.current_device())


def convert_to_tensor(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x)


def zeros(shape, dtype=float32):
    return torch.zeros([int(v) for v in shape], dtype=dtype, device=_dev())


def cast(x, dtype):
    return x.to(dtype)


def identity(x):
    return x.clone(memory_format=torch.contiguous_format)     # TensorFlow tensors are always dense row-major


def transpose(x, perm):
    return x.permute(*perm).contiguous()                      # tf.transpose materialises its result


def py_function(func, inp, Tout):
    out = func(*inp)
    return out.to(Tout) if isinstance(out, torch.Tensor) else out


def custom_gradient(f):
    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, *args):
            with torch.enable_grad():
                out, grad_fn = f(*[a.detach() if isinstance(a, torch.Tensor) else a for a in args])
            ctx.grad_fn_ = grad_fn
            return out

        @staticmethod
        def backward(ctx, dout):
            g = ctx.grad_fn_(dout.contiguous())
            return g if isinstance(g, tuple) else (g,)

    return lambda *args: _Fn.apply(*args)


experimental = types.SimpleNamespace(dlpack=types.SimpleNamespace(to_dlpack=_dlpack.to_dlpack, from_dlpack=_dlpack.from_dlpack))

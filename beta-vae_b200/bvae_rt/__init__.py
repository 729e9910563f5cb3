"""Runtime glue of the B200 Beta-VAE-SE path: ctypes binding (lib), autograd wrappers (ops),
data-parallel gradient reducer (dp), fused optimizer (optim) and the GPU self-check."""
from . import lib  # noqa: F401

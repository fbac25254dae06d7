"""Host-side binding of libfusionfpn.so: ctypes loader (lib), tensor-level ops (ops) and the autograd
Functions (functional) that the reference-shaped nn.Modules under models/ call; trainer (data-parallel step) and
checkpoint (reference .ckpt layout, evaluation ensemble) are imported on demand."""
from . import lib, ops, functional  # noqa: F401
from .functional import set_compute_dtype, get_compute_dtype  # noqa: F401
from .ops import set_conv_impl  # noqa: F401

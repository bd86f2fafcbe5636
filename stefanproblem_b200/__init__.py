"""stefanproblem_b200: B200-native (sm_100a) replacement of StefanProblem's hot path.

Host-side mirror of the reference's Julia call surface (module names follow the
reference: InteriorPenalty, LocalDG, DG1D, FiniteDifference) over the C ABI of
libstefan_b200.so.  PyTorch is used for device memory, streams and
torch.distributed only.
"""
from ._lib import StefanError, lib  # noqa: F401

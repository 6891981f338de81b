"""radar_sounder_crw_b200 -- B200-native (sm_100a) CRW hot path.

Drop-in for the contrastive-random-walk hot path of jdalcorso/radar-sounder-crw:

* training loss    ``CRW(encoder, tau, pos_embed).forward(seq) -> (loss, A)``      (reference src/model.py)
* label propagation ``propagate(...)``, ``LabelPropVOS_CRW(cfg).predict(...)``,
  ``batched_affinity(...)``                                    (reference src/utils.py, src/imported/*)

Everything after the encoder runs in hand-written CUDA kernels behind the C ABI declared in
``include/crw_b200.h`` (``lib/libcrw_b200.so``), exposed to PyTorch as ``torch.library`` custom ops
in the ``crw_b200::`` namespace.  There is no CPU or PyTorch fallback: the library is loaded on first
use, and the first op call raises when it has not been built (``_lib.lib()``).
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401
from . import parallel  # noqa: F401
from . import optim  # noqa: F401
from .model import CRW  # noqa: F401
from .labelprop import LabelPropVOS_CRW  # noqa: F401
from .maskedatt import MaskedAttention, batched_affinity  # noqa: F401
from .utils import propagate, propagate_batch, pos_embed, ndiag_matrix, create_model, seed_labels, fuse_reversed  # noqa: F401
from .dataset import RGDataset, trim_miguel  # noqa: F401
from .encoder import CNN, Resnet, UNetEncoder  # noqa: F401

__version__ = "0.1.0"

"""B200-native (sm_100a) spectral convolution for the FNO-family backbones of
amazon-science/dlwp-benchmark: drop-in ``FNO`` / ``TFNO`` / ``SpectralConv`` (neuralop API,
as used by src/nsbench/models/fno/fno.py and src/dlwpbench/models/fno/fno.py) and ``AFNO2D``
(src/*/models/fourcastnet/fourcastnet.py).  All arithmetic on the hot path runs in the
hand-written CUDA kernels of ``libspectral_b200.so``; there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .plan import Plan, get_plan  # noqa: F401
from .spectral_conv import SpectralConv  # noqa: F401
from .fno import FNO, TFNO, FNOBlocks, MLP  # noqa: F401
from .afno import AFNO2D  # noqa: F401
from .fourcastnet import AFNONet, AFNONetNS, Block, Mlp, PatchEmbed  # noqa: F401
from .shim import install_neuralop_shim, patch_fourcastnet  # noqa: F401
from .rollout import Rollout, sequence_forward, dlwp_sequence_forward, DLWPRollout  # noqa: F401

__version__ = "0.1.0"

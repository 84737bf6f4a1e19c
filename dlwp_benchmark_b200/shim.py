"""Make the UNMODIFIED reference wrappers import and run on this implementation.

* ``install_neuralop_shim()`` registers ``neuralop`` / ``neuralop.models`` in ``sys.modules`` so
  that ``from neuralop.models import FNO, TFNO`` (src/nsbench/models/fno/fno.py:7,
  src/dlwpbench/models/fno/fno.py:7, src/dlwpbench/models/fourcastnet/fourcastnet.py:17)
  resolves to the B200 classes.
* ``patch_fourcastnet(module)`` rebinds ``AFNO2D`` in an imported reference ``fourcastnet``
  module (nsbench resolves the global at Block construction, fourcastnet.py:145; dlwpbench does
  ``eval("AFNO2D")``, fourcastnet.py:173,259) -- see INTEGRATION.md.
"""
from __future__ import annotations

import sys
import types


def install_neuralop_shim(force: bool = False):
    from . import fno, spectral_conv
    if "neuralop" in sys.modules and not force and not getattr(sys.modules["neuralop"], "__b200_shim__", False):
        raise RuntimeError("a real `neuralop` is already imported; pass force=True to shadow it, or inject "
                           "dlwp_benchmark_b200.SpectralConv through neuralop's FNO(SpectralConv=...) hook")
    pkg = types.ModuleType("neuralop")
    pkg.__b200_shim__ = True
    models = types.ModuleType("neuralop.models")
    layers = types.ModuleType("neuralop.layers")
    sc = types.ModuleType("neuralop.layers.spectral_convolution")
    models.FNO, models.TFNO = fno.FNO, fno.TFNO
    models.FNO2d = fno.FNO
    models.TFNO2d = fno.TFNO
    sc.SpectralConv = spectral_conv.SpectralConv
    layers.spectral_convolution = sc
    pkg.models, pkg.layers = models, layers
    pkg.FNO, pkg.TFNO = fno.FNO, fno.TFNO
    sys.modules.update({"neuralop": pkg, "neuralop.models": models, "neuralop.layers": layers,
                        "neuralop.layers.spectral_convolution": sc})
    return pkg


def patch_fourcastnet(module, blocks: bool = True):
    """Swap the reference module's AFNO2D (and, with ``blocks``, its Block / Mlp / PatchEmbed) for the B200 ones: same
    constructors, parameters and forward, so the reference's own ``AFNONet`` builds on them unchanged."""
    from .afno import AFNO2D
    from . import fourcastnet as f
    module.AFNO2D = AFNO2D
    if blocks:
        module.Block, module.Mlp, module.PatchEmbed = f.Block, f.Mlp, f.PatchEmbed
    return module

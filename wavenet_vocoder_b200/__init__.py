# coding: utf-8
"""wavenet_vocoder_b200 — the autoregressive synthesis path of r9y9/wavenet_vocoder
(``WaveNet.incremental_forward`` and the output samplers) as a persistent sm_100a CUDA kernel
behind the reference's own class surface.  See DESIGN.md / INTEGRATION.md."""
from .version import version as __version__
from .wavenet import WaveNet, receptive_field_size

__all__ = ["WaveNet", "receptive_field_size", "__version__"]

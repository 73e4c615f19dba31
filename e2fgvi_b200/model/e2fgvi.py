"""``model.e2fgvi`` — fixed 432x240 variant (reference: model/e2fgvi.py). Discriminator is training-only and
out of scope."""
from .generator import BaseNetwork, Encoder, InpaintGenerator, deconv  # noqa: F401

__all__ = ["InpaintGenerator", "Encoder", "deconv", "BaseNetwork"]

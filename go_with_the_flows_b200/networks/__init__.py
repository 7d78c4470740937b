"""Drop-in replacements for the reference's `lib.networks` modules (same names and signatures)."""
from . import layers, flows, decoders, encoders, models, flow_mixture, losses, optimizers  # noqa: F401

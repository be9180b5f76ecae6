from .neutts import NeuTTSAir

__all__ = ["NeuTTSAir"]

from .neutts import NeuTTS

__all__ = ["NeuTTS"]

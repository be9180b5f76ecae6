"""Back-compat alias package (reference: neuttsair/neutts.py:4-11)."""
from neutts.neutts import NeuTTS


class NeuTTSAir(NeuTTS):
    """Same class under its earlier name; no behaviour of its own."""

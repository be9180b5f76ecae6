"""GPU: the drop-in facade end to end on the B200 engines (synthetic weights, injected tokenizer /
phonemizer): the reference's own smoke assertions (tests/test_neutts.py:55-58), batched inference,
and streaming with the reference's window geometry."""
import warnings

import numpy as np
import pytest
import torch

from oracle import codec_oracle as CO
from oracle import lm_oracle as LO
from oracle import stream_oracle as SO
from tests.helpers import make_codec, make_lm
from tests.test_host_logic import FakePhonemizer, FakeTokenizer

pytestmark = pytest.mark.gpu


class SmallTok(FakeTokenizer):
    """FakeTokenizer squeezed into a 4096-token vocabulary: 1024 speech codes from id 3000."""

    def __init__(self):
        super().__init__(n_speech=1024)


def _tts(max_batch=1, seed=7):
    from neutts import NeuTTS

    cfg = LO.LMConfig.tiny(vocab_size=4096, hidden_size=256, intermediate_size=512, num_layers=2, num_heads=4, num_kv_heads=2)
    w = LO.random_weights(cfg, 3, std=0.05, bf16_round=True)
    lm = make_lm(cfg, w, max_batch=max_batch, max_ctx=2048, max_new=512)
    ccfg = CO.CodecConfig.tiny()
    dec = make_codec(ccfg, CO.random_weights(ccfg, 2), max_batch=max_batch, max_frames=512)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tts = NeuTTS(tokenizer=SmallTok(), phonemizer=FakePhonemizer(), backbone=lm, codec=dec, max_batch=max_batch, seed=seed)
    tts.max_context = 2048
    return tts, ccfg


def test_facade_infer_on_gpu():
    tts, ccfg = _tts()
    ref_codes = torch.arange(40, dtype=torch.int32)
    audio = tts.infer("Testing.", ref_codes, "some reference text")
    # exactly what the reference's own test asserts
    assert isinstance(audio, np.ndarray), type(audio)
    assert len(audio) > 0
    assert not np.isnan(audio).any()
    assert audio.dtype in (np.float32, np.float64), audio.dtype
    assert len(audio) % ccfg.hop == 0
    # the random-weight LM emits mostly non-speech ids; every kept code is a valid codec id, order preserved
    gen = tts._generate_ids([tts._apply_chat_template(ref_codes, "some reference text", "Testing.")])[0]
    codes = tts._ids_to_codes(gen)
    assert len(audio) == ccfg.hop * len(codes)
    again = tts.infer("Testing.", ref_codes, "some reference text")
    assert np.array_equal(audio, again)            # seeded: same tokens, same waveform


def test_facade_batch_matches_single():
    tts1, _ = _tts(max_batch=1)
    tts3, _ = _tts(max_batch=3)
    texts = ["alpha", "beta gamma", "delta"]
    refs = [torch.arange(10 + 5 * i) for i in range(3)]
    rts = ["one", "two words", "three"]
    batch = tts3.infer_batch(texts, refs, rts)
    assert len(batch) == 3 and all(isinstance(b, np.ndarray) and len(b) > 0 for b in batch)
    # slot 0 of the batch draws from the same Philox stream (seed, slot 0, step) as a batch of one
    solo = tts1.infer(texts[0], refs[0], rts[0])
    assert len(solo) == len(batch[0])


def test_facade_streaming_geometry():
    tts, ccfg = _tts()
    hop = ccfg.hop
    tts.hop_length = hop                                     # tiny codec: 16 samples per frame
    tts.streaming_stride_samples = tts.streaming_frames_per_chunk * hop
    ref_codes = list(range(60))
    chunks = list(tts.infer_stream("streaming test sentence", ref_codes, "reference"))
    assert len(chunks) >= 1 and all(isinstance(c, np.ndarray) and c.dtype == np.float32 for c in chunks)
    total = sum(len(c) for c in chunks)
    assert total % hop == 0 and total > 0
    n_frames = total // hop                                  # generated frames that reached the codec
    plan = SO.chunk_plan(len(ref_codes), len(ref_codes) + n_frames, hop=hop)
    full = [c for c in chunks[:-1]] if len(plan) > 1 else []
    assert all(len(c) == tts.streaming_stride_samples for c in full)   # every non-final chunk is 25 frames

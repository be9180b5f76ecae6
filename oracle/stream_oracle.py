"""CPU restatement of the reference's streaming cross-fade (TEST INFRASTRUCTURE ONLY).

``linear_overlap_add`` follows ``neutts/neutts.py:46-70`` (itself after encodec's utility): frames
placed ``stride`` apart, each weighted by a triangular window ``0.5 - |t - 0.5|`` with
``t = linspace(0, 1, len + 2)[1:-1]``, summed and divided by the summed weights.
``chunk_plan`` restates the window bookkeeping of ``_infer_stream_ggml`` (``:401-421``, ``:443-465``).
"""
from __future__ import annotations

import numpy as np


def linear_overlap_add(frames, stride: int) -> np.ndarray:
    assert len(frames)
    dtype = frames[0].dtype
    total = max(stride * i + f.shape[-1] for i, f in enumerate(frames))
    sum_w = np.zeros(total, dtype=dtype)
    out = np.zeros(total, dtype=dtype)
    for i, f in enumerate(frames):
        n = f.shape[-1]
        t = np.linspace(0, 1, n + 2, dtype=dtype)[1:-1]
        w = np.abs(0.5 - (t - 0.5))
        out[i * stride: i * stride + n] += w * f
        sum_w[i * stride: i * stride + n] += w
    assert sum_w.min() > 0
    return out / sum_w


def chunk_plan(n_ref: int, n_total: int, hop=480, frames=25, lookforward=5, lookback=50, overlap=1):
    """(tokens_start, tokens_end, sample_start, sample_end) of every codec call the reference makes
    while streaming n_total - n_ref generated frames; the last tuple is the ragged tail (sample_end None)."""
    plan, n_dec = [], n_ref
    have = n_ref
    while True:
        # tokens arrive one at a time; a chunk fires once frames + lookforward undecoded tokens exist
        have = min(n_total, max(have, n_dec + frames + lookforward))
        if have - n_dec < frames + lookforward:
            break
        t0 = max(n_dec - lookback - overlap, 0)
        t1 = n_dec + frames + lookforward + overlap
        s0 = (n_dec - t0) * hop
        plan.append((t0, min(t1, have), s0, s0 + (frames + 2 * overlap) * hop))
        n_dec += frames
    if n_total > n_dec:
        rem = n_total - n_dec
        t0 = max(n_total - (lookback + overlap + rem), 0)
        s0 = (n_total - t0 - rem - overlap) * hop
        plan.append((t0, n_total, s0, None))
    return plan

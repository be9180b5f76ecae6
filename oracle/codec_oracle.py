"""CPU oracle for hot path B (NeuCodec token -> 24 kHz waveform decoder).

TEST INFRASTRUCTURE ONLY (same import rule as ``oracle/lm_oracle.py``).

PARITY UNPINNED.  The arithmetic the reference reaches through
``neutts/neutts.py:288-291`` (``self.codec.decode_code``) lives in the third-party
``neucodec`` package (``requirements.txt:2``: ``neucodec>=0.0.4``, no upper pin).
That package is not vendored in /root/reference, not installed in this image, not
in /opt/wheelhouse, and there is no network: neither its source nor a checkpoint
can be consulted.  This file restates the *published* architecture of the decoder
(NeuCodec = XCodec2-style: FSQ codebook -> Linear -> Vocos backbone with
transformer blocks -> ISTFT head) as described in SURVEY.md §3.4:

    decode_code(codes[B,1,N]) :
        FSQ indices -> 8 base-4 digits -> {-1,-.5,0,.5}      (vector_quantize_pytorch FSQ,
                                                              levels [4]*8, 4^8 = 65536 =
                                                              examples/finetune_config.yaml:7)
        quantizer.project_out  Linear(8 -> 2048)
        fc_post_a              Linear(2048 -> 1024)
        backbone.embed         Conv1d(1024,1024,k=7,p=3)
        backbone.prior_net     2 x ResnetBlock(GroupNorm32 -> swish -> Conv k3 -> GroupNorm32 -> swish -> Conv k3, +x)
        backbone.transformers  12 x [x + Attn(RMSNorm x) ; x + MLP(RMSNorm x)], 16 heads x 64,
                               fused bias-free QKV, RoPE, bidirectional SDPA, MLP 1024->4096->SiLU->1024
        backbone.post_net      2 x ResnetBlock
        backbone.final_layer_norm LayerNorm(1024, eps 1e-6)
        head.out               Linear(1024 -> n_fft+2 = 1922) -> (log-mag | phase)
        head.istft             mag = exp(.).clip(max=1e2); S = mag e^{i phase};
                               irfft(1920) * hann ; overlap-add hop 480 ; "same" trim 720 ; / envelope
        -> [B, 1, 480 N]                                      (hop 480: neutts/neutts.py:86)

Everything the reference itself pins about this path is honoured: 480 samples per
code (``neutts/neutts.py:84,86``), 65536 codes, int codes in / float PCM out, shapes
[B,1,N] -> [B,1,480N].  Layer shapes are config-driven so a real state_dict can
override them (``neutts_air_b200/loader.py``).

RoPE note: upstream is believed to call torchtune's RotaryPositionalEmbeddings on a
[b, h, t, d] tensor, which makes the rotation depend on the *head* index and
therefore cancel in q.k (a no-op on the output).  ``rope_axis`` selects "time"
(rotary over frames, interleaved pairs, base 10000) or "head" (the quirk).  Both
are restated here literally; the CUDA path implements "time" and skips "head".
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F


@dataclass
class CodecConfig:
    fsq_levels: int = 4
    fsq_dims: int = 8
    quant_dim: int = 2048
    hidden: int = 1024
    depth: int = 12
    heads: int = 16
    head_dim: int = 64
    mlp_mult: int = 4
    groups: int = 32
    embed_kernel: int = 7
    n_fft: int = 1920
    hop: int = 480
    rope_base: float = 10000.0
    rope_axis: str = "time"     # "time" | "head" (see module docstring)
    norm_eps: float = 1e-6
    mag_clip: float = 1e2

    @staticmethod
    def tiny(**kw) -> "CodecConfig":
        base = dict(quant_dim=64, hidden=128, depth=2, heads=2, head_dim=64, groups=32,
                    n_fft=64, hop=16)
        base.update(kw)
        return CodecConfig(**base)

    @property
    def codebook_size(self) -> int:
        return self.fsq_levels ** self.fsq_dims


@dataclass
class CodecWeights:
    project_out_w: torch.Tensor = None   # [quant_dim, 8]
    project_out_b: torch.Tensor = None
    fc_post_a_w: torch.Tensor = None     # [hidden, quant_dim]
    fc_post_a_b: torch.Tensor = None
    embed_w: torch.Tensor = None         # [hidden, hidden, 7]
    embed_b: torch.Tensor = None
    prior: list = field(default_factory=list)    # resnet dicts: n1w,n1b,c1w,c1b,n2w,n2b,c2w,c2b
    blocks: list = field(default_factory=list)   # dicts: att_norm, wqkv, wproj, ffn_norm, fc1, fc2
    post: list = field(default_factory=list)
    final_ln_w: torch.Tensor = None
    final_ln_b: torch.Tensor = None
    head_w: torch.Tensor = None          # [n_fft+2, hidden]
    head_b: torch.Tensor = None


def random_weights(cfg: CodecConfig, seed: int = 0) -> CodecWeights:
    """Seeded synthetic decoder.  Scales are chosen so activations stay O(1) through
    the stack and the PCM has speech-like level (RMS ~ 0.05-0.2, no clip at 1e2)."""
    g = torch.Generator().manual_seed(seed)
    C = cfg.hidden

    def rn(*shape, std):
        return torch.randn(*shape, generator=g) * std

    def lin(o, i, gain=1.0):
        return rn(o, i, std=gain / math.sqrt(i))

    def resnet():
        return dict(n1w=1.0 + rn(C, std=0.05), n1b=rn(C, std=0.05),
                    c1w=rn(C, C, 3, std=1.0 / math.sqrt(3 * C)), c1b=rn(C, std=0.02),
                    n2w=1.0 + rn(C, std=0.05), n2b=rn(C, std=0.05),
                    c2w=rn(C, C, 3, std=0.5 / math.sqrt(3 * C)), c2b=rn(C, std=0.02))

    w = CodecWeights()
    w.project_out_w = lin(cfg.quant_dim, cfg.fsq_dims, 1.5)
    w.project_out_b = rn(cfg.quant_dim, std=0.1)
    w.fc_post_a_w = lin(C, cfg.quant_dim)
    w.fc_post_a_b = rn(C, std=0.05)
    w.embed_w = rn(C, C, cfg.embed_kernel, std=1.0 / math.sqrt(cfg.embed_kernel * C))
    w.embed_b = rn(C, std=0.02)
    w.prior = [resnet() for _ in range(2)]
    for _ in range(cfg.depth):
        w.blocks.append(dict(att_norm=1.0 + rn(C, std=0.05), wqkv=lin(3 * C, C, 1.5), wproj=lin(C, C, 0.5),
                             ffn_norm=1.0 + rn(C, std=0.05), fc1=lin(cfg.mlp_mult * C, C),
                             fc2=lin(C, cfg.mlp_mult * C, 0.5)))
    w.post = [resnet() for _ in range(2)]
    w.final_ln_w = 1.0 + rn(C, std=0.05)
    w.final_ln_b = rn(C, std=0.05)
    nb = cfg.n_fft // 2 + 1
    hw = lin(2 * nb, C)
    hb = torch.zeros(2 * nb)
    hw[:nb] *= 0.5                      # log-magnitude spread
    # spectral tilt: log-mag bias falls with frequency so the PCM is speech-like in level
    hb[:nb] = 2.5 - 3.5 * torch.linspace(0, 1, nb)
    hw[nb:] *= 2.0                      # phases spread over several radians
    w.head_w, w.head_b = hw, hb
    return w


# ------------------------------------------------------------------------------------------
# stages
# ------------------------------------------------------------------------------------------

def fsq_dequant(codes: torch.Tensor, cfg: CodecConfig) -> torch.Tensor:
    """vector_quantize_pytorch FSQ.indices_to_codes: digit_i = (idx // L^i) % L, then
    (digit - L//2) / (L//2).  codes: int64 [..] -> float [.., fsq_dims]."""
    L = cfg.fsq_levels
    basis = L ** torch.arange(cfg.fsq_dims, dtype=torch.int64)
    digits = (codes.long()[..., None] // basis) % L
    half = L // 2
    return (digits.float() - half) / half


def group_norm_swish(x, w, b, groups, eps):
    h = F.group_norm(x, groups, w, b, eps)
    return h * torch.sigmoid(h)


def resnet_block(x: torch.Tensor, p: dict, cfg: CodecConfig) -> torch.Tensor:
    """x: [B, C, T].  GroupNorm -> swish -> conv3 -> GroupNorm -> swish -> (dropout off) -> conv3, + x."""
    h = group_norm_swish(x, p["n1w"], p["n1b"], cfg.groups, cfg.norm_eps)
    h = F.conv1d(h, p["c1w"], p["c1b"], padding=1)
    h = group_norm_swish(h, p["n2w"], p["n2b"], cfg.groups, cfg.norm_eps)
    h = F.conv1d(h, p["c2w"], p["c2b"], padding=1)
    return x + h


def rms_norm(x, w, eps):
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def rope_interleaved(x: torch.Tensor, positions: torch.Tensor, base: float) -> torch.Tensor:
    """torchtune RotaryPositionalEmbeddings arithmetic: pairs (x[2i], x[2i+1]) rotated by
    pos * base^(-2i/d).  x: [..., P, d] with positions [P] broadcast on dim -2."""
    d = x.shape[-1]
    theta = 1.0 / (base ** (torch.arange(0, d, 2).float() / d))
    ang = positions.float()[:, None] * theta[None, :]           # [P, d/2]
    cos, sin = ang.cos(), ang.sin()
    xs = x.float().reshape(*x.shape[:-1], d // 2, 2)
    out = torch.stack((xs[..., 0] * cos - xs[..., 1] * sin, xs[..., 1] * cos + xs[..., 0] * sin), dim=-1)
    return out.reshape(x.shape).to(x.dtype)


def transformer_block(x: torch.Tensor, p: dict, cfg: CodecConfig) -> torch.Tensor:
    """x: [B, T, C]."""
    B, T, C = x.shape
    h = rms_norm(x, p["att_norm"], cfg.norm_eps)
    qkv = h @ p["wqkv"].T                                        # 'b t (r h d)'
    qkv = qkv.view(B, T, 3, cfg.heads, cfg.head_dim).permute(2, 0, 3, 1, 4)  # r b h t d
    q, k, v = qkv[0], qkv[1], qkv[2]
    if cfg.rope_axis == "time":
        pos = torch.arange(T)
        q, k = rope_interleaved(q, pos, cfg.rope_base), rope_interleaved(k, pos, cfg.rope_base)
    elif cfg.rope_axis == "head":
        pos = torch.arange(cfg.heads)
        q = rope_interleaved(q.transpose(1, 2), pos, cfg.rope_base).transpose(1, 2)
        k = rope_interleaved(k.transpose(1, 2), pos, cfg.rope_base).transpose(1, 2)
    else:
        raise ValueError(cfg.rope_axis)
    s = (q @ k.transpose(-1, -2)) * (cfg.head_dim ** -0.5)
    a = torch.softmax(s, dim=-1) @ v                             # bidirectional
    a = a.transpose(1, 2).reshape(B, T, C)
    x = x + a @ p["wproj"].T
    h = rms_norm(x, p["ffn_norm"], cfg.norm_eps)
    x = x + F.silu(h @ p["fc1"].T) @ p["fc2"].T
    return x


def istft_same(spec: torch.Tensor, cfg: CodecConfig) -> torch.Tensor:
    """Vocos ISTFT(padding="same"): irfft(n_fft, norm=backward) * hann -> fold(hop) ->
    trim (win-hop)/2 each side -> / window envelope.  spec: complex [B, n_fft/2+1, T] -> [B, hop*T]."""
    n_fft, hop = cfg.n_fft, cfg.hop
    B, _, T = spec.shape
    window = torch.hann_window(n_fft)
    pad = (n_fft - hop) // 2
    frames = torch.fft.irfft(spec, n_fft, dim=1, norm="backward") * window[None, :, None]
    out_size = (T - 1) * hop + n_fft
    y = F.fold(frames, output_size=(1, out_size), kernel_size=(1, n_fft), stride=(1, hop))[:, 0, 0, pad:out_size - pad]
    wsq = window.square().expand(1, T, -1).transpose(1, 2)
    env = F.fold(wsq, output_size=(1, out_size), kernel_size=(1, n_fft), stride=(1, hop)).squeeze()[pad:out_size - pad]
    assert (env > 1e-11).all()
    return y / env


def head_spec(x: torch.Tensor, w: CodecWeights, cfg: CodecConfig) -> torch.Tensor:
    """ISTFTHead up to the complex spectrogram.  x: [B, T, C] -> complex [B, nb, T]."""
    o = (x @ w.head_w.T + w.head_b).transpose(1, 2)
    mag, ph = o.chunk(2, dim=1)
    mag = torch.exp(mag).clip(max=cfg.mag_clip)
    return mag * (torch.cos(ph) + 1j * torch.sin(ph))


def decode_code(codes: torch.Tensor, w: CodecWeights, cfg: CodecConfig, collect: dict | None = None) -> torch.Tensor:
    """codes: int [B, 1, N] -> float32 [B, 1, hop*N]  (the seam at neutts/neutts.py:288-291)."""
    assert codes.dim() == 3 and codes.shape[1] == 1
    z = fsq_dequant(codes[:, 0, :], cfg)                          # [B, N, 8]
    x = z @ w.project_out_w.T + w.project_out_b                   # [B, N, 2048]
    x = x @ w.fc_post_a_w.T + w.fc_post_a_b                       # [B, N, C]
    if collect is not None:
        collect["fc_post_a"] = x.clone()
    x = x.transpose(1, 2)
    x = F.conv1d(x, w.embed_w, w.embed_b, padding=cfg.embed_kernel // 2)
    if collect is not None:
        collect["embed"] = x.transpose(1, 2).clone()
    for p in w.prior:
        x = resnet_block(x, p, cfg)
    x = x.transpose(1, 2)
    if collect is not None:
        collect["prior"] = x.clone()
    for p in w.blocks:
        x = transformer_block(x, p, cfg)
    if collect is not None:
        collect["transformers"] = x.clone()
    x = x.transpose(1, 2)
    for p in w.post:
        x = resnet_block(x, p, cfg)
    x = x.transpose(1, 2)
    x = F.layer_norm(x, (cfg.hidden,), w.final_ln_w, w.final_ln_b, cfg.norm_eps)
    if collect is not None:
        collect["final"] = x.clone()
    spec = head_spec(x, w, cfg)
    return istft_same(spec, cfg)[:, None, :]

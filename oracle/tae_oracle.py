"""CPU oracle for the TAE hot path — TEST INFRASTRUCTURE ONLY.

This file is a plain restatement of the reference's algorithm (eminorhan/tae, `tae.py` and the optimizer-side
helpers of `util/misc.py`) as pure functions over a state_dict.  It is the *checker* for the CUDA path:
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import it.
Nothing under `tae_b200/` imports it, and the product path raises when the CUDA extension is missing.

Pinning: the reference publishes no golden vectors.  The oracle is pinned instead against
outputs of the UNMODIFIED reference module executed in the build container: `tests/golden/make_golden.py`
imports `/root/reference/tae.py`, runs it on seeded inputs (fp32 and bf16-autocast, forward + backward + one
AdamW step) and commits the results as small fixtures; `tests/test_oracle_golden.py` checks this restatement
against those fixtures and against the seed-pinned scalars of BASELINE.md §4 for the real patch16 model.

Third-party arithmetic: every FLOP of the reference runs inside PyTorch (unpinned by the reference; the build
image has torch 2.11.0).  The oracle therefore uses torch's primitive tensor ops (matmul, exp, erf, mean ...)
but restates every composite the reference calls (conv-as-im2col, LayerNorm, attention, GELU, MSE, AdamW).

Each function cites the reference file:line it follows (paths relative to the reference repo root).
`mode='fp32'` computes everything in fp32 (or fp64 if the inputs are fp64).  `mode='bf16'` reproduces the rounding
points of `torch.autocast(dtype=torch.bfloat16)`: Linear/conv/attention/GELU outputs are bf16, LayerNorm,
residual stream and loss are fp32.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict

import numpy as np
import torch


@dataclass(frozen=True)
class TAEConfig:
    img_size: int = 256
    patch_size: int = 16
    embed_dim: int = 1024
    vocab_size: int = 16
    depth: int = 15
    num_heads: int = 16
    decoder_embed_dim: int = 1024
    decoder_depth: int = 15
    decoder_num_heads: int = 16
    mlp_ratio: float = 4.0
    eps: float = 1e-6

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def num_patches(self) -> int:
        return self.grid * self.grid

    def as_kwargs(self) -> dict:
        d = asdict(self)
        d.pop("eps")
        return d


# the factory table of tae.py:434-483
ZOO = {16: (1024, 15, 16, (16, 64, 256)), 32: (2048, 18, 32, (64, 256, 1024)),
       64: (2560, 21, 32, (256, 1024, 4096)), 128: (2560, 22, 32, (1024, 4096, 16384))}


def zoo_config(name: str) -> TAEConfig:
    """'tae_patch16_vocab256_px256' -> TAEConfig  (tae.py:434-483)."""
    parts = name.split("_")
    patch, vocab = int(parts[1][len("patch"):]), int(parts[2][len("vocab"):])
    dim, depth, heads, vocabs = ZOO[patch]
    assert vocab in vocabs and parts[3] == "px256", name
    return TAEConfig(256, patch, dim, vocab, depth, heads, dim, depth, heads)


def state_dict_spec(cfg: TAEConfig):
    """Ordered (key, shape) list of the checkpoint `model` dict (tae.py:155-169)."""
    D, Dd, p, V, N = cfg.embed_dim, cfg.decoder_embed_dim, cfg.patch_size, cfg.vocab_size, cfg.num_patches
    spec = [("pos_embed", (1, N, D)), ("decoder_pos_embed", (1, N, Dd)),
            ("patch_embed.proj.weight", (D, 3, p, p)), ("patch_embed.proj.bias", (D,))]

    def block(prefix, d):
        hdim = int(d * cfg.mlp_ratio)
        return [(f"{prefix}.norm1.weight", (d,)), (f"{prefix}.norm1.bias", (d,)),
                (f"{prefix}.attn.qkv.weight", (3 * d, d)), (f"{prefix}.attn.qkv.bias", (3 * d,)),
                (f"{prefix}.attn.proj.weight", (d, d)), (f"{prefix}.attn.proj.bias", (d,)),
                (f"{prefix}.norm2.weight", (d,)), (f"{prefix}.norm2.bias", (d,)),
                (f"{prefix}.mlp.fc1.weight", (hdim, d)), (f"{prefix}.mlp.fc1.bias", (hdim,)),
                (f"{prefix}.mlp.fc2.weight", (d, hdim)), (f"{prefix}.mlp.fc2.bias", (d,))]

    for i in range(cfg.depth):
        spec += block(f"blocks.{i}", D)
    spec += [("norm.weight", (D,)), ("norm.bias", (D,)), ("dict_proj.weight", (V, D)),
             ("decoder_embed.weight", (Dd, V)), ("decoder_embed.bias", (Dd,))]
    for i in range(cfg.decoder_depth):
        spec += block(f"decoder_blocks.{i}", Dd)
    spec += [("decoder_norm.weight", (Dd,)), ("decoder_norm.bias", (Dd,)),
             ("decoder_pred.weight", (3 * p * p, Dd)), ("decoder_pred.bias", (3 * p * p,))]
    return spec


# ----------------------------------------------------------------------------------------------------
# Integer index maps (numpy, bit-exact gates)
# ----------------------------------------------------------------------------------------------------
def patchify_np(imgs: np.ndarray, p: int) -> np.ndarray:
    """tae.py:196-208: out[n, h*g+w, (i*p+j)*3+c] = imgs[n, c, h*p+i, w*p+j]."""
    n, c, H, W = imgs.shape
    assert c == 3 and H == W and H % p == 0
    g = H // p
    x = imgs.reshape(n, 3, g, p, g, p)            # n c h i w j
    x = x.transpose(0, 2, 4, 3, 5, 1)             # n h w i j c
    return np.ascontiguousarray(x).reshape(n, g * g, p * p * 3)


def unpatchify_np(x: np.ndarray, p: int) -> np.ndarray:
    """tae.py:210-222: exact inverse of patchify_np."""
    n, L, K = x.shape
    g = int(L ** .5)
    assert g * g == L and K == 3 * p * p
    y = x.reshape(n, g, g, p, p, 3)               # n h w i j c
    y = y.transpose(0, 5, 1, 3, 2, 4)             # n c h i w j
    return np.ascontiguousarray(y).reshape(n, 3, g * p, g * p)


def im2col_np(imgs: np.ndarray, p: int) -> np.ndarray:
    """The gather a stride-p, kernel-p conv performs (tae.py:43,50-52):
    cols[n*N + h*g+w, c*p*p + i*p + j] = imgs[n, c, h*p+i, w*p+j]  — K-order (c,i,j), unlike patchify's (i,j,c)."""
    n, c, H, W = imgs.shape
    g = H // p
    x = imgs.reshape(n, 3, g, p, g, p)            # n c h i w j
    x = x.transpose(0, 2, 4, 1, 3, 5)             # n h w c i j
    return np.ascontiguousarray(x).reshape(n * g * g, 3 * p * p)


def patchify(imgs: torch.Tensor, p: int) -> torch.Tensor:
    n, _, H, _ = imgs.shape
    g = H // p
    return imgs.reshape(n, 3, g, p, g, p).permute(0, 2, 4, 3, 5, 1).reshape(n, g * g, p * p * 3)


def unpatchify(x: torch.Tensor, p: int) -> torch.Tensor:
    n, L, _ = x.shape
    g = int(L ** .5)
    return x.reshape(n, g, g, p, p, 3).permute(0, 5, 1, 3, 2, 4).reshape(n, 3, g * p, g * p)


def im2col(imgs: torch.Tensor, p: int) -> torch.Tensor:
    n, _, H, _ = imgs.shape
    g = H // p
    return imgs.reshape(n, 3, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(n, g * g, 3 * p * p)


# ----------------------------------------------------------------------------------------------------
# Building blocks
# ----------------------------------------------------------------------------------------------------
def _lp(t: torch.Tensor, mode: str) -> torch.Tensor:
    """autocast rounding point: cast to bf16 in 'bf16' mode, identity otherwise."""
    return t.to(torch.bfloat16) if mode == "bf16" else t


def linear(x, w, b, mode):
    """nn.Linear under autocast: operands (and bias) cast to bf16, fp32 accumulation, ONE rounding of the output
    (the tensor-core GEMM semantics of the reference's cuBLASLt call); plain matmul in fp32 mode."""
    if mode == "bf16":
        y = x.to(torch.bfloat16).float() @ w.to(torch.bfloat16).float().t()
        if b is not None:
            y = y + b.to(torch.bfloat16).float()
        return y.to(torch.bfloat16)
    y = x @ w.t()
    if b is not None:
        y = y + b
    return y


def layer_norm(x, w, b, eps):
    """nn.LayerNorm (tae.py:122,125,159,168): fp32, biased variance; autocast keeps it in fp32."""
    x = x.float() if x.dtype in (torch.bfloat16, torch.float16) else x
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu(x):
    """nn.GELU(), approximate='none' (tae.py:96,102): 0.5 x (1 + erf(x / sqrt 2)); computed in fp32, output in x.dtype."""
    xf = x.float() if x.dtype in (torch.bfloat16, torch.float16) else x
    y = 0.5 * xf * (1.0 + torch.erf(xf * (1.0 / math.sqrt(2.0))))
    return y.to(x.dtype)


def attention(x, w_qkv, b_qkv, w_proj, b_proj, num_heads, mode):
    """Attention.forward (tae.py:72-82): qkv rows [0:D]=q, [D:2D]=k, [2D:3D]=v, each head-major; softmax(q k^T / sqrt(hd)) v."""
    B, N, C = x.shape
    hd = C // num_heads
    qkv = linear(x, w_qkv, b_qkv, mode).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    # F.scaled_dot_product_attention (flash semantics): fp32 scores and softmax, probabilities rounded to the input
    # dtype for the P@V product, fp32 accumulation, output in the input dtype
    if mode == "bf16":
        s = q.float() @ k.float().transpose(-2, -1) * (hd ** -0.5)
        pr = torch.softmax(s, dim=-1)
        o = (pr.to(torch.bfloat16).float() @ v.float()).to(torch.bfloat16)
    else:
        s = q @ k.transpose(-2, -1) * (hd ** -0.5)
        o = torch.softmax(s, dim=-1) @ v
    o = o.transpose(1, 2).reshape(B, N, C)
    return linear(o, w_proj, b_proj, mode)


def block(x, sd, prefix, num_heads, eps, mode, acts=None):
    """Block.forward (tae.py:128-131): pre-LN residual; fp32 stream + low-precision branch -> fp32."""
    g = lambda k: sd[f"{prefix}.{k}"]
    a = attention(layer_norm(x, g("norm1.weight"), g("norm1.bias"), eps), g("attn.qkv.weight"), g("attn.qkv.bias"),
                  g("attn.proj.weight"), g("attn.proj.bias"), num_heads, mode)
    x = x + a
    h = linear(layer_norm(x, g("norm2.weight"), g("norm2.bias"), eps), g("mlp.fc1.weight"), g("mlp.fc1.bias"), mode)
    m = linear(gelu(h), g("mlp.fc2.weight"), g("mlp.fc2.bias"), mode)
    x = x + m
    if acts is not None:
        acts[prefix] = x
    return x


def patch_embed(imgs, w, b, p, mode):
    """PatchEmbed.forward (tae.py:46-54): Conv2d(3, D, k=p, s=p) == im2col(c,i,j) @ W.view(D,-1)^T + b."""
    cols = im2col(imgs, p)
    return linear(cols, w.reshape(w.shape[0], -1), b, mode)


def forward_encoder(sd, imgs, cfg: TAEConfig, mode="fp32", acts=None):
    """TAE.forward_encoder (tae.py:224-238)."""
    x = patch_embed(imgs, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], cfg.patch_size, mode)
    x = x + sd["pos_embed"]
    if acts is not None:
        acts["embed"] = x
    for i in range(cfg.depth):
        x = block(x, sd, f"blocks.{i}", cfg.num_heads, cfg.eps, mode, acts)
    x = layer_norm(x, sd["norm.weight"], sd["norm.bias"], cfg.eps)
    return linear(x, sd["dict_proj.weight"], None, mode)


def forward_decoder(sd, z, cfg: TAEConfig, mode="fp32", acts=None):
    """TAE.forward_decoder (tae.py:240-254)."""
    x = linear(z, sd["decoder_embed.weight"], sd["decoder_embed.bias"], mode)
    x = x + sd["decoder_pos_embed"]
    if acts is not None:
        acts["decoder_embed"] = x
    for i in range(cfg.decoder_depth):
        x = block(x, sd, f"decoder_blocks.{i}", cfg.decoder_num_heads, cfg.eps, mode, acts)
    x = layer_norm(x, sd["decoder_norm.weight"], sd["decoder_norm.bias"], cfg.eps)
    return linear(x, sd["decoder_pred.weight"], sd["decoder_pred.bias"], mode)


def forward_loss(imgs, pred, p):
    """TAE.forward_loss (tae.py:256-265): mean over all elements of (pred - patchify(imgs))^2, in fp32."""
    target = patchify(imgs, p)
    predf = pred.float() if pred.dtype in (torch.bfloat16, torch.float16) else pred
    return ((predf - target) ** 2).mean()


def forward(sd, imgs, cfg: TAEConfig, mode="fp32", acts=None):
    """TAE.forward (tae.py:267-271) -> (loss, pred, latent)."""
    latent = forward_encoder(sd, imgs, cfg, mode, acts)
    pred = forward_decoder(sd, latent, cfg, mode, acts)
    loss = forward_loss(imgs, pred, cfg.patch_size)
    return loss, pred, latent


def forward_backward(sd, imgs, cfg: TAEConfig, mode="fp32"):
    """Forward + autograd backward over the restated graph; returns (loss, pred, latent, grads dict)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    loss, pred, latent = forward(leaves, imgs, cfg, mode)
    loss.backward()
    return loss.detach(), pred.detach(), latent.detach(), {k: v.grad for k, v in leaves.items()}


# ----------------------------------------------------------------------------------------------------
# Optimizer-side helpers
# ----------------------------------------------------------------------------------------------------
def add_weight_decay_names(named_shapes, weight_decay, skip_list=()):
    """util/misc.py:364-379: 1-D params and '*.bias' -> no_decay group; everything else -> decay group.
    named_shapes: iterable of (name, shape).  Returns (no_decay_names, decay_names)."""
    no_decay, decay = [], []
    for name, shape in named_shapes:
        if len(shape) == 1 or name.endswith(".bias") or name in skip_list:
            no_decay.append(name)
        else:
            decay.append(name)
    return no_decay, decay


def adjust_learning_rate(max_lr, min_lr, it, switch_it):
    """util/misc.py:400-412: step schedule."""
    return max_lr if it < switch_it else min_lr


def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.0):
    """torch.optim.AdamW (train.py:109; decoupled weight decay, bias-corrected), single-tensor restatement.
    Returns new (p, m, v); `step` is the 1-based step count."""
    p = p * (1.0 - lr * weight_decay)
    m = m + (g - m) * (1.0 - beta1)
    v = beta2 * v + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def grad_norm(grads) -> torch.Tensor:
    """util/misc.py:274-286 (norm_type 2): norm of the per-tensor norms."""
    return torch.norm(torch.stack([torch.norm(g.detach().float(), 2.0) for g in grads if g is not None]), 2.0)


def init_state_dict(cfg: TAEConfig, seed: int = 0, device="cpu"):
    """A cheap deterministic init for property tests at full size (NOT the reference's init distribution order —
    use tae_b200.tae factories under torch.manual_seed for init parity).  Xavier-uniform weights, zero biases,
    small-normal pos-embeds."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    for k, shape in state_dict_spec(cfg):
        if k.endswith("pos_embed"):
            t = torch.randn(shape, generator=g) * 0.02
        elif "norm" in k and k.endswith("weight"):
            t = torch.ones(shape)
        elif k.endswith("bias"):
            t = torch.zeros(shape) if "patch_embed" not in k else (torch.rand(shape, generator=g) - 0.5) * 0.05
        else:
            fan_out = shape[0]
            fan_in = int(np.prod(shape[1:]))
            bound = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[k] = t.to(device)
    return sd

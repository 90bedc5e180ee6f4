"""Per-kernel timing probe on a B200 (CUDA events, L2 flushed between timed launches by rotating large buffers).

    python tools/gpu_probe.py [--quick]

Prints one line per kernel/shape: time, achieved TFLOP/s or GB/s, fraction of the measured peak.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tae_b200 import ops  # noqa: E402
from tae_b200._lib import (EPI_BF16, EPI_BF16_DGELU, EPI_BF16_GELU, EPI_F32_ACC, EPI_F32_RESID)  # noqa: E402


def peaks():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in ev:
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--attn-only", action="store_true")
    ap.add_argument("--gemm-only", action="store_true")
    ap.add_argument("--small-grids", action="store_true", help="attention and LayerNorm at the patch32/64/128 shapes only")
    ap.add_argument("--only", default=None, help="with --gemm-only: just the layers whose name starts with this (qkv, proj, fc1, fc2)")
    args = ap.parse_args()
    hbm, tf, how = peaks()
    dev = "cuda"
    bf = torch.bfloat16
    M = 16384 if args.quick else 65536
    print(f"# peaks ({how}): HBM {hbm} GB/s, bf16 {tf} TFLOP/s; M={M}")

    def rnd(*s, dtype=bf):
        return (torch.randn(*s, device=dev) * 0.5).to(dtype)

    if args.small_grids:
        # attention on the short grids (B=256): bytes = qkv read (+ dout) and out / dqkv written
        for N, H, hd in ((64, 32, 64), (16, 32, 80), (4, 32, 80)):
            B, D = 256, H * hd
            qkv, dout = rnd(B * N, 3 * D), rnd(B * N, D)
            med, _ = timeit(lambda: ops.attention_fwd(qkv, B, N, H, hd))
            by = B * N * D * 2.0 * 4
            print(f"attention fwd N={N} H={H} hd={hd}: {med * 1e3:.1f} us  {by / med / 1e6:.0f} GB/s ({by / med / 1e6 / hbm:.2f} of HBM peak)")
            out, lse = ops.attention_fwd(qkv, B, N, H, hd)
            med, _ = timeit(lambda: ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd))
            by = B * N * D * 2.0 * 8
            print(f"attention bwd N={N} (O staged): {med * 1e3:.1f} us  {by / med / 1e6:.0f} GB/s ({by / med / 1e6 / hbm:.2f})")
            if ops.attention_takes_delta(N, hd):
                delta = (dout.float() * out.float()).view(B, N, H, hd).sum(-1).permute(0, 2, 1).contiguous()
                med, _ = timeit(lambda: ops.attention_bwd(qkv, None, dout, lse, B, N, H, hd, delta=delta))
                by = B * N * D * 2.0 * 7
                print(f"attention bwd N={N} (delta given): {med * 1e3:.1f} us  {by / med / 1e6:.0f} GB/s ({by / med / 1e6 / hbm:.2f})")
        # LayerNorm at the residual-stream shapes of patch32 / patch64 / patch128 (+ patch16 for reference)
        for rows, D in ((65536, 1024), (16384, 2048), (4096, 2560), (1024, 2560)):
            xf = torch.randn(rows, D, device=dev)
            w, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
            med, _ = timeit(lambda: ops.layernorm_fwd(xf, w, b, 1e-6))
            by = rows * D * 6.0
            print(f"layernorm fwd rows={rows} D={D}: {med * 1e3:.1f} us  {by / med / 1e6:.0f} GB/s ({by / med / 1e6 / hbm:.2f})")
            y, mean, rstd = ops.layernorm_fwd(xf, w, b, 1e-6)
            dy, dres = rnd(rows, D), torch.randn(rows, D, device=dev)
            med, _ = timeit(lambda: ops.layernorm_bwd(dy, xf, mean, rstd, w, dres))
            by = rows * D * 16.0
            print(f"layernorm bwd rows={rows} D={D} (+finalize): {med * 1e3:.1f} us  {by / med / 1e6:.0f} GB/s ({by / med / 1e6 / hbm:.2f})")
        return

    # ---- GEMMs of one patch16 block (D=1024) ----
    D = 1024
    x = rnd(M, D)
    shapes = [("qkv  fwd", D, 3 * D), ("proj fwd", D, D), ("fc1  fwd", D, 4 * D), ("fc2  fwd", 4 * D, D)]
    if args.only:
        shapes = [t for t in shapes if t[0].startswith(args.only)]
    for name, K, N in ([] if args.attn_only else shapes):
        A, W = rnd(M, K), rnd(N, K)
        bias = torch.randn(N, device=dev)
        fl = 2.0 * M * N * K
        if name.startswith("fc1"):
            f = lambda: ops.gemm(A, W, epilogue=EPI_BF16_GELU, bias=bias)
        elif name.startswith(("proj", "fc2")):
            res = torch.randn(M, N, device=dev)
            f = lambda: ops.gemm(A, W, epilogue=EPI_F32_RESID, bias=bias, resid=res)
        else:
            f = lambda: ops.gemm(A, W, epilogue=EPI_BF16, bias=bias)
        med, best = timeit(f)
        print(f"gemm {name} M={M} N={N} K={K}: {med:.3f} ms  {fl / med / 1e9:.0f} TFLOP/s ({fl / med / 1e9 / tf:.2f} of peak)  best {fl / best / 1e9:.0f}")
        tm, tb = timeit(lambda: torch.matmul(A, W.t()))
        print(f"     cuBLAS (torch.matmul, no epilogue): {tm:.3f} ms  {fl / tm / 1e9:.0f} TFLOP/s")
        # dgrad: dX[M,K] = dY[M,N] W[N,K]
        dY = rnd(M, N)
        if name.startswith("fc2"):
            haux = rnd(M, K)
            f = lambda: ops.gemm(dY, W, b_mn=True, epilogue=EPI_BF16_DGELU, aux=haux)
        else:
            f = lambda: ops.gemm(dY, W, b_mn=True, epilogue=EPI_BF16)
        med, best = timeit(f)
        print(f"gemm {name[:4]} dgrad: {med:.3f} ms  {fl / med / 1e9:.0f} TFLOP/s ({fl / med / 1e9 / tf:.2f})")
        # wgrad: dW[N,K] = dY^T A
        dW = torch.empty(N, K, device=dev)
        f = lambda: ops.gemm(dY, A, a_mn=True, b_mn=True, epilogue=EPI_F32_ACC, out=dW)
        med, best = timeit(f)
        print(f"gemm {name[:4]} wgrad: {med:.3f} ms  {fl / med / 1e9:.0f} TFLOP/s ({fl / med / 1e9 / tf:.2f})")
        del A, W, dY, dW

    if args.gemm_only:
        return
    # ---- attention (patch16: N=256, hd=64, H=16) ----
    B = M // 256
    qkv = rnd(M, 3 * D)
    med, _ = timeit(lambda: ops.attention_fwd(qkv, B, 256, 16, 64))
    fl = 4.0 * B * 16 * 256 * 256 * 64
    print(f"attention fwd B={B} N=256 H=16: {med:.3f} ms  {fl / med / 1e9:.0f} TFLOP/s")
    out, lse = ops.attention_fwd(qkv, B, 256, 16, 64)
    dout = rnd(M, D)
    med, _ = timeit(lambda: ops.attention_bwd(qkv, out, dout, lse, B, 256, 16, 64))
    print(f"attention bwd: {med:.3f} ms  {2.0 * fl / med / 1e9:.0f} TFLOP/s (algorithmic 2x fwd; 2.5x executed)")
    delta = (dout.float() * out.float()).view(B, 256, 16, 64).sum(-1).permute(0, 2, 1).contiguous()
    med, _ = timeit(lambda: ops.attention_bwd(qkv, None, dout, lse, B, 256, 16, 64, delta=delta))
    print(f"attention bwd (delta given): {med:.3f} ms  {2.0 * fl / med / 1e9:.0f} TFLOP/s")

    if args.attn_only:
        return
    # ---- LayerNorm ----
    xf = torch.randn(M, D, device=dev)
    w, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    med, _ = timeit(lambda: ops.layernorm_fwd(xf, w, b, 1e-6))
    by = M * D * 6.0
    print(f"layernorm fwd: {med:.3f} ms  {by / med / 1e6:.0f} GB/s ({by / med / 1e6 / hbm:.2f} of peak)")
    y, mean, rstd = ops.layernorm_fwd(xf, w, b, 1e-6)
    dy = rnd(M, D)
    dres = torch.randn(M, D, device=dev)
    med, _ = timeit(lambda: ops.layernorm_bwd(dy, xf, mean, rstd, w, dres))
    by = M * D * 16.0
    print(f"layernorm bwd (+resid add, +bf16 cast): {med:.3f} ms  {by / med / 1e6:.0f} GB/s ({by / med / 1e6 / hbm:.2f})")

    # ---- colsum, loss, adamw ----
    big = rnd(M, 4 * D)
    med, _ = timeit(lambda: ops.colsum(big))
    by = M * 4 * D * 2.0
    print(f"colsum [M,4D]: {med:.3f} ms  {by / med / 1e6:.0f} GB/s ({by / med / 1e6 / hbm:.2f})")
    Bi = M // 256
    imgs = torch.randn(Bi, 3, 256, 256, device=dev)
    pred = rnd(Bi, 256, 768)
    med, _ = timeit(lambda: ops.mse_loss(pred, imgs, 16, want_grad=True))
    by = Bi * 196608 * 8.0
    print(f"mse loss+grad: {med:.3f} ms  {by / med / 1e6:.0f} GB/s ({by / med / 1e6 / hbm:.2f})")
    med, _ = timeit(lambda: ops.im2col(imgs, 16))
    by = Bi * 196608 * 6.0
    print(f"im2col: {med:.3f} ms  {by / med / 1e6:.0f} GB/s ({by / med / 1e6 / hbm:.2f})")
    n = 380_000_000 if not args.quick else 50_000_000
    p, g, m, v = (torch.randn(n, device=dev) * 0.01 for _ in range(4))
    v = v.abs()
    pb = torch.empty(n, dtype=bf, device=dev)
    med, _ = timeit(lambda: ops.adamw_step(p, g, m, v, pb, lr=1e-4, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.0, step=3), iters=5)
    by = n * 30.0
    print(f"adamw n={n}: {med:.3f} ms  {by / med / 1e6:.0f} GB/s ({by / med / 1e6 / hbm:.2f})")


if __name__ == "__main__":
    main()

"""Developer check (GPU): vmae_gemm_bf16 vs torch.matmul for every operand-major combination and epilogue.
Run:  python tools/dev_check_gemm.py [kk|kmn|mnmn|perf]   (each mode in its own process: a trap kills the context)."""
import ctypes as C
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ssl_vit_video_analytics_b200 import _lib as L, ops  # noqa: E402

lib = C.CDLL(str(L.LIB_PATH))
lib.vmae_gemm_bf16.argtypes = L.SIGNATURES["vmae_gemm_bf16"]
lib.vmae_last_error.restype = C.c_char_p
dev = torch.device("cuda:0")


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def gemm(A, lda, a_major, B, ldb, b_major, M, N, K, bias=None, residual=None, aux_in=None, aux_out=None, out=None,
         out_fp32=False, accumulate=False, act=0, k_splits=1):
    rc = lib.vmae_gemm_bf16(ptr(A), lda, a_major, ptr(B), ldb, b_major, M, N, K, ptr(bias), ptr(residual),
                            residual.stride(0) if residual is not None else 0, ptr(aux_in),
                            aux_in.stride(0) if aux_in is not None else 0, ptr(aux_out),
                            aux_out.stride(0) if aux_out is not None else 0, ptr(out), out.stride(0),
                            1 if out_fp32 else 0, 1 if accumulate else 0, act, k_splits,
                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError(f"rc={rc}: {lib.vmae_last_error().decode()}")


def report(name, got, ref):
    torch.cuda.synchronize()
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item() + 1e-9
    ok = err / scale < 2e-2
    print(f"{'OK ' if ok else 'BAD'} {name}: max_abs_err={err:.4e} ref_max={scale:.3e} rel={err/scale:.3e}", flush=True)
    return ok


def mode_kk():
    ok = True
    g = torch.Generator(device="cuda").manual_seed(0)
    for (M, N, K) in [(128, 64, 64), (256, 384, 384), (1000, 1152, 384), (512, 1536, 384), (384, 384, 1536), (300, 192, 384), (128, 96, 200), (2048, 256, 128)]:
        x = torch.randn(M, K, device=dev, generator=g).bfloat16()
        w = (torch.randn(N, K, device=dev, generator=g) * 0.05).bfloat16()
        b = torch.randn(N, device=dev, generator=g)
        ref = x.float() @ w.float().t()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        gemm(x, K, 0, w, K, 0, M, N, K, out=out)
        ok &= report(f"kk plain {M}x{N}x{K}", out, ref)
        gemm(x, K, 0, w, K, 0, M, N, K, bias=b, out=out)
        ok &= report(f"kk bias  {M}x{N}x{K}", out, ref + b)
        pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        gemm(x, K, 0, w, K, 0, M, N, K, bias=b, aux_out=pre, out=out, act=1)
        ok &= report(f"kk gelu  {M}x{N}x{K}", out, torch.nn.functional.gelu(ref + b))
        pr = (ref + b).requires_grad_(True)
        torch.nn.functional.gelu(pr).sum().backward()
        ok &= report(f"kk gelu' saved {M}x{N}x{K}", pre, pr.grad)
        res = torch.randn(M, N, device=dev, generator=g)
        o32 = torch.empty(M, N, device=dev, dtype=torch.float32)
        gemm(x, K, 0, w, K, 0, M, N, K, bias=b, residual=res, out=o32, out_fp32=True)
        ok &= report(f"kk resid {M}x{N}x{K}", o32, ref + b + res)
        h = torch.randn(M, N, device=dev, generator=g).bfloat16()
        gemm(x, K, 0, w, K, 0, M, N, K, aux_in=h, out=out)
        ok &= report(f"kk dgelu {M}x{N}x{K}", out, ref * h.float())
    return ok


def mode_kmn():
    ok = True
    g = torch.Generator(device="cuda").manual_seed(1)
    for (M, Np, Kp) in [(128, 64, 64), (256, 384, 384), (1000, 1152, 384), (512, 1536, 384), (384, 384, 1536), (300, 192, 384)]:
        dy = torch.randn(M, Np, device=dev, generator=g).bfloat16()
        w = (torch.randn(Np, Kp, device=dev, generator=g) * 0.05).bfloat16()
        ref = dy.float() @ w.float()  # [M, Kp]
        out = torch.empty(M, Kp, device=dev, dtype=torch.bfloat16)
        gemm(dy, Np, 0, w, Kp, 1, M, Kp, Np, out=out)
        ok &= report(f"kmn dgrad M={M} N'={Np} K'={Kp}", out, ref)
    return ok


def mode_mnmn():
    ok = True
    g = torch.Generator(device="cuda").manual_seed(2)
    for (T, Np, Kp, ks) in [(64, 128, 64, 1), (512, 384, 384, 1), (4096, 1152, 384, 4), (10240, 1536, 384, 8), (3000, 384, 1536, 3), (1000, 192, 384, 2)]:
        dy = (torch.randn(T, Np, device=dev, generator=g) * 0.1).bfloat16()
        x = torch.randn(T, Kp, device=dev, generator=g).bfloat16()
        ref = dy.float().t() @ x.float()  # [Np, Kp]
        out = torch.zeros(Np, Kp, device=dev, dtype=torch.float32)
        gemm(dy, Np, 1, x, Kp, 1, Np, Kp, T, out=out, out_fp32=True, accumulate=True, k_splits=ks)
        ok &= report(f"mnmn wgrad T={T} N'={Np} K'={Kp} ks={ks}", out, ref)
    return ok


def _time(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def mode_perf():
    g = torch.Generator(device="cuda").manual_seed(3)
    peak = 1644.1
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 100352   # 100352 = decoder rows (64 x 1568); 10240 = encoder rows (64 x 160)
    for (N, K, tag) in [(1536, 384, "fc1"), (384, 1536, "fc2"), (1152, 384, "qkv"), (384, 384, "proj")]:
        x = torch.randn(M, K, device=dev, generator=g).bfloat16()
        w = (torch.randn(N, K, device=dev, generator=g) * 0.05).bfloat16()
        b = torch.randn(N, device=dev, generator=g)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        o32 = torch.empty(M, N, device=dev, dtype=torch.float32)
        res = torch.randn(M, N, device=dev, generator=g)
        pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        fl = 2.0 * M * N * K
        rows = []
        rows.append(("plain", _time(lambda: gemm(x, K, 0, w, K, 0, M, N, K, out=out))))
        rows.append(("bias", _time(lambda: gemm(x, K, 0, w, K, 0, M, N, K, bias=b, out=out))))
        rows.append(("bias+gelu+pre", _time(lambda: gemm(x, K, 0, w, K, 0, M, N, K, bias=b, aux_out=pre, out=out, act=1))))
        rows.append(("bias+resid(f32)", _time(lambda: gemm(x, K, 0, w, K, 0, M, N, K, bias=b, residual=res, out=o32, out_fp32=True))))
        rows.append(("dgelu(aux_in)", _time(lambda: gemm(x, K, 0, w, K, 0, M, N, K, aux_in=pre, out=out))))
        rows.append(("cuBLAS", _time(lambda: torch.matmul(x, w.t(), out=out))))
        # dgrad: dy [M,N] @ W [N,K] -> [M,K]
        dy = torch.randn(M, N, device=dev, generator=g).bfloat16()
        dx = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
        rows.append(("dgrad", _time(lambda: gemm(dy, N, 0, w, K, 1, M, K, N, out=dx))))
        dw = torch.zeros(N, K, device=dev, dtype=torch.float32)
        rows.append(("wgrad auto", _time(lambda: ops.linear_wgrad(dy, x))))
        for ks in ((4, 8, 12, 16, 24, 32, 48) if len(sys.argv) > 3 else (4, 12, 24)):
            rows.append((f"wgrad ks={ks}", _time(lambda: gemm(dy, N, 1, x, K, 1, N, K, M, out=dw, out_fp32=True, accumulate=True, k_splits=ks))))
        print(f"perf {tag} M={M} N={N} K={K}: " + " | ".join(f"{n} {ms*1e3:.0f}us {fl/ms/1e9:.0f}TF" for n, ms in rows), flush=True)
    return True


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "kk"
    t0 = time.time()
    ok = {"kk": mode_kk, "kmn": mode_kmn, "mnmn": mode_mnmn, "perf": mode_perf,
          "all": lambda: mode_kk() & mode_kmn() & mode_mnmn()}[mode]()
    print(f"mode {mode}: {'ALL OK' if ok else 'FAILURES'} ({time.time()-t0:.1f}s)")
    sys.exit(0 if ok else 1)

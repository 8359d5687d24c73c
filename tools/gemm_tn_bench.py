"""TFLOP/s of the learner-body GEMM (csrc/gemm_tn.cu) on the Qwen2.5-7B shapes of a 16 384-token micro-batch,
next to cuBLAS (torch.mm) on the same operands."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pipelinerl_b200 import _lib  # noqa: E402


def time_ms(fn, iters=10):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    lib = _lib.load()
    dev = torch.device("cuda:0")
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    shapes = {"qkv fwd": (T, 4608, 3584), "o fwd": (T, 3584, 3584), "gate_up fwd": (T, 37888, 3584),
              "down fwd": (T, 3584, 18944), "gate_up dgrad": (T, 3584, 37888), "gate_up wgrad": (37888, 3584, T),
              "down wgrad": (3584, 18944, T)}
    out = {}
    for name, (M, N, K) in shapes.items():
        A = (torch.randn(M, K, device=dev) * 0.1).to(torch.bfloat16)
        B = (torch.randn(N, K, device=dev) * 0.1).to(torch.bfloat16)
        f32 = "wgrad" in name
        C = torch.zeros(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device=dev)
        st = _lib.stream_ptr()

        def ours():
            _lib.check(lib.prl_gemm_tn(A.data_ptr(), K, B.data_ptr(), K, M, N, K, C.data_ptr(), N, int(f32), int(f32),
                                       None, None, 0, 1.0, st))
        Bt = B.t()

        def cublas():
            torch.mm(A, Bt)
        t_o, t_c = time_ms(ours), time_ms(cublas)
        fl = 2.0 * M * N * K
        out[name] = {"M": M, "N": N, "K": K, "ours_ms": round(t_o, 4), "ours_TFLOPs": round(fl / t_o / 1e9, 1),
                     "cublas_ms": round(t_c, 4), "cublas_TFLOPs": round(fl / t_c / 1e9, 1)}
        if "fwd" in name:  # the prefill path: fp32 [1, M, N] output through either kernel
            P = torch.zeros(1, M, N, dtype=torch.float32, device=dev)
            t_sw = time_ms(lambda: _lib.check(lib.prl_gemm_bf16_splitk(B.data_ptr(), None, A.data_ptr(), M, N, K, 1,
                                                                       P.data_ptr(), st)))
            t_f32 = time_ms(lambda: _lib.check(lib.prl_gemm_tn(A.data_ptr(), K, B.data_ptr(), K, M, N, K, P.data_ptr(), N,
                                                               1, 0, None, None, 0, 1.0, st)))
            out[name].update({"swapab_pair_f32_ms": round(t_sw, 4), "tn_f32_ms": round(t_f32, 4)})
            del P
        print(name, out[name], flush=True)
        del A, B, C
    # operands as stored (MN-major): dgrad reads W [out, in] as B[K=out, N=in]; wgrad reads dY [T, out], X [T, in]
    mn = {"down dgrad (B mn)": (T, 18944, 3584, False, True), "gate_up dgrad (B mn)": (T, 3584, 37888, False, True),
          "gate_up wgrad (A,B mn)": (37888, 3584, T, True, True), "down wgrad (A,B mn)": (3584, 18944, T, True, True),
          "qkv wgrad (A,B mn)": (4608, 3584, T, True, True)}
    for name, (M, N, K, a_mn, b_mn) in mn.items():
        A = (torch.randn((K, M) if a_mn else (M, K), device=dev) * 0.1).to(torch.bfloat16)
        B = (torch.randn((K, N) if b_mn else (N, K), device=dev) * 0.1).to(torch.bfloat16)
        f32 = "wgrad" in name
        C = torch.zeros(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device=dev)
        st = _lib.stream_ptr()

        def ours():
            _lib.check(lib.prl_gemm_ex(A.data_ptr(), A.stride(0), int(a_mn), B.data_ptr(), B.stride(0), int(b_mn), M, N, K,
                                       C.data_ptr(), N, int(f32), int(f32), None, None, 0, 1.0, st))
        Am, Bm = (A.t() if a_mn else A), (B if b_mn else B.t())

        def cublas():
            torch.mm(Am, Bm)
        t_o, t_c = time_ms(ours), time_ms(cublas)
        fl = 2.0 * M * N * K
        out[name] = {"M": M, "N": N, "K": K, "ours_ms": round(t_o, 4), "ours_TFLOPs": round(fl / t_o / 1e9, 1),
                     "cublas_ms": round(t_c, 4), "cublas_TFLOPs": round(fl / t_c / 1e9, 1)}
        print(name, out[name], flush=True)
        del A, B, C
    x = torch.randn(T, 18944, device=dev).to(torch.bfloat16)
    y = torch.empty(18944, T, dtype=torch.bfloat16, device=dev)
    t = time_ms(lambda: _lib.check(lib.prl_transpose_bf16(x.data_ptr(), T, 18944, 18944, y.data_ptr(), T, _lib.stream_ptr())))
    out["transpose 16384x18944"] = {"ms": round(t, 4), "GBs": round(2 * x.numel() * 2 / t / 1e6, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

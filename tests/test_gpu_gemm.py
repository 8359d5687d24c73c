"""tcgen05 weight-streaming GEMM (csrc/gemm_tc.cu) against a plain fp32 torch matmul of the same bf16
operands.  Tolerance: fp32 accumulation of bf16 products -> 1e-4 relative to the row scale."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def run_gemm(W, X, W_lo=None, split_k=0):
    from pipelinerl_b200 import _lib
    lib = _lib.load()
    M, K = X.shape
    N = W.shape[0]
    if split_k == 0:
        split_k = lib.prl_gemm_auto_split_k(M, N, K)
    part = torch.full((split_k, M, N), float("nan"), dtype=torch.float32, device=X.device)
    _lib.check(lib.prl_gemm_bf16_splitk(W.data_ptr(), W_lo.data_ptr() if W_lo is not None else None, X.data_ptr(),
                                        M, N, K, split_k, part.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    return part


SHAPES = [
    (1, 128, 64, 1), (16, 256, 128, 2), (37, 300, 200, 1), (64, 4608, 3584, 0), (64, 3584, 18944, 0),
    (33, 1152, 512, 3), (128, 640, 1024, 2), (200, 384, 256, 1), (300, 256, 192, 1), (64, 37888, 3584, 0),
]


@pytest.mark.parametrize("M,N,K,split_k", SHAPES)
def test_gemm_matches_fp32_matmul(cuda_device, M, N, K, split_k):
    g = torch.Generator().manual_seed(M * 131 + N * 7 + K)
    X = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(cuda_device)
    W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(cuda_device)
    part = run_gemm(W, X, split_k=split_k)
    assert torch.isfinite(part).all(), "unwritten partial tile"
    got = part.sum(0)
    want = X.float() @ W.float().t()
    scale = want.abs().max().item()
    assert (got - want).abs().max().item() <= 2e-4 * scale + 1e-6


def test_gemm_hi_lo_is_fp32_equivalent(cuda_device):
    g = torch.Generator().manual_seed(3)
    M, N, K = 48, 1024, 1536
    Wf = (torch.randn(N, K, generator=g) * 0.02).to(cuda_device)
    hi = Wf.to(torch.bfloat16)
    lo = (Wf - hi.float()).to(torch.bfloat16)
    X = (torch.randn(M, K, generator=g)).to(torch.bfloat16).to(cuda_device)
    got = run_gemm(hi, X, W_lo=lo).sum(0)
    want = (X.double() @ Wf.double().t()).float()
    bf16_only = X.float() @ hi.float().t()
    err = (got - want).abs().max().item()
    err_bf16 = (bf16_only - want).abs().max().item()
    assert err <= 3e-5 * want.abs().max().item(), (err, err_bf16)
    assert err < err_bf16 / 20  # the residual stream really is used


def test_gemm_rejects_bad_arguments(cuda_device):
    from pipelinerl_b200 import _lib
    lib = _lib.load()
    X = torch.zeros(4, 60, dtype=torch.bfloat16, device=cuda_device)
    W = torch.zeros(128, 60, dtype=torch.bfloat16, device=cuda_device)
    out = torch.zeros(1, 4, 128, device=cuda_device)
    rc = lib.prl_gemm_bf16_splitk(W.data_ptr(), None, X.data_ptr(), 4, 128, 60, 1, out.data_ptr(), None)
    assert rc == -1 and b"K" in lib.prl_last_error()

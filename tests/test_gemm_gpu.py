"""tcgen05 GEMM (csrc/gemm_tc.cu, rl_gemm_bf16) against a plain PyTorch fp32 reference of the same
op on the same bf16-rounded operands.  Tolerance: bf16 inputs, fp32 accumulation -> the only
difference is summation order: rtol 2e-3 of the row scale (atol = 2e-3 * sqrt(K))."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]

EPI_F32, EPI_ATOMIC, EPI_BIAS_ELU_BF16, EPI_BIAS_F32, EPI_DELU_BF16, EPI_BF16 = range(6)


def gemm(A, B, C, bias=None, aux=None, db=None, M=None, N=None, K=None, transposed=0, epilogue=0, split_k=1):
    from rapid_locomotion_rl_b200 import _lib
    lib = _lib.lib()
    P = _lib.ptr
    _lib.check(lib.rl_gemm_bf16(P(A), P(B), P(C), P(bias), P(aux), P(db), M, N, K, A.stride(0), B.stride(0),
                                C.stride(0), 0 if aux is None else aux.stride(0), transposed, epilogue, split_k,
                                _lib.current_stream()))
    torch.cuda.synchronize()


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 64, 128), (1000, 512, 64), (24000, 256, 512),
                                   (333, 18, 128), (4000, 12, 128), (129, 1, 128), (4096, 256, 640), (77, 32, 24)])
def test_nt_plain(M, N, K):
    torch.manual_seed(M + N + K)
    A, B = rnd(M, K), rnd(N, K)
    C = torch.full((M, N + 3), 7.0, device="cuda")[:, :N] if N % 4 else torch.empty(M, N, device="cuda")
    ldc_ok = C.stride(0)
    gemm(A, B, C, M=M, N=N, K=K)
    ref = A.float() @ B.float().t()
    torch.testing.assert_close(C, ref, rtol=2e-3, atol=2e-3 * K ** 0.5)
    assert ldc_ok == C.stride(0)


def test_nt_padded_pitch_and_k_tail():
    """K not a multiple of 64 (TMA zero fill) and pitches wider than the logical extents."""
    M, N, K = 500, 96, 60
    A_full, B_full = rnd(M, 64), rnd(N, 64)
    C = torch.empty(M, N, device="cuda")
    gemm(A_full, B_full, C, M=M, N=N, K=K)
    ref = A_full[:, :K].float() @ B_full[:, :K].float().t()
    torch.testing.assert_close(C, ref, rtol=2e-3, atol=2e-2)


def test_nt_bias_elu_bf16_and_bias_f32():
    M, N, K = 3000, 512, 64
    A, B, bias = rnd(M, K), rnd(N, K, scale=0.2), torch.randn(N, device="cuda")
    ref = A.float() @ B.float().t() + bias
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    gemm(A, B, C, bias=bias, M=M, N=N, K=K, epilogue=EPI_BIAS_ELU_BF16)
    torch.testing.assert_close(C.float(), torch.nn.functional.elu(ref), rtol=1e-2, atol=2e-2)
    C2 = torch.empty(M, N, device="cuda")
    gemm(A, B, C2, bias=bias, M=M, N=N, K=K, epilogue=EPI_BIAS_F32)
    torch.testing.assert_close(C2, ref, rtol=2e-3, atol=2e-2)


def test_nt_misaligned_bias_and_offset_views():
    """Bias pointers are 4 B-aligned views of the flat parameter buffer; outputs may be column slices."""
    M, N, K = 700, 256, 128
    A, B = rnd(M, K), rnd(N, K, scale=0.2)
    bias = torch.randn(N + 3, device="cuda")[3:]          # 12 B offset
    big = torch.zeros(M, 1024, device="cuda", dtype=torch.bfloat16)
    C = big[:, 512:512 + N]                                # column slice, pitch 1024
    gemm(A, B, C, bias=bias, M=M, N=N, K=K, epilogue=EPI_BIAS_ELU_BF16)
    ref = torch.nn.functional.elu(A.float() @ B.float().t() + bias)
    torch.testing.assert_close(C.float(), ref, rtol=1e-2, atol=2e-2)
    assert (big[:, :512] == 0).all() and (big[:, 512 + N:] == 0).all()
    # unaligned bf16 destination (direct epilogue path): 42-column offset, N = 18
    dst = torch.zeros(M, 64, device="cuda", dtype=torch.bfloat16)
    B18 = rnd(18, K, scale=0.2)
    from rapid_locomotion_rl_b200 import _lib
    lib = _lib.lib()
    _lib.check(lib.rl_gemm_bf16(A.data_ptr(), B18.data_ptr(), dst.data_ptr() + 42 * 2, bias.data_ptr(), None, None, M, 18, K,
                                K, K, 64, 0, 0, 6, 1, _lib.current_stream()))
    torch.cuda.synchronize()
    torch.testing.assert_close(dst[:, 42:60].float(), A.float() @ B18.float().t() + bias[:18], rtol=1e-2, atol=2e-2)
    assert (dst[:, :42] == 0).all() and (dst[:, 60:] == 0).all()


def test_nt_delu_epilogue():
    """dgrad through an ELU: dX = (dY W) * elu'(x) with elu' taken from the stored ELU output."""
    M, N, K = 2000, 256, 128      # dX [M,N] = dY [M,K] * Wt [N,K]^T
    dY, Wt = rnd(M, K), rnd(N, K, scale=0.2)
    pre = torch.randn(M, N, device="cuda")
    y = torch.nn.functional.elu(pre).to(torch.bfloat16)
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    gemm(dY, Wt, C, aux=y, M=M, N=N, K=K, epilogue=EPI_DELU_BF16)
    yf = y.float()
    ref = (dY.float() @ Wt.float().t()) * torch.where(yf > 0, torch.ones_like(yf), yf + 1)
    torch.testing.assert_close(C.float(), ref, rtol=1e-2, atol=3e-2)


@pytest.mark.parametrize("M,N,K,split", [(512, 64, 24000, 16), (256, 512, 4096, 4), (128, 256, 1000, 1),
                                         (12, 128, 3000, 8), (18, 32, 777, 3), (1, 128, 2048, 2), (256, 640, 9000, 5)])
def test_tn_wgrad_splitk_and_db(M, N, K, split):
    """dW[M,N] = dY[K,M]^T X[K,N] with split-K atomics, db[m] = sum_k dY[k,m] from the ones-MMA."""
    torch.manual_seed(K)
    Mp, Np = (M + 7) // 8 * 8, (N + 7) // 8 * 8
    dY, X = rnd(K, Mp, scale=0.1), rnd(K, Np)
    C = torch.zeros(M, N, device="cuda")
    db = torch.zeros(M, device="cuda")
    gemm(dY, X, C, db=db, M=M, N=N, K=K, transposed=1, epilogue=EPI_ATOMIC if split > 1 else EPI_F32, split_k=split)
    ref = dY[:, :M].float().t() @ X[:, :N].float()
    torch.testing.assert_close(C, ref, rtol=3e-3, atol=3e-3 * K ** 0.5 * 0.1)
    torch.testing.assert_close(db, dY[:, :M].float().sum(0), rtol=3e-3, atol=3e-3 * K ** 0.5 * 0.1)


def test_bad_arguments():
    from rapid_locomotion_rl_b200 import _lib
    A, B = rnd(128, 64), rnd(128, 64)
    C = torch.empty(128, 128, device="cuda")
    with pytest.raises(_lib.RlError):
        gemm(A, B, C, M=128, N=128, K=64, epilogue=EPI_BIAS_F32)            # bias missing
    with pytest.raises(_lib.RlError):
        gemm(A[:, 1:], B, C, M=128, N=128, K=63)                            # misaligned base
    with pytest.raises(_lib.RlError):
        gemm(A, B, C, M=128, N=128, K=64, split_k=2)                        # split-K without atomics


@pytest.mark.parametrize("K", [64, 1000, 24000])
def test_wgrad_grouped_vs_torch(K):
    """rl_wgrad_grouped (persistent kernel: all layers' dW / db in one launch) against fp32 torch on the
    same bf16 operands: dW += dY^T X, db += column sums of dY; shapes of the learner's layers."""
    import ctypes as C
    from rapid_locomotion_rl_b200 import _lib
    lib = _lib.lib()
    _lib.check(lib.rl_gemm_init())
    torch.manual_seed(K)
    shapes = [(256, 512), (128, 256), (12, 128), (1, 128), (1024, 60), (256, 18), (18, 32), (256, 630)]
    probs, keep = [], []
    for M, N in shapes:
        ldy, ldx = (M + 7) // 8 * 8, (N + 7) // 8 * 8
        dY = torch.zeros(K, ldy, dtype=torch.bfloat16, device="cuda"); dY[:, :M] = torch.randn(K, M, device="cuda") * 0.1
        X = torch.zeros(K, ldx, dtype=torch.bfloat16, device="cuda"); X[:, :N] = torch.randn(K, N, device="cuda")
        dW = torch.randn(M, N, device="cuda") * 0.01          # accumulates on top of existing content
        db = torch.randn(M, device="cuda") * 0.01
        ref_w = dW + dY[:, :M].float().t() @ X[:, :N].float()
        ref_b = db + dY[:, :M].float().sum(0)
        q = _lib.RlWgradProblem()
        q.dY, q.X, q.dW, q.db = dY.data_ptr(), X.data_ptr(), dW.data_ptr(), db.data_ptr()
        q.M, q.N, q.K, q.ld_dy, q.ld_x, q.ld_dw, q.split_k = M, N, K, ldy, ldx, N, 0
        probs.append(q); keep.append((dY, X, dW, db, ref_w, ref_b))
    arr = (_lib.RlWgradProblem * len(probs))(*probs)
    _lib.check(lib.rl_wgrad_grouped(arr, len(probs), _lib.current_stream()))
    torch.cuda.synchronize()
    for (M, N), (dY, X, dW, db, ref_w, ref_b) in zip(shapes, keep):
        tol = 2e-3 * (K ** 0.5)
        torch.testing.assert_close(dW, ref_w, rtol=2e-3, atol=tol, msg="dW %dx%d" % (M, N))
        torch.testing.assert_close(db, ref_b, rtol=2e-3, atol=tol, msg="db %d" % M)

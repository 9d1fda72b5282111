"""Torch restatement of the Fs2Gemm descriptor semantics (include/fs2_b200.h) used by the GPU GEMM tests."""
import torch


def _rows(mat, idx):
    """mat[idx] with out-of-range rows reading as zero."""
    ok = (idx >= 0) & (idx < mat.shape[0])
    out = mat[idx.clamp(0, mat.shape[0] - 1)]
    return out * ok[:, None].to(out.dtype)


def _cols(mat, idx):
    ok = (idx >= 0) & (idx < mat.shape[1])
    out = mat[:, idx.clamp(0, mat.shape[1] - 1)]
    return out * ok[None, :].to(out.dtype)


def ref_gemm(mode, M, N, K, taps, A, B, a_row_off=0, a_tap_step=0, b_row_off=0, b_tap_step=0):
    """A, B: 2-D float tensors [rows, inner].  Returns mode 0/1: [M, N]; mode 2: [M, taps, N]."""
    dev = A.device
    A = A.double()
    B = B.double()
    m = torch.arange(M, device=dev)
    n = torch.arange(N, device=dev)
    k = torch.arange(K, device=dev)
    if mode == 0:
        out = torch.zeros(M, N, dtype=torch.float64, device=dev)
        for j in range(taps):
            a = _cols(_rows(A, a_row_off + m + j * a_tap_step), k)          # [M, K]
            b = _cols(_rows(B, n), j * b_tap_step + k)                      # [N, K]
            out += a @ b.t()
        return out
    if mode == 1:
        out = torch.zeros(M, N, dtype=torch.float64, device=dev)
        for j in range(taps):
            a = _cols(_rows(A, a_row_off + m + j * a_tap_step), k)          # [M, K]
            b = _cols(_rows(B, b_row_off + k), n + j * b_tap_step)          # [K, N]
            out += a @ b
        return out
    out = torch.zeros(M, taps, N, dtype=torch.float64, device=dev)
    for j in range(taps):
        a = _cols(_rows(A, a_row_off + k), m)                               # [K, M]
        b = _cols(_rows(B, b_row_off + k + j * b_tap_step), n)              # [K, N]
        out[:, j] = a.t() @ b
    return out

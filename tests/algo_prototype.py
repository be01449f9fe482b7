"""NumPy model of the decomposition the CUDA kernels use.

Not the oracle: this restates *our* algorithm (column pass / row pass with the
block-constant correction) step by step so that (a) the algebra is checked against the
oracle on CPU and (b) GPU intermediates (b2f_debug_* entry points) can be compared stage
by stage.

Block of M = R*L dual-pol samples z[n] = xP[n] + i xQ[n], n = n1 + R*n2, R = 2*nchan:
  column pass (per n1):   A[k2]   = FFT_L over n2 of z[n1 + R n2]
                          S[n1]   = A[0]                       (column sums)
                          B[m]    = IFFT_L over k2 of A[k2] * W_M^(k2*n1)   (unnormalised)
  row pass (per m):       Zc[m]   = FFT_R over n1 of B[m][n1],   c = 0..R-1
  un-mix (c < nchan):     a = Zc,  b = Z_(R-1-c),  eps_c = conj(G[R-1-c] - G[(R-c) mod R]),
                          G = FFT_R(S);  yP = (a + conj(b) - eps)/2,  yQ = (a - conj(b) + eps)/(2i)
"""
import numpy as np


def column_pass(z_blk: np.ndarray, R: int, L: int):
    """z_blk[M] complex -> B[m, n1] complex (L x R), S[n1]"""
    M = R * L
    zz = z_blk.reshape(L, R)                      # [n2, n1]
    A = np.fft.fft(zz, axis=0)                    # [k2, n1]
    S = A[0].copy()
    k2 = np.arange(L)[:, None]
    n1 = np.arange(R)[None, :]
    tw = np.exp(-2j * np.pi * (k2 * n1) / M)
    B = np.fft.ifft(A * tw, axis=0) * L           # [m, n1]
    return B, S


def eps_from_colsum(S: np.ndarray, R: int) -> np.ndarray:
    G = np.fft.fft(S)
    c = np.arange(R // 2)
    return np.conj(G[R - 1 - c] - G[(R - c) % R])


def row_pass(B: np.ndarray, eps: np.ndarray, R: int):
    """B[m, n1] -> yP[m, c], yQ[m, c] for c < R/2"""
    Z = np.fft.fft(B, axis=1)                     # [m, c]
    N = R // 2
    a = Z[:, :N]
    b = Z[:, ::-1][:, :N]                         # Z[R-1-c]
    bp = np.conj(b) - eps[None, :]
    yP = (a + bp) / 2
    yQ = (a - bp) / 2j
    return yP, yQ


def filterbank_dualpol(x: np.ndarray, nchan: int, freq_res: int):
    """x[2, nsamp] -> (yP, yQ) each [t, chan], same contract as oracle.filterbank per pol."""
    R, L = 2 * nchan, freq_res
    M = R * L
    nblk = x.shape[1] // M
    z = x[0, : nblk * M] + 1j * x[1, : nblk * M]
    outP, outQ = [], []
    for b in range(nblk):
        B, S = column_pass(z[b * M: (b + 1) * M], R, L)
        yP, yQ = row_pass(B, eps_from_colsum(S, R), R)
        outP.append(yP)
        outQ.append(yQ)
    return np.concatenate(outP), np.concatenate(outQ)

"""NumPy model of the factored tensor-core DFT dataflow used by csrc/logmel.cu.

N = 32768 = 256 (n1) x 128 (n2).  Frame g[n], n = 128*n1 + n2;  bin k = k1 + 256*k2.

  stage 1 (GEMM, constants as A):   Dc[k1,n2] = sum_n1 cos(2pi k1 n1/256) g[128 n1+n2]      k1 in [0,128)
                                    Ds[k1,n2] = sum_n1 -sin(2pi k1 n1/256) g[128 n1+n2]     row 0 := (-1)^n1  (= Y[n2,128])
  twiddle (CUDA cores):             Z[k1,n2]  = (Dc + i Ds) * exp(-2pi i k1 n2 / 32768)
  stage 2 (GEMM, constants as B):   X[k1+256 k2] = sum_n2 Z[k1,n2] exp(-2pi i n2 k2/128)    k2 in [0,128)
  row 128 (CUDA cores):             X[128+256 k2] = sum_n2 Y[n2,128] exp(-2pi i n2 (2 k2+1)/256), k2 in [0,64)
  power + bin map:                  k2<64 -> bin k1+256k2 ; k2>=64 -> bin 32768-(k1+256k2) (Hermitian mirror)

This file is test infrastructure: it pins the index algebra of the kernel against np.fft.rfft.
"""
import numpy as np

N, N1, N2 = 32768, 256, 128


def bf16_round(x):
    """Round-to-nearest-even float32 -> bfloat16 (returned as float32)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def fp16_round(x):
    return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)


def split2(x, rnd):
    x = np.asarray(x, dtype=np.float32)
    hi = rnd(x)
    lo = rnd(x - hi)
    return hi, lo


def mm3(a, b, rnd):
    """Split-operand product hi*hi + lo*hi + hi*lo accumulated in float64 (stands in for fp32 TMEM)."""
    if rnd is None:
        return a.astype(np.float64) @ b.astype(np.float64)
    ah, al = split2(a, rnd)
    bh, bl = split2(b, rnd)
    f = np.float64
    return ah.astype(f) @ bh.astype(f) + al.astype(f) @ bh.astype(f) + ah.astype(f) @ bl.astype(f)


def stage1_constants():
    k1 = np.arange(128)[:, None]
    n1 = np.arange(N1)[None, :]
    ang = 2 * np.pi * ((k1 * n1) % N1) / N1
    C = np.cos(ang)
    S = -np.sin(ang)
    S[0, :] = (-1.0) ** np.arange(N1)          # row 0 of the sine block carries k1 = 128
    return C, S


def stage2_constants():
    n2 = np.arange(N2)[:, None]
    k2 = np.arange(N2)[None, :]
    ang = 2 * np.pi * ((n2 * k2) % N2) / N2
    return np.cos(ang), np.sin(ang)            # exp(-i a) = cos a - i sin a


def factored_power_spectrum(g, rnd=None):
    """g: windowed, zero-padded frame (32768,) -> power spectrum P[0..16384] via the kernel dataflow."""
    X1 = np.asarray(g, dtype=np.float32).reshape(N1, N2)            # [n1, n2]
    C, S = stage1_constants()
    Dc = mm3(C.astype(np.float32), X1, rnd)                         # [k1, n2]
    Ds = mm3(S.astype(np.float32), X1, rnd)
    v = Ds[0].copy()                                                # Y[n2, 128]
    Y = Dc + 1j * Ds
    Y[0] = Dc[0]                                                    # k1 = 0 is purely real
    k1 = np.arange(128)[:, None]
    n2 = np.arange(N2)[None, :]
    Z = Y * np.exp(-2j * np.pi * (k1 * n2) / N)
    Zr = Z.real.astype(np.float32)
    Zi = Z.imag.astype(np.float32)
    C2, S2 = stage2_constants()
    re = mm3(Zr, C2.astype(np.float32), rnd) + mm3(Zi, S2.astype(np.float32), rnd)    # [k1, k2]
    im = mm3(Zi, C2.astype(np.float32), rnd) - mm3(Zr, S2.astype(np.float32), rnd)
    P2 = re * re + im * im
    X2 = re + 1j * im
    P = np.zeros(N // 2 + 1)
    X = np.zeros(N // 2 + 1, dtype=np.complex128)
    for kk1 in range(128):
        for kk2 in range(128):
            k = kk1 + 256 * kk2
            if k <= N // 2:
                P[k] = P2[kk1, kk2]
                X[k] = X2[kk1, kk2]
            elif kk1 >= 1:
                P[N - k] = P2[kk1, kk2]
                X[N - k] = np.conj(X2[kk1, kk2])
    # row 128
    kk2 = np.arange(64)[:, None]
    nn2 = np.arange(N2)[None, :]
    V = (v[None, :] * np.exp(-2j * np.pi * ((nn2 * (2 * kk2 + 1)) % 256) / 256)).sum(1)
    P[128 + 256 * np.arange(64)] = np.abs(V) ** 2
    X[128 + 256 * np.arange(64)] = V
    return P, X


# ---------------------------------------------------------------------------------------------------------
# v2 dataflow: stage 1 with the even/odd fold over n1 (K = 128), stage 2 with one radix-2 DIF step over n2
# (two 64-point complex DFT GEMMs sharing one constant matrix).
#
#   U[m,n2] = X[m,n2] + X[256-m,n2] (m=1..127), U[0] = X[0];  V[m,n2] = X[m,n2] - X[256-m,n2], V[0] = 0
#   Dc[k1,n2] = sum_m cos(2 pi k1 m/256) U[m,n2] + (-1)^k1 X[128,n2]          (GEMM + CUDA-core correction)
#   Ds[k1,n2] = sum_m -sin(2 pi k1 m/256) V[m,n2]
#   Y128[n2]  = sum_m (-1)^m U[m,n2] + X[128,n2]                               (CUDA cores, producer side)
#   Z = (Dc + i Ds) exp(-2 pi i k1 n2/32768)
#   E[k1,n] = Z[k1,n] + Z[k1,n+64];   O[k1,n] = (Z[k1,n] - Z[k1,n+64]) exp(-2 pi i n/128)     n in [0,64)
#   X[k1 + 256 (2j)]   = sum_n E[k1,n] exp(-2 pi i n j/64)
#   X[k1 + 256 (2j+1)] = sum_n O[k1,n] exp(-2 pi i n j/64)
def stage1_fold_constants():
    k1 = np.arange(128)[:, None]
    m = np.arange(128)[None, :]
    ang = 2 * np.pi * ((k1 * m) % N1) / N1
    return np.cos(ang), -np.sin(ang)


def stage2_radix2_constants():
    """B matrices [K = 64 (n), N = 128 (64 re | 64 im)] for the real and imaginary parts of the A operand."""
    n = np.arange(64)[:, None]
    j = np.arange(64)[None, :]
    ang = 2 * np.pi * ((n * j) % 64) / 64
    c, s = np.cos(ang), np.sin(ang)
    b_re = np.concatenate([c, -s], axis=1)      # multiplies Er: re += Er c, im += -Er s
    b_im = np.concatenate([s, c], axis=1)       # multiplies Ei: re += Ei s, im +=  Ei c
    return b_re, b_im


def factored_power_spectrum_v2(g, rnd=None, scale=1.0):
    X1 = (np.asarray(g, dtype=np.float64) * scale).astype(np.float32).reshape(N1, N2)
    U = X1[:128].copy()
    V = np.zeros_like(U)
    U[1:] = X1[1:128] + X1[255:128:-1]
    V[1:] = X1[1:128] - X1[255:128:-1]
    C, S = stage1_fold_constants()
    sign_k1 = (-1.0) ** np.arange(128)[:, None]
    Dc = mm3(C.astype(np.float32), U, rnd) + sign_k1 * X1[128][None, :].astype(np.float64)
    Ds = mm3(S.astype(np.float32), V, rnd)
    y128 = ((-1.0) ** np.arange(128)[:, None] * U.astype(np.float64)).sum(0) + X1[128]
    k1 = np.arange(128)[:, None]
    n2 = np.arange(N2)[None, :]
    Z = (Dc + 1j * Ds) * np.exp(-2j * np.pi * (k1 * n2) / N)
    E = Z[:, :64] + Z[:, 64:]
    O = (Z[:, :64] - Z[:, 64:]) * np.exp(-2j * np.pi * np.arange(64)[None, :] / 128)
    b_re, b_im = stage2_radix2_constants()
    f32 = np.float32

    def dft64(A):
        D = mm3(A.real.astype(f32), b_re.astype(f32), rnd) + mm3(A.imag.astype(f32), b_im.astype(f32), rnd)
        return D[:, :64] + 1j * D[:, 64:]

    Xe, Xo = dft64(E), dft64(O)
    X = np.zeros(N // 2 + 1, dtype=np.complex128)
    for kk1 in range(128):
        for j in range(64):
            for par, val in ((0, Xe[kk1, j]), (1, Xo[kk1, j])):
                k = kk1 + 256 * (2 * j + par)
                if k <= N // 2:
                    X[k] = val
                elif kk1 >= 1:
                    X[N - k] = np.conj(val)
    kk2 = np.arange(64)[:, None]
    nn2 = np.arange(N2)[None, :]
    X[128 + 256 * np.arange(64)] = (y128[None, :] * np.exp(-2j * np.pi * ((nn2 * (2 * kk2 + 1)) % 256) / 256)).sum(1)
    X = X / scale
    return np.abs(X) ** 2, X

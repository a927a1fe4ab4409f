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

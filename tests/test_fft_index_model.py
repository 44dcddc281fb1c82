"""Host-side model of the index arithmetic of the STFT kernel
(amt-saga_b200/csrc/stft.cu: dif_pass, zpos, real-FFT split).  It replays the
kernel's formulas in numpy (complex128) and checks them against numpy.fft, so
the digit-reversal / twiddle-table layout is proven before any GPU run."""
import numpy as np
import pytest

SHAPES = {128: (16, 8, 1), 256: (16, 16, 1), 512: (32, 16, 1), 1024: (32, 32, 1),
          2048: (16, 16, 8), 4096: (16, 16, 16)}


def bitrev(i, R):
    r, b = 0, 1
    while b < R:
        r = (r << 1) | (i & 1)
        i >>= 1
        b <<= 1
    return r


def small_dif(v):
    """radix-2 DIF exactly as fft_reg: result in bit-reversed order."""
    R = len(v)
    v = list(v)
    half = R // 2
    while half >= 1:
        for b in range(0, R, 2 * half):
            for k in range(half):
                a, c = v[b + k], v[b + k + half]
                v[b + k] = a + c
                v[b + k + half] = (a - c) * np.exp(-2j * np.pi * (k * (16 // half)) / 32)
        half //= 2
    return v


def dif_pass(buf, M, L, R, tw, last):
    LS = L // R
    for q in range(M // R):
        b, j = divmod(q, LS)
        base = b * L + j
        v = small_dif([buf[base + r * LS] for r in range(R)])
        for i in range(R):
            rp = bitrev(i, R)
            o = v[i]
            if not last and rp > 0:
                o = o * tw[rp * LS + j]
            buf[base + rp * LS] = o


def zpos(k, M, R0, R1, R2):
    L1 = M // R0
    L2 = L1 // R1
    if R2 == 1:
        return (k % R0) * L1 + (k // R0)
    return (k % R0) * L1 + ((k // R0) % R1) * L2 + (k // (R0 * R1))


def tw_table(L, R):
    LS = L // R
    t = np.zeros(L, dtype=np.complex128)
    for rp in range(R):
        for j in range(LS):
            t[rp * LS + j] = np.exp(-2j * np.pi * ((j * rp) % L) / L)
    return t


@pytest.mark.parametrize("M", sorted(SHAPES))
def test_kernel_index_model_matches_rfft(M):
    R0, R1, R2 = SHAPES[M]
    N = 2 * M
    rng = np.random.default_rng(M)
    x = rng.standard_normal(N)
    buf = (x[0::2] + 1j * x[1::2]).astype(np.complex128)
    L1, L2 = M // R0, M // R0 // R1
    dif_pass(buf, M, M, R0, tw_table(M, R0), False)
    dif_pass(buf, M, L1, R1, tw_table(L1, R1) if R2 > 1 else None, R2 == 1)
    if R2 > 1:
        dif_pass(buf, M, L2, R2, None, True)
    Z = np.array([buf[zpos(k, M, R0, R1, R2)] for k in range(M)])
    assert np.allclose(Z, np.fft.fft(x[0::2] + 1j * x[1::2]), atol=1e-9)
    X = np.zeros(M + 1, dtype=np.complex128)
    for k in range(M // 2 + 1):
        zk, zm = Z[k], Z[(M - k) & (M - 1)]
        E = complex(0.5 * (zk.real + zm.real), 0.5 * (zk.imag - zm.imag))
        O = complex(0.5 * (zk.imag + zm.imag), -0.5 * (zk.real - zm.real))
        Tw = np.exp(-2j * np.pi * k / N) * O
        X[k] = E + Tw
        X[M - k] = np.conj(E - Tw)
    assert np.allclose(X, np.fft.rfft(x), atol=1e-9)


@pytest.mark.parametrize("M", [128, 1024, 2048])
def test_inverse_rebuild_model_matches_irfft(M):
    """istft_kernel: rebuild conj(Z) from the half spectrum, run the forward
    passes, read x[2n] = Re/M, x[2n+1] = -Im/M at zpos(n)."""
    R0, R1, R2 = SHAPES[M]
    N = 2 * M
    rng = np.random.default_rng(M + 1)
    X = rng.standard_normal(M + 1) + 1j * rng.standard_normal(M + 1)
    buf = np.zeros(M, dtype=np.complex128)
    for k in range(M // 2 + 1):
        xk, xm = X[k], X[M - k]
        if k == 0:
            xk, xm = complex(xk.real, 0), complex(xm.real, 0)
        E = complex(0.5 * (xk.real + xm.real), 0.5 * (xk.imag - xm.imag))
        D = complex(0.5 * (xk.real - xm.real), 0.5 * (xk.imag + xm.imag))
        w = np.exp(-2j * np.pi * k / N)
        O = complex(D.real * w.real + D.imag * w.imag, D.imag * w.real - D.real * w.imag)
        buf[k] = complex(E.real - O.imag, -(E.imag + O.real))
        if k != 0 and k != M // 2:
            buf[M - k] = complex(E.real + O.imag, -(O.real - E.imag))
    L1, L2 = M // R0, M // R0 // R1
    dif_pass(buf, M, M, R0, tw_table(M, R0), False)
    dif_pass(buf, M, L1, R1, tw_table(L1, R1) if R2 > 1 else None, R2 == 1)
    if R2 > 1:
        dif_pass(buf, M, L2, R2, None, True)
    x = np.zeros(N)
    for n in range(N):
        z = buf[zpos(n >> 1, M, R0, R1, R2)]
        x[n] = (-z.imag if n & 1 else z.real) / M
    assert np.allclose(x, np.fft.irfft(X, n=N), atol=1e-10)

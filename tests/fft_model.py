"""Host model (numpy, complex128) of the data flow of the windowed-correlation kernels in
ffsubsync_b200/csrc/corr.cu: in-place mixed-radix decimation-in-frequency passes that leave
the spectrum in digit-reversed ("position") order, the real-FFT untangle done directly in
that order, spectrum accumulation over sub blocks, the retangle and the decimation-in-time
inverse.  It exists so the index math can be checked against np.fft on the CPU; it is test
infrastructure, not product code."""
import numpy as np


class Plan:
    def __init__(self, radices):
        self.radices = list(radices)
        self.M = int(np.prod(radices))
        self.P = 2 * self.M

    # position <-> frequency maps for the in-place DIF output
    def freq_of_pos(self):
        M = self.M
        p = np.arange(M)
        f = np.zeros(M, dtype=np.int64)
        weight = 1          # weight of the current digit inside f (first radix = least significant)
        span = M
        for r in self.radices:
            sub = span // r
            digit = (p // sub) % r
            f += digit * weight
            weight *= r
            span = sub
        return f

    def pos_of_freq(self):
        f_of_p = self.freq_of_pos()
        inv = np.empty_like(f_of_p)
        inv[f_of_p] = np.arange(self.M)
        return inv


def dif_forward(plan: Plan, x: np.ndarray) -> np.ndarray:
    """Natural-order input -> DFT_M(x) stored at digit-reversed positions."""
    M = plan.M
    x = x.astype(np.complex128).copy()
    span = M
    for r in plan.radices:
        sub = span // r
        v = x.reshape(M // span, r, sub)                     # [block, q, j]
        k = np.arange(r)
        dft = np.exp(-2j * np.pi * np.outer(k, k) / r)       # [k, q]
        y = np.einsum("kq,bqj->bkj", dft, v)
        tw = np.exp(-2j * np.pi * np.outer(k, np.arange(sub)) / span)  # [k, j] = w_span^(j k)
        x = (y * tw[None]).reshape(M)
        span = sub
    return x


def dit_inverse(plan: Plan, X: np.ndarray) -> np.ndarray:
    """Digit-reversed spectrum -> natural-order M * IDFT (unnormalised)."""
    M = plan.M
    x = X.astype(np.complex128).copy()
    spans = []
    span = M
    for r in plan.radices:
        spans.append((span, r))
        span //= r
    for span, r in reversed(spans):
        sub = span // r
        v = x.reshape(M // span, r, sub)                     # [block, k, j]
        k = np.arange(r)
        tw = np.exp(+2j * np.pi * np.outer(k, np.arange(sub)) / span)
        idft = np.exp(+2j * np.pi * np.outer(k, k) / r)      # [q, k]
        y = np.einsum("qk,bkj->bqj", idft, v * tw[None])
        x = y.reshape(M)
    return x


def untangle(plan: Plan, Z: np.ndarray):
    """Z = position-order DFT_M of z[n] = x[2n] + i x[2n+1].  Returns the half spectrum of the
    real length-P sequence x in position order: H[p] = X[f(p)] for f >= 1, and slot 0 packed as
    (X[0], X[M]) in (re, im)."""
    M, P = plan.M, plan.P
    f = plan.freq_of_pos()
    pos = plan.pos_of_freq()
    partner = pos[(M - f) % M]
    Zf, Zg = Z, np.conj(Z[partner])
    w = np.exp(-2j * np.pi * f / P)
    H = 0.5 * (Zf + Zg) - 0.5j * w * (Zf - Zg)
    H[0] = (Z[0].real + Z[0].imag) + 1j * (Z[0].real - Z[0].imag)
    return H


def spectrum_product(plan: Plan, A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """conj(A) * B on packed half spectra (slot 0 holds two real bins)."""
    C = np.conj(A) * B
    C[0] = A[0].real * B[0].real + 1j * (A[0].imag * B[0].imag)
    return C


def retangle(plan: Plan, C: np.ndarray) -> np.ndarray:
    """Packed half spectrum (position order) -> position-order spectrum Zc whose inverse
    M-point transform is c[2n] + i c[2n+1] (times 2M: normalisation is applied at the end)."""
    M, P = plan.M, plan.P
    f = plan.freq_of_pos()
    pos = plan.pos_of_freq()
    partner = pos[(M - f) % M]
    Cf, Cg = C, np.conj(C[partner])
    w = np.exp(+2j * np.pi * f / P)
    Zc = (Cf + Cg) + 1j * w * (Cf - Cg)
    c0, cM = C[0].real, C[0].imag
    Zc[0] = (c0 + cM) + 1j * (c0 - cM)
    return Zc


def window_correlation(plan: Plan, ref_p: np.ndarray, sub_p: np.ndarray, o_min: int, W: int) -> np.ndarray:
    """scores[m] = sum_j sub_p[j] * ref_p[j + o_min + m], m in [0, W), out-of-range = 0, computed
    the way the kernels do: L = P - W + 1 samples of sub per block, P samples of ref per block."""
    P, M = plan.P, plan.M
    assert 1 <= W <= P
    L = P - W + 1
    S, R = len(sub_p), len(ref_p)
    acc = np.zeros(M, dtype=np.complex128)
    for j0 in range(0, S, L):
        a = np.zeros(P)
        seg = sub_p[j0:j0 + L]
        a[:len(seg)] = seg
        b = np.zeros(P)
        i0 = j0 + o_min
        lo, hi = max(i0, 0), min(i0 + P, R)
        if hi > lo:
            b[lo - i0:hi - i0] = ref_p[lo:hi]
        A = untangle(plan, dif_forward(plan, a[0::2] + 1j * a[1::2]))
        B = untangle(plan, dif_forward(plan, b[0::2] + 1j * b[1::2]))
        acc += spectrum_product(plan, A, B)
    z = dit_inverse(plan, retangle(plan, acc)) / (2.0 * M)
    c = np.empty(P)
    c[0::2], c[1::2] = z.real, z.imag
    return c[:W]

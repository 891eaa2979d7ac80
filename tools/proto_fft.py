"""numpy emulation of csrc/fft.cuh (lane x register layout) to validate the index algebra."""
import numpy as np

def brev5(x):
    return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4)

W32 = np.exp(-2j * np.pi * np.arange(16) / 32)
CM = np.exp(1j * np.pi * np.arange(32) / 64)
l_ = np.arange(32)
TW = np.array([[np.exp(1j * np.pi * ((l * (1 - 4 * k1)) % 4096) / 2048) for l in range(32)] for k1 in range(32)])  # [k1][l]

def fft32_dit(x, inv):  # x: [lane][32] complex, in-place DIT, bit-reversed in -> natural out
    x = x.copy()
    half = 1
    while half < 32:
        for base in range(0, 32, 2 * half):
            for t in range(half):
                w = W32[t * (16 // half)]
                if inv: w = np.conj(w)
                a = x[:, base + t].copy(); b = x[:, base + t + half] * w
                x[:, base + t] = a + b; x[:, base + t + half] = a - b
        half *= 2
    return x

def fwd1024(xin):  # xin[lane][brev5(m)] already twisted by C_m
    x = fft32_dit(xin, False)                    # x[l][k1]
    y = x * TW.T                                 # y[l][k1] = x[l][k1]*TW[k1][l]
    tb = y                                       # tbuf[l][k1]
    x2 = np.zeros_like(x)
    for l in range(32):
        x2[:, brev5(l)] = tb[l, :]               # lane k1 reads tbuf[l][k1]
    return fft32_dit(x2, False)                  # [lane=k1][k2]

def inv1024(xin):  # xin[lane=k1][brev5(k2)]
    x = fft32_dit(xin, True)                     # x[k1][l]
    tb = x                                       # tbuf[k1][l]
    x2 = np.zeros_like(x)
    for k1 in range(32):
        x2[:, brev5(k1)] = tb[k1, :] * np.conj(TW[k1, :])   # lane l reads tbuf[k1][l]
    return fft32_dit(x2, True)                   # [lane=l][m]

def to_home(poly):  # poly[2048] -> folded complex [lane][m]
    z = poly[:1024] + 1j * poly[1024:]
    return z.reshape(32, 32).T                   # [l][m] = z[l + 32 m]

def forward(poly):
    h = to_home(poly.astype(np.float64))
    x = np.zeros((32, 32), complex)
    for m in range(32):
        x[:, brev5(m)] = h[:, m] * CM[m]
    return fwd1024(x)                            # [lane][q]

def backward(F):  # F[lane][q] -> real poly (float), includes 1/1024
    x = np.zeros((32, 32), complex)
    for q in range(32):
        x[:, brev5(q)] = F[:, q]
    y = inv1024(x)
    for m in range(32):
        y[:, m] = y[:, m] * np.conj(CM[m])
    y = y / 1024
    z = y.T.reshape(1024)                        # z[l+32m]
    return np.concatenate([z.real, z.imag])

def negacyclic(a, b):
    n = len(a); out = np.zeros(n, dtype=object)
    for i in range(n):
        for j in range(n):
            k = i + j
            if k < n: out[k] += int(a[i]) * int(b[j])
            else: out[k - n] -= int(a[i]) * int(b[j])
    return out

if __name__ == "__main__":
    rng = np.random.default_rng(1)
    a = rng.integers(-2**10, 2**10, 2048); b = rng.integers(-2**10, 2**10, 2048)
    got = backward(forward(a) * forward(b))
    # fast exact reference via numpy polynomial convolution
    full = np.convolve(a.astype(object), b.astype(object))
    ref = full[:2048].copy(); ref[:2047] -= full[2048:]
    err = np.max(np.abs(got - ref.astype(np.float64)))
    print("max abs err", err)
    assert err < 1e-3
    # check frequency ordering claim: F[lane=k1][q=k2] = sum_j z_j w_j exp(-2 pi i j k/1024), k=k1+32k2
    z = (a[:1024] + 1j * a[1024:]) * np.exp(1j * np.pi * np.arange(1024) / 2048)
    X = np.fft.fft(z)
    F = forward(a)
    k = np.arange(1024)
    print("order err", np.max(np.abs(F[k % 32, k // 32] - X)))

"""numpy model of the four-warps-per-polynomial negacyclic FFT used by pbs_kernel_lat4 (thread (lane, r) holds 8 of the
32 points of a column; cross-warp radix-4 stage, 8-point transforms, one transposition).  Checks the forward result against
the direct definition (same frequency layout as fft.cuh) and the round trip."""
import numpy as np

N2 = 1024
w32 = np.exp(-2j * np.pi * np.arange(32) / 32)
w4 = np.array([1, -1j, -1, 1j])
rng = np.random.default_rng(0)
z = rng.integers(-2**22, 2**22, N2) + 1j * rng.integers(-2**22, 2**22, N2)
j = np.arange(N2)
zt = z * np.exp(1j * np.pi * j / 2048)
F_ref = np.array([np.sum(zt * np.exp(-2j * np.pi * j * k / N2)) for k in range(N2)])

def Tp(k1, l):
    return np.exp(-2j * np.pi * l * k1 / 1024) * np.exp(1j * np.pi * l / 2048)

def dft8(x, inv=False):
    s = 1 if inv else -1
    k = np.arange(8)
    return np.array([np.sum(x * np.exp(s * 2j * np.pi * k * kk / 8)) for kk in range(8)])

def fwd_pass(quarters, r):
    """quarters[q][mm] = x[mm + 8 q]; warp r returns outputs with index 4*kappa + r."""
    t = sum(w4[(q * r) % 4] * quarters[q] for q in range(4))
    t = t * w32[(np.arange(8) * r) % 32]
    return dft8(t)

def inv_send(vals, r):
    """vals[kappa] = input with index 4*kappa + r; returns u_r[mm] (what warp r publishes)."""
    return dft8(vals, inv=True) * np.conj(w32[(np.arange(8) * r) % 32])

def inv_combine(u, q):
    """warp q: result for index mm + 8 q."""
    return sum(np.conj(w4[(q * r) % 4]) * u[r] for r in range(4))

C = np.exp(1j * np.pi * np.arange(32) / 64)
x = np.zeros((32, 4, 8), complex)                             # [lane l][quarter q][mm] = z_{l+32m} * C_m, m = mm + 8q
for l in range(32):
    for q in range(4):
        for mm in range(8):
            m = mm + 8 * q
            x[l, q, mm] = z[l + 32 * m] * C[m]
Y = np.zeros((32, 32), complex)                               # [l][k1]
for l in range(32):
    for r in range(4):
        out = fwd_pass(x[l], r)
        for kap in range(8):
            k1 = 4 * kap + r
            Y[l, k1] = out[kap] * Tp(k1, l)
F = np.zeros(N2, complex)
for k1 in range(32):
    quarters = [Y[8 * q:8 * q + 8, k1] for q in range(4)]     # after the transposition thread (lane k1, q) holds l = ll + 8 q
    for r in range(4):
        out = fwd_pass(quarters, r)
        for kap in range(8):
            F[k1 + 32 * (4 * kap + r)] = out[kap]
print("forward max rel err", np.abs(F - F_ref).max() / np.abs(F_ref).max())

G = F
y = np.zeros((32, 32), complex)                               # [k1][l]
for k1 in range(32):
    u = [inv_send(np.array([G[k1 + 32 * (4 * kap + r)] for kap in range(8)]), r) for r in range(4)]
    for q in range(4):
        y[k1, 8 * q:8 * q + 8] = inv_combine(u, q)
xo = np.zeros(N2, complex)
for l in range(32):
    u = [inv_send(np.array([y[4 * kap + r, l] * np.conj(Tp(4 * kap + r, l)) for kap in range(8)]), r) for r in range(4)]
    for q in range(4):
        res = inv_combine(u, q)
        for mm in range(8):
            xo[l + 32 * (mm + 8 * q)] = res[mm] * np.conj(C[mm + 8 * q])
print("round trip max err", np.abs(xo / 1024 - z).max())

"""numpy model of the two-warps-per-polynomial negacyclic FFT used by pbs_kernel_lat (thread (lane, h) holds
16 of the 32 points of a column; cross-thread radix-2 stage, 16-point transforms, one transposition).
Checks the forward result against the direct definition (same frequency layout as fft.cuh) and the round trip."""
import numpy as np

N2 = 1024
w32 = np.exp(-2j * np.pi * np.arange(32) / 32)
rng = np.random.default_rng(0)
z = rng.integers(-2**22, 2**22, N2) + 1j * rng.integers(-2**22, 2**22, N2)
j = np.arange(N2)
zt = z * np.exp(1j * np.pi * j / 2048)                       # twisted input
F_ref = np.array([np.sum(zt * np.exp(-2j * np.pi * j * k / N2)) for k in range(N2)])

def Tp(k1, l):
    return np.exp(-2j * np.pi * l * k1 / 1024) * np.exp(1j * np.pi * l / 2048)

def dft16(x, inv=False):
    s = 1 if inv else -1
    k = np.arange(16)
    return np.array([np.sum(x * np.exp(s * 2j * np.pi * k * kk / 16)) for kk in range(16)])

def fwd_pass(own, recv, h):
    """own/recv: 16 values (own half of the 32 points); returns outputs index 2*kappa + h."""
    t = own + (1 if h == 0 else -1) * recv
    if h:
        t = t * (-w32[:16])
    return dft16(t)

def inv_pass(vals, h):
    """vals: 16 inputs with index 2*kappa + h; returns the value to send (E or O*w) -- combine separately."""
    e = dft16(vals, inv=True)
    if h:
        e = e * np.conj(w32[:16])
    return e

# ---- forward
C = np.exp(1j * np.pi * np.arange(32) / 64)
x = np.zeros((32, 2, 16), complex)                            # [lane l][h][mm] = z_{l+32m} * C_m, m = mm + 16h
for l in range(32):
    for h in range(2):
        for mm in range(16):
            m = mm + 16 * h
            x[l, h, mm] = z[l + 32 * m] * C[m]
Y = np.zeros((32, 32), complex)                               # [l][k1]
for l in range(32):
    for h in range(2):
        out = fwd_pass(x[l, h], x[l, 1 - h], h)
        for kap in range(16):
            k1 = 2 * kap + h
            Y[l, k1] = out[kap] * Tp(k1, l)
F = np.zeros(N2, complex)
for k1 in range(32):
    for hp in range(2):
        own = Y[16 * hp:16 * hp + 16, k1]
        recv = Y[16 * (1 - hp):16 * (1 - hp) + 16, k1]
        out = fwd_pass(own, recv, hp)
        for kap in range(16):
            F[k1 + 32 * (2 * kap + hp)] = out[kap]
print("forward max rel err", np.abs(F - F_ref).max() / np.abs(F_ref).max())

# ---- inverse (unnormalised), then untwist: must give 1024 * z
G = F
y = np.zeros((32, 32), complex)                               # [k1][l]
for k1 in range(32):
    snd = [inv_pass(np.array([G[k1 + 32 * (2 * kap + hp)] for kap in range(16)]), hp) for hp in range(2)]
    y[k1, 0:16] = snd[0] + snd[1]                             # thread 0: own + recv
    y[k1, 16:32] = snd[0] - snd[1]                            # thread 1: recv - own
xo = np.zeros(N2, complex)
for l in range(32):
    snd = [inv_pass(np.array([y[2 * kap + h, l] * np.conj(Tp(2 * kap + h, l)) for kap in range(16)]), h) for h in range(2)]
    lo, hi = snd[0] + snd[1], snd[0] - snd[1]
    for mm in range(16):
        xo[l + 32 * mm] = lo[mm] * np.conj(C[mm])
        xo[l + 32 * (mm + 16)] = hi[mm] * np.conj(C[mm + 16])
print("round trip max err", np.abs(xo / 1024 - z).max())

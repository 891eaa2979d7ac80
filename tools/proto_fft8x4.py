"""numpy model of the four-warps-per-polynomial negacyclic FFT used by pbs_lat4_kernel (pbs_kernel_lat4.cuh): thread
(lane, warp (c, e)) holds the input quarter q = c + 2 e (8 of the 32 points of a column) and produces the output residue
r = 2 c + e; a 32-point pass = cross-warp radix-2 level with distance 16 (through tensor memory in the kernel), cross-warp radix-2
level with distance 8 (through shared memory), 8-point transform in registers; one transposition between the passes.  Checks the
forward result against the direct definition (same frequency layout as fft.cuh) and the round trip.  Run by tests/test_oracle.py."""
import numpy as np

N2 = 1024
w32 = np.exp(-2j * np.pi * np.arange(32) / 32)
MM = np.arange(8)


def Tp(k1, l):
    return np.exp(-2j * np.pi * l * k1 / 1024) * np.exp(1j * np.pi * l / 2048)


def dft8(x, inv=False):
    return np.fft.ifft(x) * 8 if inv else np.fft.fft(x)


def fwd_pass(quarters):
    """quarters[q][mm] = x[mm + 8 q]; returns out[(c, e)][kappa] = output with index 4 kappa + 2 c + e."""
    lvl1 = {}
    for c in range(2):
        lvl1[(c, 0)] = quarters[c] + quarters[c + 2]                       # warp e = 0: own + other
        lvl1[(c, 1)] = (quarters[c] - quarters[c + 2]) * w32[MM + 8 * c]   # warp e = 1: (other - own) W32^(mm + 8 c)
    out = {}
    for e in range(2):
        s0, s1 = lvl1[(0, e)], lvl1[(1, e)]
        out[(0, e)] = dft8(s0 + s1)                                       # warp c = 0
        out[(1, e)] = dft8((s0 - s1) * w32[2 * MM])                       # warp c = 1
    return out


def inv_pass(vals):
    """vals[(c, e)][kappa] = input with index 4 kappa + 2 c + e; returns y[q][mm] = result for index mm + 8 q, q = c + 2 e."""
    u = {k: dft8(v, inv=True) for k, v in vals.items()}
    s = {}
    for e in range(2):
        t = np.conj(w32[2 * MM]) * u[(1, e)]                              # warp c = 1 publishes u conj(W32^(2 mm))
        s[(0, e)] = u[(0, e)] + t
        s[(1, e)] = u[(0, e)] - t
    y = [None] * 4
    for c in range(2):
        t = np.conj(w32[MM + 8 * c]) * s[(c, 1)]                          # warp e = 1 publishes s conj(W32^(mm + 8 c))
        y[c] = s[(c, 0)] + t
        y[c + 2] = s[(c, 0)] - t
    return y


def run(seed=0):
    rng = np.random.default_rng(seed)
    z = rng.integers(-2**22, 2**22, N2) + 1j * rng.integers(-2**22, 2**22, N2)
    j = np.arange(N2)
    F_ref = np.fft.fft(z * np.exp(1j * np.pi * j / 2048))
    C = np.exp(1j * np.pi * np.arange(32) / 64)
    # forward: pass 1 over m for every column l, twiddle, transposition, pass 2 over l for every k1
    Y = np.zeros((32, 32), complex)                                        # [l][k1]
    for l in range(32):
        out = fwd_pass([np.array([z[l + 32 * (mm + 8 * q)] * C[mm + 8 * q] for mm in range(8)]) for q in range(4)])
        for (c, e), v in out.items():
            for kap in range(8):
                k1 = 4 * kap + 2 * c + e
                Y[l, k1] = v[kap] * Tp(k1, l)
    F = np.zeros(N2, complex)
    for k1 in range(32):
        out = fwd_pass([Y[8 * q:8 * q + 8, k1] for q in range(4)])        # thread (lane k1, quarter q) holds l = ll + 8 q
        for (c, e), v in out.items():
            for kap in range(8):
                F[k1 + 32 * (4 * kap + 2 * c + e)] = v[kap]
    fwd_err = np.abs(F - F_ref).max() / np.abs(F_ref).max()
    # inverse (unnormalised), untwist: must give 1024 z
    y = np.zeros((32, 32), complex)                                        # [k1][l]
    for k1 in range(32):
        res = inv_pass({(c, e): np.array([F[k1 + 32 * (4 * kap + 2 * c + e)] for kap in range(8)]) for c in range(2) for e in range(2)})
        for q in range(4):
            y[k1, 8 * q:8 * q + 8] = res[q]
    xo = np.zeros(N2, complex)
    for l in range(32):
        res = inv_pass({(c, e): np.array([y[4 * kap + 2 * c + e, l] * np.conj(Tp(4 * kap + 2 * c + e, l)) for kap in range(8)])
                        for c in range(2) for e in range(2)})
        for q in range(4):
            for mm in range(8):
                xo[l + 32 * (mm + 8 * q)] = res[q][mm] * np.conj(C[mm + 8 * q])
    rt_err = np.abs(xo / 1024 - z).max()
    return fwd_err, rt_err


if __name__ == "__main__":
    fe, re_ = run()
    print("forward max rel err", fe)
    print("round trip max err", re_)

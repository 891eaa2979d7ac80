"""Development aid: one KS+PBS (and a batch) of PARAM_MESSAGE_1_CARRY_1 / 3_3 on the generic kernel with random keys (timing / ncu target).
usage: perf_generic.py [1_1|3_3] [batch]"""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import tfhe_rs_string_b200 as T
which = sys.argv[1] if len(sys.argv) > 1 else "1_1"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n, k, N, pb, pl, kb, kl, mm, cm = (684, 3, 512, 18, 1, 4, 3, 2, 2) if which == "1_1" else (864, 1, 8192, 15, 2, 3, 6, 8, 8)
eng = T.Engine(T.Params(n, k, N, pb, pl, kb, kl, mm, cm), device=0)
rng = np.random.default_rng(0)
eng.load_ksk(rng.integers(0, 2**64, k * N * kl * (n + 1), dtype=np.uint64))
eng.load_bsk_standard(rng.integers(0, 2**64, n * pl * (k + 1) * (k + 1) * N, dtype=np.uint64))
ids = np.full(batch, eng.generate_lookup_table(lambda x: x), dtype=np.uint32)
cts = rng.integers(0, 2**64, (batch, k * N + 1), dtype=np.uint64)
eng.ks_pbs_batch(cts, ids)
t0 = time.perf_counter()
for _ in range(3): eng.ks_pbs_batch(cts, ids)
print(f"{which} batch {batch}: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms per call")

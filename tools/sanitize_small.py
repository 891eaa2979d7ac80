"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): KS + PBS (1..4 ciphertexts per CTA
paths), PBS->KS is not needed here.  Random keys; only memory safety / hazards are of interest."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tfhe_rs_string_b200 as T

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24       # CMUX steps (lwe_dimension) kept tiny for the tool
p = T.Params(n, 1, 2048, 23, 1, 3, 5, 4, 4)
eng = T.Engine(p, 0)
rng = np.random.default_rng(1)
eng.load_ksk(rng.integers(0, 2**64, 2048 * 5 * (n + 1), dtype=np.uint64))
eng.load_bsk_standard(rng.integers(0, 2**64, n * 4 * 2048, dtype=np.uint64))
lid = eng.generate_lookup_table(lambda x: x)
for batch in (1, 3, 150, 300, 450, 600):
    cts = rng.integers(0, 2**64, (batch, 2049), dtype=np.uint64)
    out = eng.ks_pbs_batch(cts, np.full(batch, lid, dtype=np.uint32))
    assert out.shape == (batch, 2049)
eng.close()
print("sanitize_small ok")

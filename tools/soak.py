"""Soak / determinism check: the same inputs must give bit-identical outputs on every repetition for every
launch configuration (a race in the kernels' shared-memory / TMEM hand-overs would show up as a diff or a hang)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tfhe_rs_string_b200 as T

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
p = T.Params.message_2_carry_2()
eng = T.Engine(p, 0)
rng = np.random.default_rng(3)
eng.load_ksk(rng.integers(0, 2**64, 2048 * 5 * 743, dtype=np.uint64))
eng.load_bsk_standard(rng.integers(0, 2**64, 742 * 4 * 2048, dtype=np.uint64))
lids = [eng.generate_lookup_table(lambda x, k=k: (x * k + 1) % 16) for k in range(1, 6)]
t0 = time.time()
for batch in (1, 37, 148, 149, 296, 297, 444, 592, 593, 1000, 4096):
    d_in = torch.from_numpy(rng.integers(-2**63, 2**63, (batch, 2049), dtype=np.int64)).cuda()
    d_ids = torch.from_numpy(np.array([lids[i % 5] for i in range(batch)], dtype=np.int32)).cuda()
    d_out = torch.empty_like(d_in)
    ref = None
    for r in range(reps if batch < 4096 else max(3, reps // 4)):
        d_out.zero_()
        eng.ks_pbs_batch_device(d_in, d_ids, d_out, batch)
        eng.sync()
        h = d_out.cpu()
        if ref is None:
            ref = h
        else:
            assert torch.equal(ref, h), f"non-deterministic output at batch {batch}, repetition {r}"
    print(f"batch {batch}: {r + 1} repetitions identical", flush=True)
print(f"soak ok in {time.time() - t0:.1f} s")

"""Quick device-resident timing of the KS and PBS kernels (development aid; bench.py is the contract)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tfhe_rs_string_b200 as T

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p = T.Params.message_2_carry_2()
eng = T.Engine(p, 0)
rng = np.random.default_rng(0)
# random keys are fine for timing: the kernels are data-oblivious except the (measure-zero) a~ == 0 skip
eng.load_ksk(rng.integers(0, 2**64, 2048 * 5 * 743, dtype=np.uint64))
eng.load_bsk_standard(rng.integers(0, 2**64, 742 * 4 * 2048, dtype=np.uint64))
lid = eng.generate_lookup_table(lambda x: x)
d_in = torch.from_numpy(rng.integers(0, 2**63, (B, 2049), dtype=np.int64)).cuda()
d_out = torch.empty_like(d_in)
d_ids = torch.full((B,), lid, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
eng.set_profiling(True)
for v in [3]:
    for _ in range(2):
        eng.ks_pbs_batch_device(d_in, d_ids, d_out, B)
    eng.sync(); eng.kernel_times(reset=True)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.ks_pbs_batch_device(d_in, d_ids, d_out, B)
    eng.sync()
    wall = (time.perf_counter() - t0) / reps
    kt = eng.kernel_times(reset=True)
    pbs_ms = kt["pbs_ms"] / kt["pbs_launches"]; ks_ms = kt["ks_ms"] / kt["ks_launches"]
    print(json.dumps({"variant": v, "batch": B, "wall_ms": wall * 1e3, "ks_ms": ks_ms, "pbs_ms": pbs_ms,
                      "ks_pbs_per_s": B / wall, "pbs_only_per_s": B / (pbs_ms * 1e-3),
                      "fp64_tflops_algorithmic": B * 1.945e8 / (pbs_ms * 1e-3) / 1e12}))

"""Device-resident latency of the batched call sites at the BASELINE.json shapes (configs 2-5).
Development aid; random keys / inputs (timing is data-oblivious)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tfhe_rs_string_b200 as T

p = T.Params.message_2_carry_2()
eng = T.Engine(p, 0)
rng = np.random.default_rng(0)
eng.load_ksk(rng.integers(0, 2**64, 2048 * 5 * 743, dtype=np.uint64))
eng.load_bsk_standard(rng.integers(0, 2**64, 742 * 4 * 2048, dtype=np.uint64))
CASES = [
    ("radix_eq", [1024, 4], "config 2: 1024 FheUint8 pairs, eq"),
    ("radix_add", [1024, 4], "config 2: 1024 FheUint8 pairs, add"),
    ("string_eq", [256, 64, 64, 4], "config 3: 256 pairs of 64-char strings, eq"),
    ("string_eq", [1, 64, 64, 4], "config 3: ONE pair of 64-char strings, eq (latency)"),
    ("string_to_uppercase", [256, 64, 4], "config 3: 256 strings x 64 chars, to_uppercase"),
    ("string_contains", [1, 256, 8, 4], "config 4: contains, 8-char pattern in 256 chars"),
    ("string_find", [1, 256, 8, 4], "config 4: find, 8-char pattern in 256 chars"),
    ("trivium", [16, 0, 0], "config 5: trivium, 1152 warm-up + 1024 output bits"),
]
sel = sys.argv[1].split(",") if len(sys.argv) > 1 else None
for op, shape, label in CASES:
    if sel and op not in sel:
        continue
    prog = T.Program(eng, op, shape)
    n_in, n_out = prog.info["n_inputs"], prog.info["n_outputs"]
    d_in = torch.from_numpy(rng.integers(0, 2**63, (n_in, 2049), dtype=np.int64)).cuda()
    d_out = torch.empty((n_out, 2049), dtype=torch.int64, device="cuda")
    prog.run_device(d_in, d_out); eng.sync()
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        prog.run_device(d_in, d_out)
    eng.sync()
    ms = (time.perf_counter() - t0) / reps * 1e3
    print(json.dumps({"op": op, "shape": shape, "label": label, "ms": round(ms, 3), "n_pbs": prog.info["n_pbs"],
                      "depth": prog.info["depth"], "pbs_per_s": round(prog.info["n_pbs"] / ms * 1e3)}))
    prog.close()

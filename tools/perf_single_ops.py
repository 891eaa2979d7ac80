import json, os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import tfhe_rs_string_b200 as T
p = T.Params.message_2_carry_2(); eng = T.Engine(p, 0); rng = np.random.default_rng(0)
eng.load_ksk(rng.integers(0, 2**64, 2048 * 5 * 743, dtype=np.uint64)); eng.load_bsk_standard(rng.integers(0, 2**64, 742 * 4 * 2048, dtype=np.uint64))
for op, shape in [("radix_eq",[1,4]),("radix_ne",[1,4]),("radix_add",[1,4]),("radix_sub",[1,4]),("radix_scalar_gt",[1,4,100]),("radix_bitand",[1,4]),("string_find",[1,3,2,4]),("string_eq",[1,64,64,4]),("string_to_uppercase",[1,64,4]),("string_contains",[1,256,8,4])]:
    prog = T.Program(eng, op, shape)
    d_in = torch.from_numpy(rng.integers(0, 2**63, (prog.info["n_inputs"], 2049), dtype=np.int64)).cuda()
    d_out = torch.empty((prog.info["n_outputs"], 2049), dtype=torch.int64, device="cuda")
    prog.run_device(d_in, d_out); eng.sync()
    t0 = time.perf_counter()
    for _ in range(5): prog.run_device(d_in, d_out)
    eng.sync()
    print(json.dumps({"op": op, "shape": shape, "ms": round((time.perf_counter()-t0)/5*1e3, 2), "n_pbs": prog.info["n_pbs"], "depth": prog.info["depth"]}))

"""Device-resident timing of the keyswitch kernel alone (development aid)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tfhe_rs_string_b200 as T
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = T.Engine(T.Params.message_2_carry_2(), 0)
rng = np.random.default_rng(0)
eng.load_ksk(rng.integers(0, 2**64, 2048 * 5 * 743, dtype=np.uint64))
d_in = torch.from_numpy(rng.integers(-2**63, 2**63, (B, 2049), dtype=np.int64)).cuda()
d_out = torch.empty((B, 743), dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
eng.set_profiling(True)
for _ in range(2): eng.keyswitch_batch_device(d_in, d_out, B)
eng.sync(); eng.kernel_times(reset=True)
for _ in range(5): eng.keyswitch_batch_device(d_in, d_out, B)
kt = eng.kernel_times(reset=True)
ms = kt["ks_ms"] / kt["ks_launches"]
print(json.dumps({"ks_variant": os.environ.get("B200TFHE_KS_VARIANT", "0"), "batch": B, "ks_ms": ms, "ks_per_s": B / ms * 1e3,
                  "u64_mac_per_s": B * 2048 * 5 * 743 / ms * 1e3, "ksk_stream_GBps_algorithmic": (B / 64) * 60866560 / ms / 1e6}))

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def toy_keys(oracle_mod):
    """Insecure fast parameters (examples/fhe_strings/ciphertext.rs:76-90 shape, n raised to 16)."""
    return oracle_mod.Keyset(oracle_mod.params_toy(16, 256), seed=7)


@pytest.fixture(scope="session")
def real_keys(oracle_mod):
    """PARAM_MESSAGE_2_CARRY_2_KS_PBS keys from the oracle keygen (seed 0xB200)."""
    return oracle_mod.Keyset(oracle_mod.params_message_2_carry_2(), seed=0xB200)


@pytest.fixture(scope="session")
def engine(real_keys):
    """GPU engine with the oracle's keys uploaded through the C ABI."""
    import tfhe_rs_string_b200 as T
    p = real_keys.params
    params = T.Params(p.lwe_dimension, p.glwe_dimension, p.polynomial_size, p.pbs_base_log, p.pbs_level,
                      p.ks_base_log, p.ks_level, p.message_modulus, p.carry_modulus)
    e = T.Engine(params, device=0)
    e.load_ksk(real_keys.ksk)
    e.load_bsk_standard(real_keys.bsk_standard)
    yield e
    e.close()


def centered(x):
    """u64 array -> signed distance to 0 on the torus."""
    return np.asarray(x, dtype=np.uint64).astype(np.int64)

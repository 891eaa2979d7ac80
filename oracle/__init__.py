"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see tfhe_oracle.h).  Never imported by the
product package `tfhe_rs_string_b200`."""

#!/usr/bin/env python3
"""Extracts the Trivium known-answer vectors the reference's own tests pin
(/root/reference/apps/trivium/src/trivium/test.rs: trivium_test_1..4, first 64 keystream bytes each; values
from the ECRYPT / avr-crypto-lib trivium-80.80 test vectors) into tests/golden/trivium_kat.json.
Run in the build container (the reference tree is not present on the GPU box)."""
import json
import re

src = open("/root/reference/apps/trivium/src/trivium/test.rs").read()
kats = []
for name, body in re.findall(r"fn (trivium_test_[1-4])\(\) \{(.*?)\n\}\n", src, re.S):
    out = re.search(r'output_0_63\s*=\s*"([0-9A-F]+)"', body).group(1)
    key = [0] * 80
    iv = [0] * 80
    for var, arr in (("key", key), ("iv", iv)):
        m = re.search(var + r'_string = "([0-9A-F]+)"', body)
        if m:   # hex string, bytes in order, bits LSB first (test.rs:155-171)
            hx = m.group(1)
            for i in range(0, len(hx), 2):
                val = int(hx[i:i + 2], 16)
                for j in range(8):
                    arr[8 * (i >> 1) + j] = (val >> j) & 1
        for idx in re.findall(var + r"\[(\d+)\] = true", body):
            arr[int(idx)] = 1
    kats.append({"name": name, "key_bits": key, "iv_bits": iv, "keystream_bytes_0_63_hex": out})
json.dump({"source": "apps/trivium/src/trivium/test.rs (reference), output_0_63 of trivium_test_1..4",
           "bit_order": "keystream bit t is bit (t % 8) of byte t // 8 (LSB first), test.rs:10-58", "kats": kats},
          open("tests/golden/trivium_kat.json", "w"), indent=1)
print(len(kats), "vectors")

#!/usr/bin/env python
"""Extracts the scalar constants (OpConstant of 32-bit float / int types) from the reference's precompiled SPIR-V modules of
the DDGI path and writes them to tests/golden/spirv_constants.json.  Run in the build container, where /root/reference is
mounted; the JSON is the committed fixture (tests/test_golden_constants.py never reads /root/reference).

    python tests/golden/make_spirv_constants.py [/root/reference/assets]
"""
import json
import os
import struct
import sys

MODULES = ["irradiance_probe_update_rgen.spv", "irradiance_probe_update_rchit.spv", "irradiance_probe_update_rmiss.spv",
           "probe_conv_comp.spv", "shadow_rgen.spv", "shadow_rmiss.spv", "mirror_reflections_rgen.spv",
           "mirror_reflections_rchit.spv", "mirror_reflections_rmiss.spv"]


def constants(path):
    data = open(path, "rb").read()
    words = struct.unpack("<%dI" % (len(data) // 4), data)
    assert words[0] == 0x07230203, "not a SPIR-V module"
    floats, ints, ftypes, itypes = set(), set(), set(), {}
    i = 5
    while i < len(words):
        op, wc = words[i] & 0xFFFF, words[i] >> 16
        if op == 22 and words[i + 2] == 32:                    # OpTypeFloat <id> 32
            ftypes.add(words[i + 1])
        elif op == 21 and words[i + 2] == 32:                  # OpTypeInt <id> 32 signedness
            itypes[words[i + 1]] = words[i + 3]
        elif op == 43:                                         # OpConstant <type> <id> <value>
            ty, val = words[i + 1], words[i + 3]
            if ty in ftypes:
                floats.add(struct.unpack("<f", struct.pack("<I", val))[0])
            elif ty in itypes:
                ints.add(struct.unpack("<i", struct.pack("<I", val))[0] if itypes[ty] else val)
        i += wc
    return sorted(floats), sorted(ints)


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets"
    out = {}
    for m in MODULES:
        p = os.path.join(root, m)
        if os.path.exists(p):
            f, n = constants(p)
            out[m] = {"float32": [repr(float(x)) for x in f], "int32": n}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spirv_constants.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print("wrote", dst, {k: (len(v["float32"]), len(v["int32"])) for k, v in out.items()})


if __name__ == "__main__":
    main()

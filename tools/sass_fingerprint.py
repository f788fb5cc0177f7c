#!/usr/bin/env python
"""Fingerprint of the DEVICE code in libjpetra_b200.so's object files: one SHA-256 per kernel over its SASS instruction
stream (cuobjdump -sass; addresses and encodings stripped, anonymous-namespace hashes normalised).  Used to show that
a change touched host code only, i.e. that the kernels are byte-for-byte the ones the last GPU run validated:
    python tools/sass_fingerprint.py > /tmp/now.txt && diff profiles/r1_sass_fingerprint_validated.txt /tmp/now.txt
"""
import glob
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "juliapetra.jl_b200", "csrc", "_build", "*.o")))
    if not objs:
        sys.exit("build the library first (make -C juliapetra.jl_b200/csrc)")
    for obj in objs:
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
        name, h, kernels = None, None, {}
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name = re.sub(r"_GLOBAL__N__[0-9a-f]+_", "_GLOBAL__N__H_", m.group(1))
                kernels[name] = [hashlib.sha256(), 0]
                continue
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", line)
            if m and name:
                kernels[name][0].update(m.group(1).encode())
                kernels[name][1] += 1
        for k in sorted(kernels):
            dem = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
            dem = re.sub(r"\(anonymous namespace\)::", "", dem).split("(")[0]
            print(f"{os.path.basename(obj):<14} {kernels[k][0].hexdigest()[:16]} {kernels[k][1]:>6} instr  {dem}")


if __name__ == "__main__":
    main()

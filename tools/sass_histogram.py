#!/usr/bin/env python
"""SASS mnemonic histogram of selected kernels of lib/libsxgpu.so (cuobjdump -sass; no GPU needed).

    python tools/sass_histogram.py bank_repeat batch_warp bank_tx >> profiles/r01_sass_mnemonics.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "sxxcvr_b200" / "lib" / "libsxgpu.so"


def main():
    want = sys.argv[1:] or ["bulk_convert_kernel"]
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    for body in re.split(r"\n\s*Function : ", sass)[1:]:
        mangled = body.split("\n", 1)[0].strip()
        if not any(w in mangled for w in want):
            continue
        name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip() or mangled
        ops = collections.Counter()
        for line in body.split("\n"):
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m:
                ops[m.group(1)] += 1
        print(f"== {name} (sm_100a SASS mnemonic counts, cuobjdump -sass sxxcvr_b200/lib/libsxgpu.so)")
        for op, n in sorted(ops.items(), key=lambda kv: (-kv[1], kv[0])):
            print(f"{n:7d} {op}")
        print()


if __name__ == "__main__":
    main()

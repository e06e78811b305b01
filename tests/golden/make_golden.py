"""Regenerates tests/golden/ref_shfun.txt from the reference's OWN header (src/sh/SH_function.h), compiled in
place by `make -C oracle ref` into oracle/_ref/ref_shfun.  Run in the build container (needs /root/reference)."""
import os
import subprocess

here = os.path.dirname(os.path.abspath(__file__))
root = os.path.dirname(os.path.dirname(here))
subprocess.check_call(["make", "-s", "-C", os.path.join(root, "oracle"), "ref"])
out = subprocess.check_output([os.path.join(root, "oracle", "_ref", "ref_shfun"), "golden"], text=True)
with open(os.path.join(here, "ref_shfun.txt"), "w") as f:
    f.write(out)
print(f"wrote {len(out.splitlines())} lines")

"""Builds a variant of the library next to the shipped one (A/B runs, profile builds):
    python tools/build_variant.py NAME [-DMACRO ...]   ->  m3l_b200/lib/variant_NAME.so
Use with M3L_B200_LIB=$PWD/m3l_b200/lib/variant_NAME.so."""
import subprocess, sys
from pathlib import Path
sys.path.insert(0, ".")
from m3l_b200 import build as B
name, defs = sys.argv[1], sys.argv[2:]
out = B.OBJ / f"variant_{name}"
out.mkdir(parents=True, exist_ok=True)
objs = []
for src in B._sources():
    o = out / (src.stem + ".o")
    cmd = [B.NVCC, *B.ARCH_FLAGS, *B.flags_for(src), *defs, "-c", str(src), "-o", str(o)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.exit(r.stderr)
    objs.append(str(o))
lib = B.LIBDIR / f"variant_{name}.so"
subprocess.run([B.NVCC, *B.ARCH_FLAGS, "-shared", "-o", str(lib), *objs, "-lcudart"], check=True)
print(lib)

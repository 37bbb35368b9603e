"""Builds libdregb200 variants with extra -D flags into gpurun_out-free scratch paths for A/B runs:
   python scripts/build_variant.py NAME -DFOO=1 -DBAR=2   ->  variants/libdregb200_NAME.so
Select at run time with DRB_LIB_PATH=variants/libdregb200_NAME.so (experiments only)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib
b = importlib.import_module("dreg-nerf_b200.build")
name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "variants")
os.makedirs(os.path.join(out, "obj_" + name), exist_ok=True)
procs = []
for src in b.SOURCES:
    obj = os.path.join(out, "obj_" + name, src.replace(".cu", ".o"))
    src_path = os.path.join(b.CSRC, src)
    ref_obj = os.path.join(b.HERE, "build", src.replace(".cu", ".o"))
    if src != "ngp.cu" and os.path.exists(ref_obj):       # flags only touch ngp.cu in these experiments
        procs.append((obj, None, ref_obj)); continue
    procs.append((obj, subprocess.Popen([b._nvcc()] + b.NVCC_FLAGS + flags + ["-c", src_path, "-o", obj]), None))
objs = []
for obj, p, ref in procs:
    if p is None: objs.append(ref); continue
    assert p.wait() == 0
    objs.append(obj)
lib = os.path.join(out, "libdregb200_%s.so" % name)
subprocess.run([b._nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"], check=True)
print(lib)

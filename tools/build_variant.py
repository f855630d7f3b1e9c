"""Build an A/B variant of the library: one csrc file recompiled with extra -D flags, linked with the
shipped objects of the others -> icp_slam-yolo_b200/lib/variants/libb200icp_<name>.so (git-ignored, travels
with gpurun).  Run a tool against it with B200ICP_LIB=<path>.

  python tools/build_variant.py s2m_prefetch scan2map.cu -DS2M_PREFETCH=1
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("b200build", os.path.join(ROOT, "icp_slam-yolo_b200", "build.py"))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)


def main():
    name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
    b.build()
    vdir = os.path.join(b.LIB_DIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    obj = os.path.join(vdir, f"{name}_{src[:-3]}.o")
    subprocess.run(["nvcc", *b.NVCC_FLAGS, *flags, "-I", b.INCLUDE, "-I", b.CSRC, "-c", "-o", obj,
                    os.path.join(b.CSRC, src)], check=True)
    objs = [obj if os.path.basename(s) == src else b._obj(s) for s in b.sources()]
    out = os.path.join(vdir, f"libb200icp_{name}.so")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out, *objs], check=True)
    os.remove(obj)
    print(out)


if __name__ == "__main__":
    main()

"""Builds experiment variants of libvqb200.so:  python tools/ab_build.py name1:-DX=1,-DY=2 name2: ...
-> vq-vae-transformer-arc-welding_b200/ab_<name>.so (git-ignored, travels to the GPU box)."""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vq-vae-transformer-arc-welding_b200")
spec = importlib.util.spec_from_file_location("b", os.path.join(PKG, "csrc", "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
for arg in sys.argv[1:]:
    name, _, flags = arg.partition(":")
    out = os.path.join(PKG, f"ab_{name}.so")
    b.build(force=True, out=out, extra_flags=[f for f in flags.split(",") if f])
    print("built", out)

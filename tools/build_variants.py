"""Build side-by-side library variants for A/B timing: tools/build_variants.py NAME=-DFLAG ..."""
import importlib.util
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("pp_build", os.path.join(ROOT, "pumi-pic_b200", "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
for arg in sys.argv[1:]:
    name, _, flags = arg.partition("=")
    out = os.path.join(ROOT, "pumi-pic_b200", "_variants", "lib_%s.so" % name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    b.build(force=True, extra_flags=[f for f in flags.split(",") if f],
            lib=out, obj=os.path.join(ROOT, "pumi-pic_b200", "_obj_" + name))
    print(out)

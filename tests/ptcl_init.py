"""Re-export of the synthetic workload generators (pumi-pic_b200/workloads.py) for the tests."""
import importlib.util
import os

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pumi-pic_b200",
                  "workloads.py")
_spec = importlib.util.spec_from_file_location("pp_workloads", _p)
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})

"""CPU: every script under tools/ and the repo-root entry points compile (the GPU box is the wrong place to
find a syntax error), and the shell scripts only call files that exist."""
import glob
import os
import py_compile
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_python_sources_compile(tmp_path):
    files = (glob.glob(os.path.join(ROOT, "tools", "*.py")) + glob.glob(os.path.join(ROOT, "pumi-pic_b200", "*.py"))
             + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")])
    assert len(files) > 10
    for f in files:
        py_compile.compile(f, cfile=str(tmp_path / (os.path.basename(f) + "c")), doraise=True)


def test_gpu_scripts_reference_existing_files():
    for sh in glob.glob(os.path.join(ROOT, "tools", "gpu_*.sh")):
        text = open(sh).read()
        for rel in set(re.findall(r"\b(tools/[\w./-]+\.py|tests/[\w./-]+\.py|bench\.py)\b", text)):
            assert os.path.exists(os.path.join(ROOT, rel)), (os.path.basename(sh), rel)

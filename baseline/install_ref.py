"""Copy the UNMODIFIED reference sources into baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU
box, where /root/reference does not exist). The reference is a script collection without setup.py / pyproject.toml, so
the contract's `pip install --target baseline/_ref /root/reference` has nothing to build; a plain copy of its package
directories and top-level modules is the install. Run in the build container (also called by __graft_entry__.build()):

    python baseline/install_ref.py
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
WANT = ("generators", "discriminators", "datasets", "util.py", "train.py", "test.py", "two_step_test.py",
        "requirements.txt")


def install(src="/root/reference"):
    """-> True if baseline/_ref is populated (copied now or already there and identical)."""
    if not os.path.isdir(src):
        return os.path.isdir(os.path.join(DST, "generators"))
    os.makedirs(DST, exist_ok=True)
    for name in WANT:
        s, d = os.path.join(src, name), os.path.join(DST, name)
        if not os.path.exists(s):
            continue
        if os.path.isdir(s):
            if os.path.isdir(d):
                cmp = filecmp.dircmp(s, d, ignore=["__pycache__"])
                if not (cmp.left_only or cmp.diff_files):
                    continue
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__"))
        elif not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copy2(s, d)
    return True


if __name__ == "__main__":
    ok = install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("baseline/_ref", "ready" if ok else "NOT available (no /root/reference here)")

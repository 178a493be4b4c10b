"""baseline/_ref: an install of the UNMODIFIED Python reference (git-ignored, shipped to the GPU box).  Run by
__graft_entry__.build() in the build container, the only place where the reference tree exists."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def install_reference(src="/root/reference"):
    """Install the reference into baseline/_ref for bench.py's cpu_baseline.python_reference (oracle/pyref_timing.py times it
    on the box's host cores).  Where the reference tree is absent, whatever is already there is used."""
    import shutil
    import subprocess
    import tempfile
    dst = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(src, "zombsole")):
        return os.path.isdir(os.path.join(dst, "zombsole"))
    if os.path.isfile(os.path.join(dst, "zombsole", "gym_env.py")):
        return True
    os.makedirs(dst, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:  # the build writes into the source tree: install from a copy
        copy = os.path.join(tmp, "reference")
        shutil.copytree(src, copy, ignore=shutil.ignore_patterns(".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", dst, copy]
        rc = subprocess.call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if rc != 0 or not os.path.isfile(os.path.join(dst, "zombsole", "gym_env.py")):
        # pip could not build it: the package is pure Python, its directory is the install
        shutil.copytree(os.path.join(src, "zombsole"), os.path.join(dst, "zombsole"), dirs_exist_ok=True)
    return True



if __name__ == "__main__":
    print(install_reference())

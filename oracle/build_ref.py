"""
TEST INFRASTRUCTURE ONLY -- installs the UNMODIFIED reference (menickname/waafle, pure Python) into oracle/_ref/.

    python oracle/build_ref.py

Runs `pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>` (the source
tree is read-only, so the build happens in a copy under /tmp).  oracle/_ref/ is git-ignored -- no reference source enters
the history -- but it travels to the GPU box with the snapshot, so that `bench.py --impl reference` and the `cpu_baseline`
leg time the real `waafle_orgscorer` there (oracle/reference_harness.py drives it) instead of the numpy port.
Called by `__graft_entry__.build()` when /root/reference is present; a no-op otherwise.
"""

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("WAAFLE_REFERENCE_SRC", "/root/reference")


def installed():
    return os.path.isfile(os.path.join(TARGET, "waafle", "waafle_orgscorer.py"))


def build(force=False):
    """Returns 'installed', 'present' or 'no-source'."""
    if installed() and not force:
        return "present"
    if not os.path.isdir(os.path.join(SOURCE, "waafle")):
        return "no-source"
    tmp = tempfile.mkdtemp(prefix="waafle_ref_src_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns("demo", "website", ".git"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        subprocess.run([sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation",
                        "--no-deps", "--target", TARGET, src], check=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if not installed():
        raise RuntimeError("pip install of the reference did not produce oracle/_ref/waafle")
    return "installed"


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

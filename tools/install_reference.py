"""Install the UNMODIFIED reference package under ``baseline/_ref`` (git-ignored; it travels to the GPU box with the
snapshot) so that ``bench.py --impl reference`` and the same-GPU comparators can run the reference's own modules.

``python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref
<copy of /root/reference>`` fails in this image: the project's build backend (hatchling, pyproject.toml) is not
installed and there is no network.  The package is pure Python, so what pip would have put into the target directory is
exactly the ``diffsptk/`` package directory: this script copies it (byte for byte, nothing else) and records the source
commit.  Build container only -- ``/root/reference`` does not exist on the GPU box.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("DIFFSPTK_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def install(quiet=False):
    src = os.path.join(SRC, "diffsptk")
    if not os.path.isdir(src):
        return False
    dst = os.path.join(DST, "diffsptk")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DST, "INSTALLED_FROM.txt"), "w") as f:
        f.write(f"copied from {src} (pure-Python package; pip --target needs hatchling, absent offline)\n")
    if not quiet:
        print("installed", dst)
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)

"""Install the UNMODIFIED reference package into baseline/_ref (git-ignored; it travels to the GPU box with the snapshot).

    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference

fails in this image: the reference's build backend (poetry-core) is neither installed nor in /opt/wheelhouse.  The package is
pure Python, so the install is repeated from a copy under /tmp whose pyproject.toml gets a setuptools [build-system] /
[project] table instead of the poetry one -- packaging metadata only; every file under tetris_gymnasium/ is installed byte
for byte (checked below).  Dependencies are not installed (--no-deps): numpy and cv2 are in the image, gymnasium is not (the
reference arm uses the stand-in under oracle/gymnasium_shim), jax / chex are absent (the functional env cannot be imported).
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("TETRIS_REFERENCE", "/root/reference")
DEST = os.path.join(ROOT, "baseline", "_ref")

PYPROJECT = """[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"

[project]
name = "tetris-gymnasium"
version = "0.3.1"
description = "reference install for the bench's reference arm (metadata rewritten: poetry-core unavailable offline)"

[tool.setuptools.packages.find]
include = ["tetris_gymnasium*"]
"""


def install(force=False):
    if not os.path.isdir(os.path.join(REF, "tetris_gymnasium")):
        return None
    if os.path.isdir(os.path.join(DEST, "tetris_gymnasium")) and not force:
        return DEST
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "ref")
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns(".git", "docs", "examples", "tests"))
        with open(os.path.join(src, "pyproject.toml"), "w") as f:
            f.write(PYPROJECT)
        if os.path.isdir(DEST):
            shutil.rmtree(DEST)
        os.makedirs(DEST, exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", DEST, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + res.stdout[-2000:] + res.stderr[-2000:])
    # every installed source file is identical to the reference's
    cmp = filecmp.dircmp(os.path.join(REF, "tetris_gymnasium"), os.path.join(DEST, "tetris_gymnasium"), ignore=["__pycache__"])

    def walk(c):
        assert not c.diff_files and not c.left_only, (c.left, c.diff_files, c.left_only)
        for sub in c.subdirs.values():
            walk(sub)

    walk(cmp)
    return DEST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))

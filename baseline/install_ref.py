"""Install the UNMODIFIED reference (yhenon/pixeltable-yolox, /root/reference) into baseline/_ref/ for the
`bench.py --impl reference` arm and the same-box torch-GPU leg.

    python baseline/install_ref.py            (build container; /root/reference does not exist on the GPU box,
                                               baseline/_ref/ is git-ignored but travels with the gpurun snapshot)

1. The contract's recipe first: pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse
   --target baseline/_ref <copy of /root/reference>. In this image it fails: the build backend is poetry-core
   (pyproject.toml:1-3) and neither `poetry` nor `poetry-core` is installed or in the wheelhouse.
2. Fallback = what that wheel would contain: the pure-Python package directory `yolox/` copied verbatim (no file of it
   is edited) plus a `pixeltable_yolox-0.4.1.dist-info/METADATA`, so that `importlib.metadata.version("pixeltable-yolox")`
   (yolox/__init__.py:1-2) resolves without a monkeypatch.
The one import the reference makes that this image cannot satisfy, `pycocotools` (yolox/data/datasets/coco.py:7, only used
by the COCO dataset / evaluator, never by the hot path), is stubbed by the caller (bench.py: reference_modules())."""
from __future__ import annotations

import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")
DST = HERE / "_ref"


def main() -> int:
    if not REF.exists():
        print(f"{REF} not present (GPU box?): nothing to install; using the existing {DST}" if DST.exists()
              else f"{REF} not present and {DST} missing: the reference arm will report unavailable")
        return 0
    if DST.exists():
        shutil.rmtree(DST)
    with tempfile.TemporaryDirectory() as tmp:
        src = Path(tmp) / "reference"
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns("assets", "datasets", "docs", ".git"))
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                            "--find-links", "/opt/wheelhouse", "--target", str(DST), str(src)], capture_output=True, text=True)
    if r.returncode == 0 and (DST / "yolox").exists():
        print(f"pip installed the reference into {DST}")
        return 0
    why = (r.stderr.strip().splitlines() or ["?"])[-1]
    print(f"pip install failed ({why}); copying the package directory instead")
    DST.mkdir(parents=True, exist_ok=True)
    shutil.copytree(REF / "yolox", DST / "yolox", ignore=shutil.ignore_patterns("__pycache__"))
    info = DST / "pixeltable_yolox-0.4.1.dist-info"
    info.mkdir()
    (info / "METADATA").write_text("Metadata-Version: 2.1\nName: pixeltable-yolox\nVersion: 0.4.1\n")
    (info / "INSTALLER").write_text("baseline/install_ref.py (verbatim copy of /root/reference/yolox)\n")
    (info / "RECORD").write_text("")
    print(f"copied {REF / 'yolox'} -> {DST / 'yolox'}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())

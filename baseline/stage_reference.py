"""Stage the files of the reference that the drop-in tests and ``bench.py --impl reference`` execute into
``baseline/_ref/`` (git-ignored, NOT gpurun-ignored: it travels to the GPU box like a built .so, where
``/root/reference`` does not exist).  Nothing under ``baseline/_ref`` is product code or is imported by the
product; reference sources never enter the repository's history.

    python baseline/stage_reference.py            # in the build container (reads /root/reference, read-only)

Staged: ``models/`` (the quantizer and VQVAE, models/vq_vae.py), the two caller scripts whose helpers the
drop-in test runs (scripts/extract_code_indices.py, scripts/decode_with_vqvae.py) and the stage-2 YAML.
"""
import os
import shutil
import sys

SRC = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/base.py", "models/types_.py", "models/vq_vae.py",
         "scripts/extract_code_indices.py", "scripts/decode_with_vqvae.py", "configs/stage2_vq.yaml"]


def stage(src: str = SRC, dst: str = DST) -> bool:
    if not os.path.isdir(src):
        return False
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dst, rel)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged into", DST if ok else "(nothing: /root/reference is absent)")
    sys.exit(0)

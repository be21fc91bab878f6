"""Stage the UNMODIFIED reference implementation of the hot path where the GPU box can see it: oracle/_ref/.

    python oracle/build_ref.py        (also run by __graft_entry__.build() whenever /root/reference is present)

/root/reference does not exist on the GPU box, and the reference is not pip-installable (no setup.py / pyproject.toml), so the
seven pure-Python modules that `PrefixedIterDecoder` / `EmbeddingNoise` / `infer.GenerationTask` import are copied byte for byte
from where they lie into oracle/_ref/ - a build output, git-ignored (never part of the repository's history) but not
gpurun-ignored, exactly like the compiled libnovic_b200.so.  A MANIFEST records every file's source path and SHA-256.
With it present,
  - `bench.py --impl reference` and the `cpu_baseline` leg time the reference's own `PrefixedIterDecoder.generate`
    (kind "reference") instead of the oracle port,
  - the `-m gpu` tests run the reference's own `infer.GenerationTask.generate` on the registered CUDA class.
TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under novic_b200/ imports it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("NOVIC_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
# embedding_decoder -> utils, logger; infer -> embedders, embedding_dataset, embedding_noise (import closure, probed in this container)
FILES = ("embedding_decoder.py", "embedding_noise.py", "embedders.py", "embedding_dataset.py", "infer.py", "utils.py", "logger.py")


def build(verbose: bool = True) -> bool:
    if not os.path.isfile(os.path.join(SRC, "embedding_decoder.py")):
        if verbose:
            print(f"reference tree not found at {SRC}: oracle/_ref left as it is")
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for name in FILES:
        src = os.path.join(SRC, name)
        shutil.copyfile(src, os.path.join(DST, name))
        manifest[name] = {"source": src, "sha256": hashlib.sha256(open(src, "rb").read()).hexdigest()}
    json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"staged {len(FILES)} unmodified reference modules into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)

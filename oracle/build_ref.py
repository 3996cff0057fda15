"""Recipe for oracle/_ref: the reference's OWN implementation of the hot path, staged for the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY (the product never imports it).  The reference is pure Python; the GPU box has no
/root/reference, so the files SURVEY.md section 8(a) cites are copied -- unmodified, byte for byte -- from where they lie
under /root/reference into oracle/_ref/ (git-ignored: reference sources never enter this repository's history; not
gpurun-ignored: the directory travels to the box like the built libmindrec.so).  `bench.py --impl reference` and the
cpu_baseline leg then time the real models.TwoTower + torch.optim.Adam through the Manager._train loop (oracle/ref_harness.py).

    python oracle/build_ref.py            # or __graft_entry__.build(), which calls build() when /root/reference exists
"""
from __future__ import annotations

import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("MIND_REFERENCE", "/root/reference")

# the hot path: model, encoders, attention, embedding; utils/* for cal_metric / _group_lists / Partition_Sampler
FILES = [
    "models/TwoTower.py", "models/TwoTowerBaseModel.py",
    "models/Embeddings/BERT.py",
    "models/Encoders/CNN.py", "models/Encoders/RNN.py", "models/Encoders/MHA.py", "models/Encoders/Pooling.py",
    "models/Modules/Attention.py",
    "utils/Manager.py", "utils/utils.py", "utils/MIND.py",
]


def available() -> bool:
    return all(os.path.exists(os.path.join(DEST, f)) for f in FILES)


def build(verbose: bool = False) -> str | None:
    """Copies the files when the reference tree is present (the build container); a no-op elsewhere (the GPU box uses what
    travelled with the snapshot).  Returns the staging directory, or None when neither source nor staged copy exists."""
    if not os.path.isdir(SRC):
        return DEST if available() else None
    for f in FILES:
        src, dst = os.path.join(SRC, f), os.path.join(DEST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            if verbose:
                print("staged", f)
    return DEST


if __name__ == "__main__":
    print(build(verbose=True))

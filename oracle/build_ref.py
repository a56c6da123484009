"""ORACLE tooling (recipe): make the unmodified reference travel to the GPU box.

    python oracle/build_ref.py        (also run by __graft_entry__.build() when /root/reference exists)

The reference is pure Python, so "building" it means packing, verbatim and with their relative paths, the few source
files the hot path consists of from /root/reference into ONE archive, oracle/_ref/ref_hotpath.tar.gz -- in a directory
that is git-ignored (never committed: no reference source enters the repository) but not gpurun-ignored, so it ships
with the snapshot like the built .so files.  `oracle/ref_loader.py` unpacks it into a temporary directory when
/root/reference is absent; `bench.py --impl reference` and the `gpu_eager_baseline` leg then execute the reference's
own ConstraintManager / CaT / term functions / curriculum / PPO() there.  A manifest with the sha256 of every file is
written next to the archive.
"""

from __future__ import annotations

import hashlib
import json
import os
import tarfile

SRC = os.environ.get("CAT_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
U = "exts/cat_envs/cat_envs/tasks/utils"
S12 = "exts/cat_envs/cat_envs/tasks/locomotion/velocity/config/solo12"
FILES = [
    f"{U}/cat/constraint_manager.py",
    f"{U}/cat/constraints.py",
    f"{U}/cat/manager_constraint_cfg.py",
    f"{U}/cat/curriculums.py",
    f"{U}/cleanrl/ppo.py",
    f"{U}/skrl/ppo.py",  # compute_gae of the skrl front-end (cut out with ast by make_golden_gae.py)
    f"{U}/mdp/commands.py",
    f"{U}/mdp/events.py",
    f"{S12}/cat_flat_env_cfg.py",  # ConstraintsCfg / CurriculumCfg: read with ast by the drop-in test
    f"{S12}/agents/clean_rl_ppo_cfg.py",
    "scripts/clean_rl/train.py",
]


def build(verbose: bool = False) -> str | None:
    if not os.path.isdir(SRC):
        return None
    manifest = {}
    os.makedirs(DST, exist_ok=True)
    archive = os.path.join(DST, "ref_hotpath.tar.gz")
    with tarfile.open(archive + ".tmp", "w:gz") as tar:
        for rel in FILES:
            src = os.path.join(SRC, rel)
            if not os.path.isfile(src):
                continue
            tar.add(src, arcname=rel)
            manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    os.replace(archive + ".tmp", archive)
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"[build_ref] {len(manifest)} reference files -> {archive}")
    return archive


if __name__ == "__main__":
    build(verbose=True)

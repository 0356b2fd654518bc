"""Stage the UNMODIFIED reference sources of the hot path under git-ignored ``baseline/_ref/`` so the reference
itself can be timed on the GPU box next to this repo (the GPU box has no
/root/reference, but git-ignored files of /root/repo travel with the snapshot, like the built .so).

TEST / BENCH INFRASTRUCTURE ONLY: nothing under ``causal-gen_b200/`` imports the staged files; only
``oracle/ref_runner.py`` does, for bench.py's reference arms (CPU baseline, ``--impl reference`` and the
reference-eager-on-B200 competitor).  Nothing is copied into tracked paths: ``baseline/_ref/`` is in .gitignore.

    python oracle/stage_reference.py        # run by __graft_entry__.build() when /root/reference exists
"""
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/src"
DST = os.path.join(ROOT, "baseline", "_ref", "src")
FILES = ["vae.py", "dmol.py", "simple_vae.py", "hps.py", "utils.py", "trainer.py", "train_setup.py", "datasets.py",
         "main.py", "pgm/resnet.py", "pgm/layers.py", "pgm/dscm.py"]


def stage() -> bool:
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    for f in FILES:
        os.makedirs(os.path.dirname(os.path.join(DST, f)), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    lic = "/root/reference/LICENSE"
    if os.path.exists(lic):
        shutil.copyfile(lic, os.path.join(os.path.dirname(DST), "LICENSE"))
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference not available", DST)

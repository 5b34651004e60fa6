"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- stages the reference's own model sources for the GPU box.

The reference is pure Python (no setup.py / pyproject: `pip install /root/reference` has nothing to build),
and /root/reference does not exist on the GPU box.  This recipe copies the handful of files that hold the
reference's LRCN classes, unmodified and with their relative paths kept, into the git-ignored directory
`oracle/_ref/` (listed in .gitignore, NOT in .gpurunignore: it travels with the gpurun snapshot like the built
`.so`, and never enters the history).  `oracle/refload.py` then loads the classes from `/root/reference` when it
is present and from `oracle/_ref/` otherwise, so that

  * `bench.py --impl reference` times the reference's OWN `LRCN` class (medsos_lrcn/src/models.py:121-234) on the
    box's host cores (`cpu_baseline.kind = "reference"`), and
  * `bench.py`'s `torch_gpu_baseline` leg runs that same class with stock PyTorch (cuDNN / cuBLASLt) on the B200.

Nothing under video-classif_b200/ imports this module or anything it stages.

    python oracle/build_ref.py          # called by __graft_entry__.build() when /root/reference exists
"""
import os
import shutil

SRC = os.environ.get("B200LRCN_REFERENCE", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

# every file oracle/refload.py reads (class sources are cut out by line range there)
FILES = [
    "medsos_lrcn/src/models.py",            # production LRCN + Mamba block            (models.py:9-234)
    "medsos_lrcn/src/all_config.py",        # the CONF_* globals the model reads
    "medsos_lrcn/src/train_eval.py",        # train_model inner step                   (train_eval.py:9-54)
    "medsos_lrcn/src/loader_data.py",       # uniform_sampling / duplicate_frames      (loader_data.py:35-51)
    "medsos_lrcn/src/models_bidir.py",      # string-programmed Adapt variants         (models_bidir.py:119-248)
    "lrcn/ucf50-lrcn.py",                   # simple-adapt LRCN                        (ucf50-lrcn.py:252-336)
    "lrcn/lrcn.py",                         # crime LRCN                               (lrcn.py:181-305)
    "lrcn/rgb_lrcn.py",                     # rgb LRCN                                 (rgb_lrcn.py:168-263)
    "lrcn/dump_lrcn.py",                    # lstm/gru switch                          (dump_lrcn.py:278-339)
    "lrcn/backup_ucf50.py",                 # LRCN2 small CNN + biGRU                  (backup_ucf50.py:105-151)
    "lrcn/videomamba.py",                   # parallel_scan                            (videomamba.py:242-284)
    "lrcn/.ipynb_checkpoints/LRCN-ucf50-checkpoint.ipynb",   # notebook LRCN           (nb:148-193)
]


def build(verbose: bool = False) -> str:
    """Copies FILES from the reference tree into oracle/_ref/ (no-op without /root/reference).  Returns DST."""
    if not os.path.isdir(os.path.join(SRC, "lrcn")):
        return DST
    for rel in FILES:
        src = os.path.join(SRC, rel)
        dst = os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src) or \
                os.path.getsize(dst) != os.path.getsize(src):
            shutil.copyfile(src, dst)
            os.chmod(dst, 0o644)
            if verbose:
                print("staged", rel)
    return DST


if __name__ == "__main__":
    print(build(verbose=True))

"""One-time converter: reference `.pkl` model files -> flat `.npz` container.

Usage: python tools/convert_models.py [reference_root]
Writes cosmopower_b200/trained_models/CP_paper/{CMB,PK}/*.npz (tracked) and stages byte-identical
copies of the `.pkl` files next to them (git-ignored; they ship to the GPU box with the snapshot).
"""
import os
import shutil
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosmopower_b200 import _loader  # noqa: E402

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELS = {
    "CMB/cmb_PP_PCAplusNN": _loader.PCANN_FIELDS,
    "PK/PKLIN_NN": _loader.NN_FIELDS,
    "PK/PKNLBOOST_NN": _loader.NN_FIELDS,
}


def main(ref_root="/root/reference", stage_pkl=True):
    src_root = os.path.join(ref_root, "cosmopower", "trained_models", "CP_paper")
    dst_root = os.path.join(REPO, "cosmopower_b200", "trained_models", "CP_paper")
    for rel, fields in MODELS.items():
        src = os.path.join(src_root, rel + ".pkl")
        if not os.path.isfile(src):
            continue
        dst = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        attrs = _loader.attributes_from_list(_loader.load_pickle_list(src), fields)
        if not os.path.isfile(dst + ".npz"):
            _loader.save_npz(dst + ".npz", attrs, fields)
        if stage_pkl and not os.path.isfile(dst + ".pkl"):
            shutil.copyfile(src, dst + ".pkl")


if __name__ == "__main__":
    main(*sys.argv[1:2])

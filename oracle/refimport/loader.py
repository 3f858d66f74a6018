"""Imports the UNMODIFIED reference (read-only at /root/reference) in the build container.

Only tests/golden/make_golden.py uses this; nothing that runs on the GPU box may (the reference tree does
not exist there).  `timm==0.4.5` is not installed, so a 6-file stub package next to this file supplies the
constants / re-exports the reference's vendored `models/` package imports.
"""
import contextlib
import io
import os
import sys

REF_ROOT = os.environ.get("ACR_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "DPT"))


def import_reference():
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    with contextlib.redirect_stdout(io.StringIO()):
        import DPT.ACR as ref_acr          # noqa
        import pamr as ref_pamr            # noqa
        import models.vision_transformer as ref_vit  # noqa
    return ref_acr, ref_pamr, ref_vit


def build_acr(num_classes, backbone_name):
    ref_acr, _, _ = import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref_acr.ACR(num_classes, backbone_name, use_pretrain=False)
    return m

"""ORACLE (test infrastructure): the UNMODIFIED upstream code, imported from the staged byte copy in
``oracle/_ref/`` (written by ``oracle/make_ref.py``; git-ignored, travels to the GPU box).

Used by ``tests/`` as the checker at the headline shapes (upstream eager on the same B200, TF32 off)
and by bench.py's ``--impl reference`` / ``cpu_baseline`` / ``gpu_eager_baseline`` legs as the thing
the product is compared WITH.  The product package never imports this.

Import quirks handled here (SURVEY.md section 8(b) gotchas): upstream's ``models/__init__.py`` rebinds
``models.Effi_MVS_plus`` to the class, so the module objects are taken from ``sys.modules``;
``models/module.py:6-7`` needs a top-level ``utils`` importable, i.e. the tree root on ``sys.path``.
"""
from __future__ import annotations

import os
import sys
import types

REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "models", "Effi_MVS_plus.py"))


def load():
    """-> namespace(E=models.Effi_MVS_plus module, M=models.module module, U=models.update module)."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python oracle/make_ref.py` where the upstream checkout is mounted")
    if "models.Effi_MVS_plus" not in sys.modules:
        if REF not in sys.path:
            sys.path.insert(0, REF)
        import models  # noqa: F401  (upstream)
    return types.SimpleNamespace(E=sys.modules["models.Effi_MVS_plus"], M=sys.modules["models.module"],
                                 U=sys.modules["models.update"])


def fusion_module():
    """upstream misc/fusion.py (hard-codes .cuda() at :9-10, so it needs a GPU as it is)."""
    load()
    import misc.fusion as f  # noqa: E402  (upstream)
    return f


def build_model(state_dict=None, ndepths: str = "48,8,8", device="cpu"):
    """Effi_MVS_plus(args).eval() exactly as upstream's test scripts build it (test_tank.py:266-271: the only
    argument fields the model reads are ndepths, GRUiters, CostNum, Effi_MVS_plus.py:321-342)."""
    up = load()
    args = types.SimpleNamespace(ndepths=ndepths, GRUiters="3,3,3", CostNum=3)
    model = up.E.Effi_MVS_plus(args)
    if state_dict is not None:
        res = model.load_state_dict(state_dict, strict=False)
        # update_block.N / CSP_R.N / CSP_C.N are second registrations of update_block_depthN / CSP_RN / CSP_CN
        # (Effi_MVS_plus.py:385-400): a de-duplicated state dict fills them through the first name
        alias = ("update_block.", "CSP_R.", "CSP_C.")
        assert not res.unexpected_keys and all("num_batches" in k or k.startswith(alias) for k in res.missing_keys), res
    return model.to(device).eval()

"""Alias of :mod:`e2fgvi_b200.model` under the reference's import path (``model.e2fgvi``, ``model.e2fgvi_hq``,
``model.modules.*``) so ``test.py`` / ``evaluate.py`` of the reference pick it up with the repo root on sys.path."""

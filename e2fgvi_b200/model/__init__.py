"""Drop-in ``model`` package: ``model.e2fgvi`` / ``model.e2fgvi_hq`` expose ``InpaintGenerator`` exactly as the
reference's ``importlib.import_module('model.' + args.model)`` expects (test.py:117-118, evaluate.py:45-46)."""

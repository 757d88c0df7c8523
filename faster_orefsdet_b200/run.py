"""Run one of the reference's own scripts, unchanged, on the B200 detector:

    python -m faster_orefsdet_b200.run fsod_train_net.py --eval-only --config-file configs/fsod/finetune_vovnet.yaml \\
        MODEL.WEIGHTS model_final.pth
    python -m faster_orefsdet_b200.run demo.py --config-file configs/fsod/finetune_vovnet.yaml --input a.jpg ...

The script's directory goes on ``sys.path`` (as ``python script.py`` would do), the reference's ``fewx`` is imported so
that its registrations happen first (fewx/__init__.py:1), ``faster_orefsdet_b200.install(override=True)`` replaces them,
and the script then runs as ``__main__`` with its own argument list.  Its later ``import fewx`` statements are no-ops
(the modules are already loaded), so ``Trainer.build_model(cfg)`` / ``DefaultPredictor(cfg)`` -> ``build_model(cfg)``
(fsod_train_net.py:95-101, predictor.py:16-37 -> d2!/engine/defaults.py, d2!/modeling/meta_arch/build.py:16-25) resolve
``cfg.MODEL.META_ARCHITECTURE`` to this package's class.
"""
from __future__ import annotations

import os
import runpy
import sys


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m faster_orefsdet_b200.run <reference script.py> [its arguments...]")
    script = os.path.abspath(argv[0])
    sys.path.insert(0, os.path.dirname(script))
    try:
        import fewx  # noqa: F401  (the reference's registrations first)
    except ImportError:
        pass           # a script that does not use fewx: only free names are taken
    import faster_orefsdet_b200
    report = faster_orefsdet_b200.install(override=True)
    print("[faster_orefsdet_b200] registry entries:", report, file=sys.stderr)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()

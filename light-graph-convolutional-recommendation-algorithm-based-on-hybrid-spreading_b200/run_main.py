#!/usr/bin/env python
"""Launcher: drive the B200 drop-in with the REFERENCE's unchanged main.py / const.py.

    python run_main.py --reference /path/to/LGCNHS [--script main.py]

Python puts a script's own directory first on sys.path, so running the reference's main.py
directly would import the reference's model/, utils/, metrics/.  This launcher puts this
directory (our model/, utils/, metrics/, processing/) ahead of the reference tree — which still
provides const.py and main.py — and runpy-executes the script.  The only edit the reference needs
is the one its own README asks for: the env/dataset/model selectors at the bottom of const.py.
"""
import argparse
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("LGCNHS_REFERENCE", "/root/reference"))
    ap.add_argument("--script", default="main.py", help="main.py | findLambda.py | evaluationMetrics.py")
    ap.add_argument("--check-imports", action="store_true",
                    help="only resolve the script's imports and report where each module came from")
    a = ap.parse_args()
    ref = os.path.abspath(a.reference)
    if not os.path.exists(os.path.join(ref, "const.py")):
        raise SystemExit(f"{ref} does not look like the LGCNHS reference (no const.py)")
    sys.path[:0] = [HERE, ref]
    if a.check_imports:
        import importlib

        names = ["const", "utils.log", "processing.handleMovielens", "processing.handleDouban",
                 "model.SpreadMethod.recommend", "model.LightGCN.recommend", "model.LightGCNOpti.recommend",
                 "model.SpreadLightGCN.recommend", "model.SpreadLightGCNOpti.recommend", "utils.trans",
                 "metrics.accurate", "metrics.diversity"]
        for n in names:
            m = importlib.import_module(n)
            print(f"{n:40s} <- {os.path.relpath(m.__file__, HERE) if m.__file__.startswith(HERE) else m.__file__}")
        return
    runpy.run_path(os.path.join(ref, a.script), run_name="__main__")


if __name__ == "__main__":
    main()

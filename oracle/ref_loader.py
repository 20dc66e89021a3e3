"""TEST INFRASTRUCTURE — import the real reference (andykee/lentil) staged under oracle/_ref/ by oracle/build_ref.sh.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs (cpu_baseline, --impl reference) may use this; nothing under
lentil_b200/ does.  Returns None when no copy has been staged (then the callers fall back to the oracle port)."""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference():
    d = os.path.join(_HERE, "_ref")
    if not os.path.isfile(os.path.join(d, "lentil", "__init__.py")):
        return None
    if d not in sys.path:
        sys.path.insert(0, d)
    import lentil
    if not os.path.abspath(lentil.__file__).startswith(d):
        raise ImportError(f"another 'lentil' shadows the staged reference: {lentil.__file__}")
    return lentil

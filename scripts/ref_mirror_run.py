"""Run the reference's own test files against the lentil_b200 mirror classes (tests/ref_mirror_plugin.py) and print the
per-test outcome.  Development aid:  python scripts/ref_mirror_run.py [pytest args]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
FILES = ["tests/test_fourier.py", "tests/test_propagate.py", "tests/test_propagate_fft.py", "tests/test_propagate_mask.py",
         "tests/test_propagate_slice.py", "tests/test_plane.py", "tests/test_wavefront.py", "tests/test_field.py",
         "tests/test_wfe.py", "tests/test_util.py", "tests/test_detector.py", "tests/test_helper.py"]
env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests"), ROOT, os.environ.get("PYTHONPATH", "")]))
sys.exit(subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "ref_mirror_plugin", "-p", "no:cacheprovider", "-rf", "--tb=line"]
                        + (sys.argv[1:] or FILES), cwd=REF, env=env).returncode)

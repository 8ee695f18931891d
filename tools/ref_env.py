"""Make the unmodified reference importable in the build container.

Used only by tools/gen_golden.py and tools/ref_*.py.  /root/reference does not exist on
the GPU box; nothing in tests/, bench.py or the product imports this module."""
import os
import sys
import tempfile
from pathlib import Path

REF = Path("/root/reference")
HERE = Path(__file__).resolve().parent


def setup():
    if not REF.exists():
        raise SystemExit("reference checkout /root/reference not present (build container only)")
    sys.path[:0] = [str(HERE / "ref_shims"), str(REF)]
    # our drop-in package has the same name: make sure the *reference* wins in this process
    for m in [k for k in sys.modules if k == "guided_diffusion" or k.startswith("guided_diffusion.")]:
        del sys.modules[m]
    repo = str(HERE.parent)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != repo]
    sys.path.append(repo)  # flair_b200.synth stays importable (after the reference)
    import scipy.signal
    import scipy.signal.windows
    scipy.signal.gaussian = scipy.signal.windows.gaussian  # removed in scipy>=1.13
    os.chdir(tempfile.mkdtemp())  # imresize() writes rot59.mat into the CWD
    import torch
    torch.set_grad_enabled(False)

"""Import alias: `import sapcu_b200` resolves to the package directory whose (hyphenated) name the build
contract fixes -- c-users-sayakdutta-self-supervised-arbitrary-scale-point-cloud-upsampling-via-snn_b200/."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_PKG = "c-users-sayakdutta-self-supervised-arbitrary-scale-point-cloud-upsampling-via-snn_b200"
_mod = importlib.import_module(_PKG)
sys.modules[__name__] = _mod

"""Drop-in import shim: `import Params` resolves to the B200 backend (put this directory on sys.path)."""
import sys as _sys
from lstm_unet_b200 import Params as _m
_sys.modules[__name__] = _m

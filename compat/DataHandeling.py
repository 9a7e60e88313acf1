"""Drop-in import shim: `import DataHandeling` resolves to the B200 backend's reader mirror (put this directory on sys.path)."""
import sys as _sys
from lstm_unet_b200 import data as _m
_sys.modules[__name__] = _m

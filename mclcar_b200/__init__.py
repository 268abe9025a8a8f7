"""mclcar_b200: B200-native engine for the Monte Carlo likelihood hot path of the R package
shazhe/mclcar.  The product is the C-ABI shared library (include/mclcar_b200.h); `api` is the
host-side mirror of the reference's R functions that binds it through ctypes."""
from . import api  # noqa: F401  (imports _lib: fails loudly when the shared library is missing)

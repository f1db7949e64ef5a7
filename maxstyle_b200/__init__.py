"""maxstyle_b200 -- B200-native MaxStyle feature-style layer (drop-in for
cherise215/MaxStyle `src/advanced/maxstyle.py`) over hand-written sm_100a CUDA kernels."""
from .layer import MaxStyle
from .optim import FusedStyleOptimizer
from .distributed import GlobalBatchMaxStyle, StyleTableExchange

__all__ = ["MaxStyle", "FusedStyleOptimizer", "GlobalBatchMaxStyle", "StyleTableExchange"]
__version__ = "0.1.0"

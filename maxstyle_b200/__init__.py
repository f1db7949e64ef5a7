"""maxstyle_b200 -- B200-native MaxStyle feature-style layer (drop-in for
cherise215/MaxStyle `src/advanced/maxstyle.py`) over hand-written sm_100a CUDA kernels."""
from .layer import MaxStyle
from .optim import FusedStyleOptimizer
from .mixstyle import MixStyle
from .distributed import GlobalBatchMaxStyle, StyleTableExchange, PeerTableExchange
from .host_pipeline import HostStepPipeline, HostStepResult
from .executor import StyleLoopExecutor
from .graphed import GraphedLayerStep
from .losses import cross_entropy_2D
from .fused import apply_max_style_fused, rescale_intensity

__all__ = ["MaxStyle", "MixStyle", "FusedStyleOptimizer", "GlobalBatchMaxStyle", "StyleTableExchange", "PeerTableExchange",
           "HostStepPipeline", "HostStepResult", "StyleLoopExecutor", "GraphedLayerStep", "cross_entropy_2D", "apply_max_style_fused",
           "rescale_intensity"]
__version__ = "0.1.0"

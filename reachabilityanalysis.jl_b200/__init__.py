"""reach-b200: B200-native LGG09 / GLGM06 set-propagation hot path.

The device path lives in libreach_cuda.so (csrc/, C ABI in include/reach.h); this package is the
host-side mirror of the reference's operator interface for that path."""
from . import sets  # noqa: F401
from .sets import *  # noqa: F401,F403
from .directions import BoxDirections, OctDirections, CustomDirections, PolarDirections  # noqa: F401
from .flowpipe import (Flowpipe, MixedFlowpipe, ReachSolution, TemplateReachSet, ReachSet,  # noqa: F401
                       TimeInterval, zeroT, rho, sigma, shift)
from .lgg09 import LGG09, Forward, NoBloating, FirstOrderZonotope, Lgg09Plan  # noqa: F401
from .glgm06 import GLGM06, Glgm06Plan, reduce_order  # noqa: F401
from .box import BOX  # noqa: F401
from .solve import solve, IVP, discretize  # noqa: F401
from ._lib import Context, ReachError, default_context  # noqa: F401

"""biceps_b200 -- B200-native BICePs posterior-sampling hot path behind the reference's API.

`import biceps_b200 as biceps` gives the names the reference exports on this path
(/root/reference/biceps/__init__.py:8-22): Ensemble, get_restraint_options, PosteriorSampler,
PosteriorSamplingTrajectory, Analysis, find_all_state_sampled_time, Convergence, Preparation, toolbox; the restraint classes
live at biceps.Restraint.Restraint_{cs,J,noe,pf} like in the reference.  Every compute step runs in
the CUDA library (lib/libbiceps_b200.so); there is no CPU fallback.
"""
__version__ = "0.1.0"
name = "biceps_b200"

from . import toolbox                                          # noqa: F401
from . import Restraint                                        # noqa: F401  (module, as in the reference)
from .Restraint import Preparation, Ensemble, get_restraint_options        # noqa: F401
from .PosteriorSampler import PosteriorSampler, PosteriorSamplingTrajectory   # noqa: F401
from .Analysis import Analysis, find_all_state_sampled_time   # noqa: F401
from .convergence import Convergence                           # noqa: F401

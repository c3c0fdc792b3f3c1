from .minimal_distortion_principle import minimal_distortion_principle  # noqa: F401
from .permutation_alignment import correlation_based_permutation_solver  # noqa: F401
from .projection_back import projection_back  # noqa: F401

PROJECTION_BACK_KEYWORDS = ["projection_back", "projection-back", "PB"]
MINIMAL_DISTORTION_PRINCIPLE_KEYWORDS = ["minimal_distortion_principle", "minimal-distortion-principle", "MDP"]

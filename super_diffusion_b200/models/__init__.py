from . import utils  # noqa: F401
from . import ddpm  # noqa: F401  (registers 'score-net')

from . import utils  # noqa: F401
from . import ddpm  # noqa: F401  (registers 'score-net')
from . import toy_mlp  # noqa: F401  (registers 'toy-mlp', the notebook's score model)

from models.networks.editline_g import *   # noqa: F401,F403
from models.networks.editline2_g import *  # noqa: F401,F403

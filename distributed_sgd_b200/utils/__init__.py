from .config import Config, load_config  # noqa: F401
from .dataset import Data, rcv1, synthetic_rcv1, write_rcv1  # noqa: F401

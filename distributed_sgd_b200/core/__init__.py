from .group import Group  # noqa: F401
from .master import Master, MasterAsync, MasterSync  # noqa: F401
from .slave import Slave  # noqa: F401

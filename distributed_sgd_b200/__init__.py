"""distributed_sgd_b200 -- B200-native data-parallel SGD hot path behind the surface of
zifeo/distributed-sgd's Slave / Master / SparseSVM (include/dsgd.h).

(The directory is named with an underscore because `distributed-sgd_b200` is not an importable
Python package name.)
"""
from . import native  # noqa: F401
from .core import Master, MasterAsync, MasterSync, Slave  # noqa: F401
from .ml import EarlyStopping, GradState, SparseSVM, SplitStrategy  # noqa: F401
from .utils import Config, Data, load_config, rcv1, synthetic_rcv1  # noqa: F401

__all__ = ["native", "Master", "MasterAsync", "MasterSync", "Slave", "EarlyStopping", "GradState", "SparseSVM",
           "SplitStrategy", "Config", "Data", "load_config", "rcv1", "synthetic_rcv1"]

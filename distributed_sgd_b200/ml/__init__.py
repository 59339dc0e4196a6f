from . import early_stopping as EarlyStopping  # noqa: F401,N812
from . import split_strategy as SplitStrategy  # noqa: F401,N812
from .grad_state import GradState  # noqa: F401
from .sparse_svm import SparseSVM  # noqa: F401

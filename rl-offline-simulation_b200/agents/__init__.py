from . import tabular  # noqa: F401
from .agent import Agent  # noqa: F401

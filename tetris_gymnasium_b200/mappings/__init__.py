from .actions import ActionsMapping  # noqa: F401
from .rewards import RewardsMapping  # noqa: F401

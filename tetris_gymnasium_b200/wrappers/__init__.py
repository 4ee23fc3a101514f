from .grouped import GroupedActionsObservations  # noqa: F401
from .observation import FeatureVectorObservation, RgbObservation  # noqa: F401

from .grouped import GroupedActionsObservations  # noqa: F401
from .observation import CnnObservation, FeatureVectorObservation, RgbObservation  # noqa: F401
from .stats import RecordEpisodeStatistics  # noqa: F401

from .configuration import Configuration  # noqa: F401
from .features import Feature, Features  # noqa: F401
from .config_proto import FeatureDeal, FeaturePooling, FeatureTower  # noqa: F401

"""reach-b200: B200-native LGG09 / GLGM06 set-propagation hot path."""
from . import sets  # noqa: F401

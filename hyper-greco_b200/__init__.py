"""hyper-greco GKR proving hot path, B200-native. See DESIGN.md."""
from . import params, witness  # noqa: F401

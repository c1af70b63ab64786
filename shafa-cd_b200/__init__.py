"""shafa-cd_b200: B200-native Shafa F/T/C/D hot path."""
from . import datagen  # noqa: F401

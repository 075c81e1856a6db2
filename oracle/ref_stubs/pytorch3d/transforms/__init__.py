from . import rotation_conversions

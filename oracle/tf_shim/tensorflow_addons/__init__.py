"""NumPy stand-in for tensorflow_addons.image.rotate (tfa 0.17.1).  TEST INFRASTRUCTURE ONLY."""
from . import image  # noqa: F401

"""`import tina` drop-in alias: existing tina.Scene scripts pick up the B200-native raster path."""
from taichi_three_b200 import *  # noqa: F401,F403
from taichi_three_b200 import __version__  # noqa: F401

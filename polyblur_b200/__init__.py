"""polyblur_b200 -- B200-native Polyblur engine with the public surface of teboli/polyblur.

    from polyblur_b200 import polyblur_deblurring, PolyblurDeblurring

(`polyblur/__init__.py:1` of the reference re-exports the same two names.)
"""
from .deblurring import GraphedPolyblur, PolyblurDeblurring, clear_cache, polyblur_deblurring  # noqa: F401
from . import blur_estimation, compat, deblurring, domain_transform, edgetaper, filters, io, utils  # noqa: F401

__version__ = "0.1.0"

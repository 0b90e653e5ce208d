"""eosvos_b200: B200-native (sm_100a) implementation of the e-OSVOS per-video fine-tuning and
inference hot path, behind the reference's own Python API.

The directory is named ``e-osvos_b200`` (the layout the build contract asks for); import it as
``eosvos_b200`` through the loader module of that name at the repository root.
"""
__version__ = "0.1.0"

"""Offline dataset products of memflow/read_data/Dataset_Parton_Level.py:400-463 on the K2 kernel."""
from .products import phasespace_products  # noqa: F401

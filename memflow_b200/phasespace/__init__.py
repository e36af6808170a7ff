"""Drop-in mirror of memflow.phasespace (reference: memflow/phasespace/__init__.py is empty)."""

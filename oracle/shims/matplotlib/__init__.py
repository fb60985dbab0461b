"""Empty matplotlib stand-in (main.py:11, Visualiser.py:8 import pyplot only)."""

"""Empty shell (src/preprocessing/chunker.py:6 imports it)."""

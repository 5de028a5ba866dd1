"""Empty shell (src/preprocessing/chunker.py:6-7 imports it)."""

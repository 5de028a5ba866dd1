"""B200-native hybrid retriever: drop-in for Domain-RAG-System's ``src/indexing`` +
``src/reranker`` call surface, backed by hand-written sm_100a CUDA behind a C ABI
(``include/drs_b200.h``).  PyTorch only owns the device buffers."""

__version__ = "0.1.0"


def load_library():
    """Load (and return) the ctypes handle of the CUDA library; ImportError if not built."""
    from . import _lib
    return _lib.load()

"""tidk-b200: the data-parallel hot path of tolkit/telomeric-identifier (tidk 0.2.7)
on NVIDIA B200 (sm_100a).

The compute lives in ``libtidk_cuda.so`` (hand-written CUDA behind the C ABI in
``include/tidk_cuda.h``).  This package is the Python host mirror of the
reference's three drivers -- ``search::search`` (search.rs:10), ``finder::finder``
(finder.rs:13) and ``explore::explore`` (explore.rs:20) -- used by the tests and
``bench.py``; the C++ ``tidk`` driver under ``csrc/host`` is the command-line host.
There is no CPU fallback: importing works anywhere, computing needs the library
and a GPU and fails loudly without them.
"""
from .api import Context, SeqBatch, TidkCudaError, library_path  # noqa: F401
from .search import search, search_rows  # noqa: F401
from .finder import finder, finder_rows  # noqa: F401
from .explore import explore, explore_table  # noqa: F401
from .fasta import read_fasta  # noqa: F401
from .clades import return_telomere_sequence, get_clades  # noqa: F401

__version__ = "0.1.0"

"""nxsearch-b200: a B200-native query-scoring engine behind nxsearch's C API.

The product is ``lib/libnxsearch.so`` (C11 host + sm_100a CUDA, built by
``nxsearch_b200._build``).  The Python modules are thin ctypes mirrors of its
C interfaces, used by the tests, the benchmark and the multi-GPU driver:

* :mod:`nxsearch_b200.capi`   -- the public nxs_* API (``include/nxs.h``)
* :mod:`nxsearch_b200.engine` -- the GPU engine C ABI (``include/nxsb200_gpu.h``)
* :mod:`nxsearch_b200.tools`  -- synthetic corpus + index file tooling
  (``include/nxsb200_tools.h``)
* :mod:`nxsearch_b200.dist`   -- document-sharded search over torch.distributed
"""

from ._lib import load_library, library_path  # noqa: F401

__all__ = ["load_library", "library_path"]
